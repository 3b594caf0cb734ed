"""TEST INFRASTRUCTURE ONLY — CPU restatement of ``Speech_Rate_Predictor`` (src/sub_model/modules.py:108-142) on top
of the TransformerEncoder restatement of oracle/avhubert_oracle.py (fairseq/fairseq/models/wav2vec/wav2vec2.py:816-902).
Pinned against the REAL class: oracle/make_golden_sr.py executes the class definition straight from the reference file
(with the reference's own TransformerEncoder, loaded by oracle/ref_import.py) and stores input / state dict / output in
tests/golden/sr_predictor.npz; tests/test_oracle_vs_reference.py re-checks live when /root/reference is present."""
import torch
import torch.nn as nn

from .avhubert_oracle import OracleConfig, _Encoder


class OracleSpeechRatePredictor(nn.Module):
    def __init__(self, num_layers):
        super().__init__()
        cfg = OracleConfig(encoder_layers=num_layers, encoder_embed_dim=256, encoder_ffn_embed_dim=1024,
                           encoder_attention_heads=4, layer_norm_first=True, conv_pos=128, conv_pos_groups=16)
        self.sr_token = nn.Parameter(torch.zeros(1, 1, 256))          # modules.py:127-128
        nn.init.xavier_uniform_(self.sr_token)
        self.linear = nn.Linear(1024, 256)                            # :129
        self.encoder = _Encoder(cfg)                                  # :130
        self.sr_predictor = nn.Linear(256, 1)                         # :131
        self.activation = nn.ReLU()                                   # :132

    def forward(self, x):                                             # :134-142
        x = self.linear(x)
        x = torch.cat([self.sr_token.expand(x.size(0), -1, -1), x], dim=1)
        x = self.encoder(x)
        return self.activation(self.sr_predictor(x[:, 0, :]))


def build(num_layers=2, seed=77):
    torch.manual_seed(seed)
    m = OracleSpeechRatePredictor(num_layers).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():       # non-trivial LayerNorm affines / biases so that every term is exercised
        for mod in m.modules():
            if isinstance(mod, nn.LayerNorm):
                mod.weight.copy_(1 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.Linear):
                mod.bias.copy_(0.05 * torch.randn(mod.bias.shape, generator=g))
        m.sr_predictor.weight.copy_(0.3 * torch.randn(m.sr_predictor.weight.shape, generator=g))
        m.sr_predictor.bias.fill_(9.0)
    return m


def synthetic_features(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, 1024, generator=g)
