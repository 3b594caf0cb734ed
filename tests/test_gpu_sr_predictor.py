"""the wav2vec2 TransformerEncoder as a stand-alone module (avh_encoder_forward) and the
Speech_Rate_Predictor built on it (src/sub_model/modules.py:108-142), CUDA path vs the CPU oracle (pinned to the real
reference class by tests/golden/sr_predictor.npz)."""
import numpy as np
import os
import pytest
import torch

from oracle import avhubert_oracle as ao
from oracle import sr_oracle

from helpers import GOLDEN, cosine, rel_err

pytestmark = pytest.mark.gpu


def _device_sr(oracle, dtype):
    from multimodalvc_b200.sr_predictor import Speech_Rate_Predictor
    m = Speech_Rate_Predictor(len(oracle.encoder.layers))
    m.load_state_dict(oracle.state_dict(), strict=True)
    return m.to("cuda", dtype).eval()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 2e-2)])
def test_speech_rate_predictor_matches_reference_golden(dtype, tol):
    z = np.load(os.path.join(GOLDEN, "sr_predictor.npz"))
    layers, seed, B, T, xs = [int(v) for v in z["meta"]]
    oracle = sr_oracle.build(layers, seed=seed)
    x = sr_oracle.synthetic_features(B, T, seed=xs)
    y = _device_sr(oracle, dtype)(x.to("cuda", dtype)).float().cpu()
    assert y.shape == (B, 1)
    assert np.abs(y.numpy() - z["y"]).max() < tol * np.abs(z["y"]).max()


@pytest.mark.parametrize("T", [1, 97, 150, 299])
def test_speech_rate_predictor_long_inputs_vs_oracle(T):
    oracle = sr_oracle.build(2, seed=3)
    x = sr_oracle.synthetic_features(4, T, seed=T)
    with torch.no_grad():
        y_ref = oracle(x)
    y = _device_sr(oracle, torch.float32)(x.cuda()).cpu()
    assert rel_err(y, y_ref) < 2e-3
    yb = _device_sr(oracle, torch.bfloat16)(x.cuda().bfloat16()).float().cpu()
    assert rel_err(yb, y_ref) < 3e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_transformer_encoder_module_with_padding_mask_and_layer(dtype):
    """TransformerEncoder.forward(x, padding_mask, layer) on its own: Base width, ragged mask, early exit (0-based
    tgt_layer) and the full stack, against the oracle's restatement of wav2vec2.py:859-902."""
    from types import SimpleNamespace
    from multimodalvc_b200.sr_predictor import TransformerEncoder
    cfg = ao.OracleConfig(encoder_layers=3, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12)
    torch.manual_seed(5)
    enc_ref = ao._Encoder(cfg).eval()
    ao.randomize_norm_stats(enc_ref)
    enc = TransformerEncoder(SimpleNamespace(encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12,
                                             encoder_layers=3, conv_pos=128, conv_pos_groups=16, layer_norm_first=True,
                                             activation_fn="gelu"))
    enc.load_state_dict(enc_ref.state_dict(), strict=True)
    enc = enc.to("cuda", dtype).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 170, 768, generator=g)
    pm = torch.zeros(3, 170, dtype=torch.bool)
    pm[1, 100:] = True
    pm[2, 7:] = True
    for layer in (None, 0, 1):
        with torch.no_grad():
            y_ref = enc_ref(x, pm, layer)
        y, _ = enc(x.to("cuda", dtype), pm.cuda(), layer=layer)
        y = y.float().cpu()
        if dtype == torch.float32:
            assert rel_err(y, y_ref) < 2e-3, layer
        else:
            assert cosine(y, y_ref) > 0.999, layer
