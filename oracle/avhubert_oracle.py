"""TEST INFRASTRUCTURE ONLY — fp32 CPU/PyTorch oracle of the AV-HuBERT encoder hot path.

A plain-PyTorch *restatement* (not a copy) of what the reference computes behind
`AVHubertModel.extract_finetune(source, padding_mask)`:

  avhubert/hubert.py:317-332   SubModel (optional ResEncoder -> Linear proj on [B,C,T])
  avhubert/hubert.py:694-745   extract_finetune (zero-fill missing modality, concat/add, LayerNorm,
                               post_extract_proj, encoder)
  avhubert/hubert.py:564-574   forward_padding_mask
  avhubert/resnet.py:35-74     BasicBlock (conv-BN-PReLU-conv-BN-(+res)-PReLU)
  avhubert/resnet.py:77-129    ResNet-18 trunk, :131-169 ResEncoder (Conv3d stem + BN3d + PReLU + MaxPool3d)
  fairseq/fairseq/models/wav2vec/wav2vec2.py:816-902  TransformerEncoder (pad zeroing, weight-normed grouped
                               pos-conv + SamePad + GELU, residual, layer loop, final LN when pre-LN)
  fairseq/fairseq/models/wav2vec/wav2vec2.py:960-1014 TransformerSentenceEncoderLayer (pre-/post-LN)
  fairseq/fairseq/modules/multihead_attention.py:170-192 -> F.multi_head_attention_forward with separate
                               q/k/v weights, concatenated biases, key_padding_mask, need_weights=True
  fairseq/fairseq/modules/gelu.py:95-96  erf GELU computed in fp32
  fairseq/fairseq/modules/same_pad.py:16-21 drop the last step for even kernels

State-dict key names equal the reference's so a reference state dict loads with strict=True
(minus `mask_emb`, `final_proj.*`, `label_embs_concat`, which the path never touches).

Pinned: tests/test_oracle_vs_reference.py loads the REAL reference model (oracle/ref_import.py) here in
this container, copies its state dict into this oracle and requires equal outputs; the golden fixtures in
tests/golden/ were produced by the real reference (oracle/make_golden.py).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class OracleConfig:
    encoder_layers: int = 12
    encoder_embed_dim: int = 768
    encoder_ffn_embed_dim: int = 3072
    encoder_attention_heads: int = 12
    audio_feat_dim: int = 104
    modality_fuse: str = "concat"
    layer_norm_first: bool = True
    conv_pos: int = 128
    conv_pos_groups: int = 16

    @staticmethod
    def named(size, **kw):
        shape = dict(base=(12, 768, 3072, 12), large=(24, 1024, 4096, 16), tiny=(2, 128, 256, 2))[size]
        cfg = OracleConfig(*shape)
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu1 = nn.PReLU(cout)
        self.relu2 = nn.PReLU(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.relu1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        r = x if self.downsample is None else self.downsample(x)
        return self.relu2(y + r)


class _Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        widths = [64, 128, 256, 512]
        cin = 64
        for i, w in enumerate(widths):
            stride = 1 if i == 0 else 2
            setattr(self, f"layer{i + 1}", nn.Sequential(_Block(cin, w, stride), _Block(w, w, 1)))
            cin = w
        for m in self.modules():   # resnet.py:92-98
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))

    def forward(self, x):
        for i in range(4):
            x = getattr(self, f"layer{i + 1}")(x)
        return x.mean(dim=(2, 3))     # AdaptiveAvgPool2d(1) + view


class _ResEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False),
            nn.BatchNorm3d(64),
            nn.PReLU(64),
            nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)),
        )
        self.trunk = _Trunk()

    def forward(self, x):                      # [B,1,T,88,88]
        B, _, T = x.shape[:3]
        x = self.frontend3D(x)                 # [B,64,T,22,22]
        x = x.transpose(1, 2).reshape(B * T, 64, x.shape[3], x.shape[4])
        x = self.trunk(x)                      # [B*T,512]
        return x.view(B, T, 512).transpose(1, 2)   # [B,512,T]


class _SubModel(nn.Module):
    def __init__(self, resnet, input_dim, dim):
        super().__init__()
        self.resnet = resnet
        self.proj = nn.Linear(input_dim, dim)

    def forward(self, x):
        if self.resnet is not None:
            x = self.resnet(x)
        return self.proj(x.transpose(1, 2)).transpose(1, 2)   # [B,D,T]


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.dim, self.heads = dim, heads
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)

    def forward(self, x, key_padding_mask):    # x: [T,B,D]
        out, _ = F.multi_head_attention_forward(
            x, x, x, self.dim, self.heads, torch.empty([0]),
            torch.cat((self.q_proj.bias, self.k_proj.bias, self.v_proj.bias)),
            None, None, False, 0.0, self.out_proj.weight, self.out_proj.bias,
            False, key_padding_mask, True, None,
            use_separate_proj_weight=True, q_proj_weight=self.q_proj.weight,
            k_proj_weight=self.k_proj.weight, v_proj_weight=self.v_proj.weight)
        return out


class _Layer(nn.Module):
    def __init__(self, dim, ffn, heads, ln_first):
        super().__init__()
        self.ln_first = ln_first
        self.self_attn = _Attn(dim, heads)
        self.self_attn_layer_norm = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, ffn)
        self.fc2 = nn.Linear(ffn, dim)
        self.final_layer_norm = nn.LayerNorm(dim)

    def _ffn(self, x):
        h = F.gelu(self.fc1(x).float()).type_as(x)
        return self.fc2(h)

    def forward(self, x, kpm):
        if self.ln_first:
            x = x + self.self_attn(self.self_attn_layer_norm(x), kpm)
            x = x + self._ffn(self.final_layer_norm(x))
        else:
            x = self.self_attn_layer_norm(x + self.self_attn(x, kpm))
            x = self.final_layer_norm(x + self._ffn(x))
        return x


class _PosConv(nn.Module):
    """weight_norm(dim=2) Conv1d: parameters weight_g [1,1,K], weight_v [D, D/groups, K], bias [D]."""

    def __init__(self, dim, k, groups):
        super().__init__()
        self.k, self.groups = k, groups
        std = math.sqrt(4.0 / (k * dim))
        v = torch.randn(dim, dim // groups, k) * std
        self.weight_g = nn.Parameter(v.norm(dim=(0, 1), keepdim=True).clone())
        self.weight_v = nn.Parameter(v)
        self.bias = nn.Parameter(torch.zeros(dim))

    def weight(self):
        v = self.weight_v
        return v * (self.weight_g / v.norm(dim=(0, 1), keepdim=True))

    def forward(self, x):                      # [B,D,T]
        y = F.conv1d(x, self.weight(), self.bias, padding=self.k // 2, groups=self.groups)
        if self.k % 2 == 0:
            y = y[:, :, :-1]
        return F.gelu(y)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg.encoder_embed_dim
        self.embedding_dim = D
        self.layer_norm_first = cfg.layer_norm_first
        self.pos_conv = nn.Sequential(_PosConv(D, cfg.conv_pos, cfg.conv_pos_groups))
        self.layers = nn.ModuleList(
            [_Layer(D, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads, cfg.layer_norm_first)
             for _ in range(cfg.encoder_layers)])
        self.layer_norm = nn.LayerNorm(D)
        for m in self.modules():               # init_bert_params
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, 0.02)
                m.bias.data.zero_()

    def forward(self, x, padding_mask=None, layer=None):   # x [B,T,D]
        if padding_mask is not None:
            x = x.masked_fill(padding_mask.unsqueeze(-1), 0.0)
        x = x + self.pos_conv(x.transpose(1, 2)).transpose(1, 2)
        if not self.layer_norm_first:
            x = self.layer_norm(x)
        x = x.transpose(0, 1)
        skip = getattr(self, "layer_skip", None)       # LayerDrop decisions of a training-mode forward (wav2vec2.py:886-888)
        for i, lyr in enumerate(self.layers):
            if not (skip and skip[i]):
                x = lyr(x, padding_mask)
            if i == layer:
                break
        x = x.transpose(0, 1)
        if self.layer_norm_first and layer is None:
            x = self.layer_norm(x)
        return x


class OracleAVHubert(nn.Module):
    def __init__(self, cfg: OracleConfig):
        super().__init__()
        self.cfg = cfg
        D = cfg.encoder_embed_dim
        self.feature_extractor_audio = _SubModel(None, cfg.audio_feat_dim, D)
        self.feature_extractor_video = _SubModel(_ResEncoder(), 512, D)
        self.embed = 2 * D if cfg.modality_fuse == "concat" else D
        self.post_extract_proj = nn.Linear(self.embed, D) if self.embed != D else None
        self.encoder = _Encoder(cfg)
        self.layer_norm = nn.LayerNorm(self.embed)

    @staticmethod
    def forward_padding_mask(features, padding_mask):
        extra = padding_mask.size(1) % features.size(1)
        if extra > 0:
            padding_mask = padding_mask[:, :-extra]
        return padding_mask.view(padding_mask.size(0), features.size(1), -1).all(-1)

    def stage_outputs(self, source, padding_mask=None, output_layer=None):
        """Like extract_finetune but also returns intermediates (for per-stage kernel parity tests)."""
        out = {}
        a, v = source["audio"], source["video"]
        D = self.cfg.encoder_embed_dim
        if v is not None:
            res = self.feature_extractor_video.resnet(v)
            out["resnet"] = res                                # [B,512,T]
            fv = self.feature_extractor_video.proj(res.transpose(1, 2)).transpose(1, 2)
        if a is not None:
            fa = self.feature_extractor_audio(a)
        if v is None:
            fv = fa.new_zeros(fa.size(0), D, fa.size(-1))
        if a is None:
            fa = fv.new_zeros(fv.size(0), D, fv.size(-1))
        feats = torch.cat([fa, fv], dim=1) if self.cfg.modality_fuse == "concat" else fa + fv
        feats = self.layer_norm(feats.transpose(1, 2))
        out["fused_ln"] = feats
        if padding_mask is not None:
            padding_mask = self.forward_padding_mask(feats, padding_mask)
        if self.post_extract_proj is not None:
            feats = self.post_extract_proj(feats)
        out["enc_in"] = feats
        x = self.encoder(feats, padding_mask, None if output_layer is None else output_layer - 1)
        out["x"] = x
        return out, padding_mask

    @torch.no_grad()
    def extract_finetune(self, source, padding_mask=None, mask=False, ret_conv=False, output_layer=None):
        out, pm = self.stage_outputs(source, padding_mask, output_layer)
        return out["x"], pm


def _oracle_extract_features(self, source, padding_mask=None, mask=False, ret_conv=False, output_layer=None):
    """avhubert/hubert.py:676-692 -> forward(features_only=True) (:591-653) in eval mode with mask=False: both
    modalities are required (forward_features on None fails in the reference); `features` is the post_extract_proj
    output, whose padded frames the encoder zeroes IN PLACE (index_put, wav2vec2.py:869-870), `x` the encoder
    output (layer k = output_layer, no final LayerNorm then)."""
    if source["audio"] is None or source["video"] is None:
        raise ValueError("extract_features needs both modalities (hubert.py:609-610)")
    out, pm = self.stage_outputs(source, padding_mask, output_layer)
    feats = out["enc_in"]
    if pm is not None:
        feats = feats.masked_fill(pm.unsqueeze(-1), 0.0)
    return (feats if ret_conv else out["x"]), pm


OracleAVHubert.extract_features = torch.no_grad()(_oracle_extract_features)


def randomize_norm_stats(model: nn.Module, seed: int = 4321):
    """Make BN folding / PReLU non-trivial for parity tests: running_mean ~ N(0,0.1),
    running_var ~ U(0.5,1.5), BN affine ~ N(1,0.1)/N(0,0.1), PReLU slopes ~ U(0.1,0.4), LN affine perturbed."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        elif isinstance(m, nn.PReLU):
            m.weight.data.copy_(0.1 + 0.3 * torch.rand(m.weight.shape, generator=g))
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        elif isinstance(m, nn.Linear):
            m.bias.data.copy_(0.02 * torch.randn(m.bias.shape, generator=g))
    for n, p in model.named_parameters():
        if n.endswith("pos_conv.0.bias"):
            p.data.copy_(0.02 * torch.randn(p.shape, generator=g))
        if n.endswith("pos_conv.0.weight_g"):
            p.data.mul_(1 + 0.1 * torch.randn(p.shape, generator=g))
    return model


def build_oracle(size="base", seed=1234, randomize=True, **kw):
    torch.manual_seed(seed)
    model = OracleAVHubert(OracleConfig.named(size, **kw)).eval()
    if randomize:
        randomize_norm_stats(model)
    return model


def synthetic_inputs(B, T, lengths=None, seed=0, audio=True, video=True):
    """video randn (normalised gray frames), audio feats randn; pad frames zero like the
    collater (hubert_dataset.py:442-447); padding_mask True = padded."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, 1, T, 88, 88, generator=g) if video else None
    a = torch.randn(B, T, 104, generator=g) if audio else None
    pm = None
    if lengths is not None:
        pm = torch.zeros(B, T, dtype=torch.bool)
        for i, n in enumerate(lengths):
            pm[i, n:] = True
            if v is not None:
                v[i, :, n:] = 0
            if a is not None:
                a[i, n:] = 0
    if a is not None:
        a = a.transpose(1, 2)     # non-contiguous [B,104,T] view, like collater_audio (hubert_dataset.py:453)
    return {"audio": a, "video": v}, pm
