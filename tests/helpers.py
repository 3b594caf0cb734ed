"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import ast
import hashlib
import os

import numpy as np
import torch

from oracle import avhubert_oracle as ao

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def load_encoder_case(name):
    """Returns dict(oracle, src, pm, output_layer, y_ref, pm_ref, over, size) rebuilt from a golden fixture."""
    z = np.load(os.path.join(GOLDEN, f"enc_{name}.npz"))
    size, B, T, lengths, audio, video, output_layer, over, chk = [str(x) for x in z["meta"]]
    B, T = int(B), int(T)
    lengths = ast.literal_eval(lengths)
    output_layer = ast.literal_eval(output_layer)
    over = ast.literal_eval(over)
    oracle = ao.build_oracle(size, seed=1234, **over)
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=11, audio=bool(int(audio)), video=bool(int(video)))
    return dict(oracle=oracle, src=src, pm=pm, output_layer=output_layer, y_ref=torch.from_numpy(z["y"]),
                pm_ref=z["pm_out"], over=over, size=size, checksum=chk)


def rel_err(a, b):
    """max |a-b| / max |b| — the 'relative' error of the fp32 gate."""
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm())).item()


def make_device_model(oracle, over, size, dtype, device="cuda", **cfg_kw):
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    cfg = AVHubertConfig.named(size, **over, **cfg_kw)
    m = AVHubertModel(cfg)
    m.remove_pretraining_modules()
    missing = m.load_state_dict(oracle.state_dict(), strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"mask_emb"}, missing
    return m.to(device=device, dtype=dtype).eval()


def to_dev(src, pm, device="cuda", dtype=None):
    out = {}
    for k, v in src.items():
        if v is None:
            out[k] = None
        else:
            v = v.to(device)
            out[k] = v.to(dtype) if dtype is not None else v
    return out, (pm.to(device) if pm is not None else None)


def load_pretrain_case(name):
    """Golden of the REAL AVHubertModel.forward(mask=True) (oracle/make_golden_pretrain.py): returns dict(oracle, head,
    src, pm, targets, over, n_dicts, z) with the seeded encoder weights rebuilt and the stored head weights."""
    from oracle import make_golden_pretrain as mg
    from oracle import pretrain_oracle as po
    B, T, lengths, over, n_dicts = mg.CASES[name]
    z = np.load(os.path.join(GOLDEN, f"pretrain_{name}.npz"))
    oracle = ao.build_oracle("tiny", seed=1234)
    assert state_checksum(oracle.state_dict()) == str(z["checksum"]), "seeded weights changed: regenerate goldens"
    src, pm, targets = mg.case_inputs(name, B, T, lengths, n_dicts)
    head = po.Head(z["mask_emb"], z["final_proj_w"], z["final_proj_b"], z["label_embs"], mg.NUM_CLASSES[:n_dicts], **over)
    return dict(oracle=oracle, head=head, src=src, pm=pm, targets=targets, over=over, n_dicts=n_dicts, z=z)


def make_pretrain_device_model(c, dtype, device="cuda"):
    """The product's AVHubertModel with the pretraining head of a golden case (dictionaries given, final_dim 32)."""
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    from oracle import make_golden_pretrain as mg
    import types
    cfg = AVHubertConfig.named("tiny", final_dim=mg.FINAL_DIM, **c["over"])
    m = AVHubertModel(cfg, types.SimpleNamespace(sample_rate=25), [list(range(n)) for n in mg.NUM_CLASSES[:c["n_dicts"]]])
    sd = dict(c["oracle"].state_dict())
    z = c["z"]
    sd["mask_emb"] = torch.from_numpy(z["mask_emb"])
    sd["final_proj.weight"], sd["final_proj.bias"] = torch.from_numpy(z["final_proj_w"]), torch.from_numpy(z["final_proj_b"])
    sd["label_embs_concat"] = torch.from_numpy(z["label_embs"])
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and not missing.missing_keys, missing
    return m.to(device=device, dtype=dtype).eval()


class _GradMultiply(torch.autograd.Function):
    """fairseq/modules/grad_multiply.py: identity forward, gradient scaled."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


def oracle_finetune_graph(o, src, pm, fgm, fuse):
    """extract_finetune of the reference in .train() as a differentiable graph on the ORACLE's modules
    (avhubert/hubert.py:538-547,694-745): feature_grad_mult > 0 -> the extractors differentiate and GradMultiply scales the
    gradient entering them; <= 0 -> extractors under no_grad; missing modality -> zeros; concat / add fusion; LayerNorm;
    post_extract_proj; encoder.  Pinned to the REAL model's autograd by tests/golden/train_grads_tiny.npz
    (oracle/make_golden_grads.py, tests/test_oracle_vs_reference.py)."""
    def features(extractor, x):
        if x is None:
            return None
        if fgm > 0:
            f = extractor(x)
            return _GradMultiply.apply(f, fgm) if fgm != 1.0 else f
        with torch.no_grad():
            return extractor(x)

    fv = features(o.feature_extractor_video, src.get("video"))
    fa = features(o.feature_extractor_audio, src.get("audio"))
    if fv is None:
        fv = torch.zeros_like(fa)
    if fa is None:
        fa = torch.zeros_like(fv)
    fused = (torch.cat([fa, fv], dim=1) if fuse == "concat" else fa + fv).transpose(1, 2)
    feats = o.layer_norm(fused)
    if o.post_extract_proj is not None:
        feats = o.post_extract_proj(feats)
    return o.encoder(feats, pm)
