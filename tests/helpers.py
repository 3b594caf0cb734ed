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
