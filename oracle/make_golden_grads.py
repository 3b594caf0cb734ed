"""TEST INFRASTRUCTURE ONLY (this container; needs /root/reference) — gradient signatures of the REAL AVHubertModel in
.train() with feature_grad_mult > 0 (avhubert/hubert.py:538-547: the feature extractors differentiate, their output
gradient scaled by GradMultiply) and with feature_grad_mult = 0 (extractors under no_grad), through the real
extract_finetune (hubert.py:694-745) and torch.autograd.  The GPU tests of the training step (tests/test_gpu_backward.py)
compare the device gradients with autograd on a graph REBUILT from the oracle's modules; this fixture pins that rebuilt
graph to the real model's own (tests/test_oracle_vs_reference.py::test_training_graph_gradients_match_the_reference).

A full set of gradients is 45 MB, so every parameter is stored as a signature: (sum, L2 norm, dot with a seeded random
vector, first 8 values).

  python -m oracle.make_golden_grads     ->  tests/golden/train_grads_tiny.npz
"""
import os
import zlib

import numpy as np
import torch

from . import avhubert_oracle as ao
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, feature_grad_mult, modality_fuse, B, T, lengths
    ("fgm05_concat", 0.5, "concat", 2, 14, [14, 9]),
    ("fgm1_add", 1.0, "add", 2, 11, None),
    ("frozen_concat", 0.0, "concat", 2, 12, [12, 7]),
]


def signature(g, seed):
    g = g.detach().double().flatten()
    r = torch.randn(g.numel(), generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    head = torch.zeros(8, dtype=torch.float64)
    head[:min(8, g.numel())] = g[:8]
    return torch.cat([torch.stack([g.sum(), g.norm(), torch.dot(g, r)]), head])


def loss_of(y, pm, seed):
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32).to(y.dtype)
    keep = torch.ones(y.shape[:2], dtype=torch.bool) if pm is None else ~pm
    return ((y * w) * keep.unsqueeze(-1)).sum()


def main():
    out = {}
    for name, fgm, fuse, B, T, lengths in CASES:
        oracle = ao.build_oracle("tiny", seed=1234, modality_fuse=fuse)
        ref, _ = ref_import.build_reference_model("tiny", modality_fuse=fuse, feature_grad_mult=fgm)
        assert not ref.load_state_dict(oracle.state_dict(), strict=False).unexpected_keys
        src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=41)
        if pm is None:
            pm = torch.zeros(B, T, dtype=torch.bool)       # the reference dereferences the mask (hubert.py:723)
        ref.train()
        ref.zero_grad()
        # the reference's frontend mutates the caller's video in place (transpose + contiguous alias): hand it a copy
        y, _ = ref.extract_finetune({k: v.clone() for k, v in src.items()}, pm)
        loss_of(y, pm, 77).backward()
        names = []
        sigs = []
        for n, p in ref.named_parameters():
            if p.grad is None:
                continue
            names.append(n)
            sigs.append(signature(p.grad, zlib.crc32(n.encode())))       # projection seeded by the parameter's name
        print(f"{name}: {len(names)} parameters with gradients, |y| {y.abs().mean().item():.4f}")
        out[f"{name}_names"] = np.array(names)
        out[f"{name}_sigs"] = torch.stack(sigs).numpy()
        out[f"{name}_meta"] = np.array([repr(fgm), fuse, str(B), str(T), repr(lengths)])
    np.savez_compressed(os.path.join(OUT, "train_grads_tiny.npz"), **out)


if __name__ == "__main__":
    main()
