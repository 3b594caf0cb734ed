"""TEST INFRASTRUCTURE ONLY (this container; needs /root/reference) — golden vectors of the TRAINING-MODE forward of the
REAL AVHubertModel (module in .train(), torch.no_grad(): what MMS-LLaMA's frozen encoder runs during training,
src/model.py:96-100,280): BatchNorm batch statistics + running-stat updates (avhubert/resnet.py), LayerDrop coins from
np.random.random() (wav2vec2.py:886-888); every dropout probability 0 so that the output is deterministic.

  python -m oracle.make_golden_train     ->  tests/golden/enc_tiny_train.npz, enc_tiny_train_layerdrop.npz
"""
import os

import numpy as np
import torch

from . import avhubert_oracle as ao
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, overrides, B, T, lengths, numpy seed for the LayerDrop coins
    ("tiny_train", {}, 3, 12, [12, 9, 5], 0),
    ("tiny_train_layerdrop", {"encoder_layers": 4, "encoder_layerdrop": 0.5}, 2, 10, [10, 6], 3),
]


def bn_modules(model):
    res = model.feature_extractor_video.resnet
    out = [res.frontend3D[1]]
    for i in range(1, 5):
        for blk in getattr(res.trunk, f"layer{i}"):
            out += [blk.bn1, blk.bn2]
            if blk.downsample is not None:
                out.append(blk.downsample[1])
    return out


def bn_flat(model):
    return torch.cat([torch.cat([m.running_mean.float(), m.running_var.float()]) for m in bn_modules(model)])


def main():
    for name, over, B, T, lengths, npseed in CASES:
        o_over = {k: v for k, v in over.items() if k != "encoder_layerdrop"}
        oracle = ao.build_oracle("tiny", seed=1234, **o_over)
        ref, _ = ref_import.build_reference_model("tiny", **over)
        assert not ref.load_state_dict(oracle.state_dict(), strict=False).unexpected_keys
        src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=17)
        ref.train()
        np.random.seed(npseed)
        with torch.no_grad():
            y, _ = ref.extract_finetune(src, pm)
        # the coins the reference drew, re-drawn from the same seed (one np.random.random() per layer, in order)
        np.random.seed(npseed)
        p = over.get("encoder_layerdrop", 0.0)
        skip = [0 if np.random.random() > p else 1 for _ in range(len(ref.encoder.layers))]
        oracle.train()
        oracle.encoder.layer_skip = skip
        with torch.no_grad():
            y_o, _ = oracle.extract_finetune(src, pm)
        err = (y - y_o).abs().max().item()
        serr = (bn_flat(ref) - bn_flat(oracle)).abs().max().item()
        print(f"{name}: skip {skip}  ref vs oracle |dy| {err:.2e}  |d running stats| {serr:.2e}")
        assert err < 2e-4 and serr < 1e-5
        np.savez_compressed(os.path.join(OUT, f"enc_{name}.npz"), y=y.numpy().astype(np.float32),
                            bn=bn_flat(ref).numpy(), skip=np.array(skip, dtype=np.uint8),
                            meta=np.array([repr(over), str(B), str(T), repr(lengths), str(npseed)]))


if __name__ == "__main__":
    main()
