"""Pins oracle/avhubert_oracle.py: (a) against the golden outputs the REAL reference produced
(tests/golden/enc_*.npz, made by oracle/make_golden.py) — runs anywhere; (b) live against the real reference
modules when /root/reference is present (this container only)."""
import os

import pytest
import torch

from oracle import avhubert_oracle as ao
from oracle import ref_import

from helpers import load_encoder_case, state_checksum

CASES = ["tiny_av_ragged", "tiny_video_only", "tiny_audio_only", "tiny_layer1", "tiny_postln", "tiny_add",
         "base_b1_t50"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name):
    c = load_encoder_case(name)
    assert state_checksum(c["oracle"].state_dict()) == c["checksum"], "seeded weights changed: regenerate goldens"
    with torch.no_grad():
        y, pm = c["oracle"].extract_finetune(c["src"], c["pm"], output_layer=c["output_layer"])
    assert (y - c["y_ref"]).abs().max().item() < 1e-4
    if c["pm"] is not None:
        assert torch.equal(pm, torch.from_numpy(c["pm_ref"]))        # padding mask: bit-exact


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present on this box")
def test_oracle_matches_live_reference_stagewise():
    ref, _ = ref_import.build_reference_model("tiny")
    o = ao.build_oracle("tiny")
    ref.load_state_dict(o.state_dict(), strict=False)
    src, pm = ao.synthetic_inputs(2, 14, lengths=[14, 9], seed=3)
    with torch.no_grad():
        y_ref, _ = ref.extract_finetune(src, pm)
        stages, _ = o.stage_outputs(src, pm)
        res_ref = ref.feature_extractor_video.resnet(src["video"])
    assert (stages["resnet"] - res_ref).abs().max().item() < 1e-5
    assert (stages["x"] - y_ref).abs().max().item() < 1e-4


def test_zero_padded_inputs_make_valid_positions_independent_of_padding():
    # SURVEY.md §7 trap: with collater-style zero padding, valid frames equal the unpadded run
    o = ao.build_oracle("tiny")
    src, pm = ao.synthetic_inputs(1, 12, lengths=[8], seed=5)
    short = {"audio": src["audio"][:, :, :8], "video": src["video"][:, :, :8]}
    with torch.no_grad():
        y_pad, _ = o.extract_finetune(src, pm)
        y_short, _ = o.extract_finetune(short, None)
    assert (y_pad[:, :8] - y_short).abs().max().item() < 1e-4


def test_sr_predictor_oracle_matches_golden_of_the_real_class(golden_dir):
    """oracle/sr_oracle.py vs the output of the REAL Speech_Rate_Predictor (tests/golden/sr_predictor.npz, made by
    oracle/make_golden_sr.py) — and live against the reference class when /root/reference is present."""
    import numpy as np
    import torch
    from oracle import sr_oracle
    z = np.load(os.path.join(golden_dir, "sr_predictor.npz"))
    layers, seed, B, T, xs = [int(v) for v in z["meta"]]
    oracle = sr_oracle.build(layers, seed=seed)
    x = sr_oracle.synthetic_features(B, T, seed=xs)
    with torch.no_grad():
        y = oracle(x)
    assert y.shape == (B, 1) and (y > 0).all()
    assert np.abs(y.numpy() - z["y"]).max() < 1e-5
    from oracle import ref_import
    if ref_import.available():
        from oracle import make_golden_sr
        ref = make_golden_sr.real_class()(layers).eval()
        ref.load_state_dict(oracle.state_dict(), strict=True)
        with torch.no_grad():
            assert (ref(x) - y).abs().max() < 1e-5


@pytest.mark.parametrize("case", ["fgm05_concat", "fgm1_add", "frozen_concat"])
def test_training_graph_gradients_match_the_reference(case):
    """The graph the GPU training tests differentiate (helpers.oracle_finetune_graph, built from the oracle's modules)
    gives the gradients torch.autograd computes for the REAL AVHubertModel.extract_finetune in .train()
    (feature_grad_mult 0.5 / 1 / 0, concat / add fusion, ragged masks): signatures of every parameter gradient from
    oracle/make_golden_grads.py (sum, L2 norm, seeded random projection, first 8 values).  Runs without /root/reference."""
    import ast
    import os
    import zlib
    import numpy as np
    from helpers import oracle_finetune_graph
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_grads_tiny.npz"))
    fgm, fuse, B, T, lengths = [z[f"{case}_meta"][i] for i in range(5)]
    fgm, B, T, lengths = float(fgm), int(B), int(T), ast.literal_eval(str(lengths))
    o = ao.build_oracle("tiny", seed=1234, modality_fuse=str(fuse)).train()
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=41)
    if pm is None:
        pm = torch.zeros(B, T, dtype=torch.bool)
    y = oracle_finetune_graph(o, src, pm, fgm, str(fuse))
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
    ((y * w) * (~pm).unsqueeze(-1)).sum().backward()
    names = [str(n) for n in z[f"{case}_names"]]
    sigs = torch.from_numpy(z[f"{case}_sigs"])
    ref_index = {n: i for i, n in enumerate(names)}
    got = {n: p.grad for n, p in o.named_parameters() if p.grad is not None}
    assert sorted(got) == sorted(names), set(got) ^ set(names)
    for n in names:
        i = ref_index[n]
        g = got[n].detach().double().flatten()
        if n.endswith("k_proj.bias"):          # exactly 0 in exact arithmetic (softmax is shift-invariant): rounding noise only
            q = sigs[ref_index[n.replace("k_proj", "q_proj")]][1].item()
            assert g.norm().item() < 1e-4 * q and sigs[i][1].item() < 1e-4 * q, n
            continue
        r = torch.randn(g.numel(), generator=torch.Generator().manual_seed(zlib.crc32(n.encode())), dtype=torch.float64)
        head = torch.zeros(8, dtype=torch.float64)
        head[:min(8, g.numel())] = g[:8]
        mine = torch.cat([torch.stack([g.sum(), g.norm(), torch.dot(g, r)]), head])
        scale = sigs[i][1].abs().item() + 1e-12                 # the gradient's norm
        # fp32 autograd on both sides, same modules: differences are summation-order noise (BatchNorm backward on ~30 frames)
        assert (mine - sigs[i]).abs().max().item() <= 2e-3 * scale * max(1.0, g.numel() ** 0.5 / 8), (n, mine, sigs[i])

