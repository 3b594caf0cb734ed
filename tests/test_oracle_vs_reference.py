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
