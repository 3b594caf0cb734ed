"""Span masking, host side + CPU oracle: masking.compute_mask_indices against golden outputs of
the REAL function (bit-exact, generator position included), live against the reference when present, structural
properties, and oracle/pretrain_oracle.py against goldens of the real AVHubertModel.forward(mask=True)."""
import os

import numpy as np
import pytest
import torch

from multimodalvc_b200 import masking
from oracle import make_golden_pretrain as mg
from oracle import pretrain_oracle as po
from oracle import ref_import

from helpers import load_pretrain_case


def test_compute_mask_indices_matches_the_real_function_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "mask_indices.npz"))
    for i, (seed, B, T, ragged, prob, length, kind, other, mm) in enumerate(mg.MASK_SWEEP):
        pm = mg.sweep_padding(seed, B, T) if ragged else None
        np.random.seed(seed)
        m, s, e, b = masking.compute_mask_indices((B, T), pm, prob, length, kind, other, min_masks=mm)
        assert np.array_equal(m, z[f"m{i}"]) and np.array_equal(s, z[f"s{i}"])
        assert np.array_equal(e, z[f"e{i}"]) and np.array_equal(b, z[f"b{i}"])
        assert np.random.rand() == float(z[f"next{i}"]), "a different number of variates was drawn"


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present on this box")
def test_compute_mask_indices_live_against_reference():
    ru = mg.real_utils()
    for seed in range(12):
        B, T = 3, 40 + 5 * seed
        pm = mg.sweep_padding(seed, B, T) if seed % 2 else None
        for kind, other in (("static", 0), ("uniform", 1), ("normal", 3.0), ("poisson", 0)):
            np.random.seed(seed)
            ref = ru.compute_mask_indices((B, T), pm, 0.5, 6, kind, other, min_masks=2)
            np.random.seed(seed)
            got = masking.compute_mask_indices((B, T), pm, 0.5, 6, kind, other, min_masks=2)
            assert all(np.array_equal(x, y) for x, y in zip(ref, got))


def test_mask_structure_properties():
    np.random.seed(3)
    T = 80
    pm = torch.arange(T)[None, :] >= torch.tensor([80, 61, 47, 30])[:, None]
    m, s, e, b = masking.compute_mask_indices((4, T), pm, 0.65, 10, "static", 0, min_masks=2)
    assert not (m & pm.numpy()).any()                       # padded frames are never masked
    assert len(set(m.sum(1).tolist())) == 1                 # every clip masks the same number of frames
    for i in range(4):                                      # the runs cover exactly the mask
        cover = np.zeros(T, bool)
        for s_, e_ in zip(s[b == i], e[b == i]):
            assert e_ > s_
            cover[s_:e_] = True
        assert np.array_equal(cover, m[i])
    # no_overlap: spans of one clip keep min_space frames apart (the reference's branch dies on np.int under numpy >= 1.24)
    np.random.seed(4)
    m2, s2, e2, b2 = masking.compute_mask_indices((2, 120), None, 0.3, 6, "static", 0, min_masks=2, no_overlap=True, min_space=2)
    assert m2.any()
    assert masking.mask_runs(np.array([0, 1, 1, 0, 1], bool))[0].tolist() == [1, 4]
    assert masking.mask_runs(np.array([], bool))[0].size == 0


def test_span_codes():
    m = np.array([[0, 1, 1, 0], [1, 0, 0, 1]], bool)
    assert masking.span_codes_constant(m, masking.EMB).tolist() == [[-1, -3, -3, -1], [-3, -1, -1, -3]]
    assert masking.span_codes_other_clip(m, [1, 0]).tolist() == [[-1, 5, 6, -1], [0, -1, -1, 3]]
    np.random.seed(0)
    s, e = masking.mask_runs(m[0])
    codes = masking.span_codes_same_clip(m, s, e, np.zeros(len(s), np.int64))
    assert (codes[0, 1:3] >= 0).all() and (codes[0, 1:3] < 4).all() and codes[1].tolist() == [-1] * 4


@pytest.mark.parametrize("name", list(mg.CASES))
def test_pretrain_oracle_reproduces_reference_golden(name):
    c = load_pretrain_case(name)
    z, head = c["z"], c["head"]
    if head.masking_type == "input":
        np.random.seed(7)
        torch.manual_seed(7)
        v_m, mi_v = po.apply_input_mask(head, c["src"]["video"], c["pm"], masking.compute_mask_indices)
        a_m, mi_a = po.apply_input_mask(head, c["src"]["audio"], c["pm"], masking.compute_mask_indices)
        assert np.array_equal(mi_v.numpy(), z["mask_video"]) and np.array_equal(mi_a.numpy(), z["mask_audio"])
        assert np.array_equal(a_m.numpy(), z["audio_masked"])
        assert np.array_equal(v_m.double().sum(dim=(-1, -2)).numpy(), z["video_masked_framesum"])
    np.random.seed(7)
    torch.manual_seed(7)
    res = po.forward(c["oracle"], head, c["src"], c["targets"], c["pm"], masking.compute_mask_indices)
    for i in range(c["n_dicts"]):
        assert np.array_equal(res["target_m_list"][i].numpy(), z[f"target_m{i}"])
        assert np.array_equal(res["target_u_list"][i].numpy(), z[f"target_u{i}"])
        assert np.abs(res["logit_m_list"][i].numpy() - z[f"logit_m{i}"]).max() < 2e-4
        assert np.abs(res["logit_u_list"][i].numpy() - z[f"logit_u{i}"]).max() < 2e-4
    assert abs(res["features_pen"].item() - float(z["features_pen"])) < 1e-4
    np.random.seed(7)
    torch.manual_seed(7)
    fo = po.forward(c["oracle"], head, c["src"], None, c["pm"], masking.compute_mask_indices, features_only=True, output_layer=1)
    assert np.abs(fo["x"].numpy() - z["fo_x"]).max() < 1e-4
    assert np.abs(fo["features"].numpy() - z["fo_features"]).max() < 1e-4
