"""CPU: the video pre-processing restatement (oracle/video_oracle.py) against outputs of the REAL
avhubert/utils.py transform classes and the real collater (tests/golden/video_reference.npz) — bit-exact."""
import os

import numpy as np

from oracle import video_oracle as vo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_transform_matches_reference_bit_for_bit():
    z = np.load(os.path.join(GOLDEN, "video_reference.npz"))
    for name in ("roi96", "odd_97x101", "exact88"):
        y = vo.video_transform(z["frames_" + name])
        assert y.dtype == np.float64 and y.shape[1:] == (88, 88)
        assert np.array_equal(y, z["out_" + name]), name


def test_center_crop_offsets_follow_the_reference_rounding():
    # delta = int(round(w - tw) / 2.): truncation of the halved difference (utils.py:86-88)
    f = np.arange(3 * 97 * 101, dtype=np.float64).reshape(3, 97, 101)
    c = vo.center_crop(f, (88, 88))
    assert np.array_equal(c, f[:, 4:92, 6:94])
    assert vo.center_crop(f[:, :88, :88], (88, 88)).shape == (3, 88, 88)


def test_collater_layout_and_mask():
    z = np.load(os.path.join(GOLDEN, "video_reference.npz"))
    lens = z["coll_lens"]
    clips, o = [], 0
    for n in lens:
        clips.append(z["coll_frames"][o:o + n])
        o += n
    items = [vo.load_video_feats(c) for c in clips]
    out, mask = vo.collater_video(items, int(lens.max()))
    assert out.shape == (3, 1, 9, 88, 88) and out.dtype == np.float32
    assert np.array_equal(out, z["coll_out"])
    assert np.array_equal(mask, z["coll_mask"])
    assert not out[0, 0, 5:].any() and not out[2, 0, 2:].any()          # zero-padded tail frames


def test_value_range_and_known_answers():
    u = np.array([[[0, 255]] * 88] * 1, dtype=np.uint8).reshape(1, 88, 2).repeat(44, axis=2)
    y = vo.video_transform(u)
    assert np.isclose(y.min(), (0.0 - 0.421) / 0.165) and np.isclose(y.max(), (1.0 - 0.421) / 0.165)
