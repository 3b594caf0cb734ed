"""Known-answer tests pinning oracle/fbank_oracle.py: python_speech_features is absent, so the
log-fbank restatement is pinned by structure + analytic answers; stacker / alignment / add_noise / collater
are pinned against outputs of the REAL avhubert/hubert_dataset.py stored in tests/golden/audio_reference.npz."""
import os

import numpy as np
import pytest

from oracle import fbank_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n,frames", [(1, 1), (400, 1), (401, 2), (560, 2), (561, 3), (32000, 199), (96000, 599),
                                      (384000, 2399)])
def test_frame_count_table(n, frames):
    assert fo.num_frames(n) == frames
    assert fo.framesig(np.zeros(n)).shape == (frames, 400)


@pytest.mark.parametrize("frames,stacked", [(199, 50), (599, 150), (2399, 600), (4, 1), (5, 2)])
def test_stacker_shapes(frames, stacked):
    x = np.arange(frames * 26, dtype=np.float32).reshape(frames, 26)
    y = fo.stacker(x, 4)
    assert y.shape == (stacked, 104)
    assert np.array_equal(y.reshape(-1)[:frames * 26], x.reshape(-1))      # bit-exact copy
    assert not y.reshape(-1)[frames * 26:].any()                           # zero padding


def test_filterbank_structure():
    fb = fo.get_filterbanks()
    assert fb.shape == (26, 257)
    bins = fo.mel_bins()
    assert bins[0] == 0 and bins[-1] == 256 and np.all(np.diff(bins) > 0)
    for j in range(26):
        nz = np.nonzero(fb[j])[0]
        assert nz.min() >= bins[j] and nz.max() < bins[j + 2]
        assert fb[j, int(bins[j + 1])] == 1.0                              # triangle peak
    assert fb.min() >= 0.0 and fb.max() == 1.0
    z = np.load(os.path.join(GOLDEN, "audio_fbank.npz"))
    assert np.array_equal(z["filterbank"], fb)


def test_preemphasis_and_padding():
    x = np.array([3, 5, -2, 7], dtype=np.int16)
    y = fo.preemphasis(x)
    assert np.allclose(y, [3, 5 - 0.97 * 3, -2 - 0.97 * 5, 7 + 0.97 * 2])
    fr = fo.framesig(y)
    assert fr.shape == (1, 400) and not fr[0, 4:].any()


def test_pure_tone_peaks_in_the_right_filter():
    t = np.arange(16000)
    for hz in (500.0, 1000.0, 3000.0):
        w = np.round(8000 * np.sin(2 * np.pi * hz * t / 16000)).astype(np.int16)
        f = fo.logfbank(w)
        bins = fo.mel_bins()
        k = hz * 512 / 16000
        j = int(np.argmax(f[5]))
        assert bins[j] <= k <= bins[j + 2]


def test_parseval_energy_identity():
    # sum of the (un-normalised) power spectrum equals frame energy: pins powspec scaling 1/NFFT
    r = np.random.RandomState(0)
    x = r.randn(1, 400)
    p = fo.powspec(x)
    full = 2 * p.sum() - p[0, 0] - p[0, -1]
    assert np.isclose(full, (x ** 2).sum())


def test_silence_maps_to_log_eps():
    f = fo.logfbank(np.zeros(2000, dtype=np.int16))
    assert np.all(f == np.log(np.finfo(float).eps))


def test_golden_fbank_reproducible():
    z = np.load(os.path.join(GOLDEN, "audio_fbank.npz"))
    for k in z.files:
        if k.startswith("wav_"):
            name = k[4:]
            assert np.array_equal(fo.logfbank(z[k]), z["fbank_" + name])
            assert np.array_equal(fo.featurize_clip(z[k]), z["feat_" + name])


def test_collater_against_real_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "audio_reference.npz"))
    lens = z["coll_lens"]
    items, o = [], 0
    for n in lens:
        items.append(z["coll_items"][o:o + n])
        o += n
    out, mask = fo.collater_audio(items, 50)
    assert np.array_equal(mask, z["coll_mask"])                          # bit-exact mask
    assert np.array_equal(out.transpose(0, 2, 1), z["coll_out"])         # reference returns [B,F,T]
    assert z["collv_out"].shape == (3, 1, 50, 4, 4) and np.array_equal(z["collv_mask"], z["coll_mask"])


def test_add_noise_against_real_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "audio_reference.npz"))
    for snr, key in [(-5, "mix_snr_m5"), (0, "mix_snr_0"), (5, "mix_snr_5"), (40, "mix_snr_40")]:
        assert np.array_equal(fo.add_noise(z["clean"], z["noise"], snr), z[key])     # same numpy ops => exact
    assert np.array_equal(fo.add_noise(z["loud"], z["noise"], -5), z["mix_loud"])
    assert np.abs(z["mix_loud"].astype(np.int32)).max() >= 32760                     # clipping branch was taken


def test_alignment():
    a = np.ones((10, 104), dtype=np.float32)
    assert fo.align_to_video(a, 12).shape == (12, 104) and not fo.align_to_video(a, 12)[10:].any()
    assert fo.align_to_video(a, 7).shape == (7, 104)
    assert fo.align_to_video(a, 10) is a
