"""Oracle-compared parity at the sizes BASELINE.json's configs name (VERDICT r1 "parity holes"): config 2 at full
size, the Large shape beyond 160 frames (the streaming attention tile, positional-conv halos, multi-chunk lip
frontend) in bf16 AND fp32 mode with ragged lengths + key-padding masks, config 3 literally (64 clips of 25..600
frames through balanced_shards + length_buckets) and config 4 with the Large encoder on noisy 24 s clips.

Every comparison is CUDA path vs the fp32 CPU oracle (oracle/avhubert_oracle.py, pinned to the real reference by
tests/test_oracle_vs_reference.py and the goldens) on the same seeded inputs.  Gates (BASELINE.json north_star):
fp32 mode max|y - y_ref| / max|y_ref| <= 2e-3, bf16 mode cosine >= 0.999, padding masks bit-exact."""
import numpy as np
import pytest
import torch

from multimodalvc_b200 import audio, sharding
from oracle import avhubert_oracle as ao
from oracle import fbank_oracle as fo

from helpers import cosine, make_device_model, rel_err, to_dev

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-3
BF16_COS = 0.999


@pytest.fixture(scope="module")
def large():
    oracle = ao.build_oracle("large", seed=1234)
    models = {}

    def get(dtype):
        if dtype not in models:
            models[dtype] = make_device_model(oracle, {}, "large", dtype)
        return models[dtype]
    return oracle, get


def test_config2_full_size_every_clip_vs_oracle(large):
    """BASELINE config 2 as benchmarked: Large, 16 x 150 frames, audio + video, bf16: cosine per clip."""
    oracle, get = large
    src, _ = ao.synthetic_inputs(16, 150, seed=21)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, None)
    m = get(torch.bfloat16)
    d_src, _ = to_dev(src, None, dtype=torch.bfloat16)
    y, pm = m.extract_finetune(d_src, None)
    assert pm is None and y.shape == (16, 150, 1024) and torch.isfinite(y).all()
    y = y.float().cpu()
    cos = [cosine(y[i], y_ref[i]) for i in range(16)]
    assert min(cos) > BF16_COS, cos
    # replay (CUDA graph) gives the same bits
    y2, _ = m.extract_finetune(d_src, None)
    y3, _ = m.extract_finetune(d_src, None)
    assert torch.equal(y2.float().cpu(), y) and torch.equal(y3.float().cpu(), y)


def test_config2_full_size_fp32_mode_vs_oracle(large):
    oracle, get = large
    src, _ = ao.synthetic_inputs(16, 150, seed=22)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, None)
    d_src, _ = to_dev(src, None)
    y, _ = get(torch.float32).extract_finetune(d_src, None)
    assert rel_err(y.cpu(), y_ref) < FP32_TOL


@pytest.mark.parametrize("T,short", [(161, 113), (300, 209), (600, 417)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_large_long_ragged_clips_vs_oracle(large, T, short, dtype):
    """T > 160 takes the streaming attention tile (attention.cu) and, at B*T > 2400, more than one lip-frontend
    chunk (api.cu b0 > 0); B=3 so that T=300/600 also span chunk boundaries.  Dense mode: pad positions compared."""
    oracle, get = large
    lengths = [T, short, max(1, T // 3)]
    src, pm = ao.synthetic_inputs(3, T, lengths=lengths, seed=T)
    with torch.no_grad():
        y_ref, pm_ref = oracle.extract_finetune(src, pm)
    d_src, d_pm = to_dev(src, pm, dtype=dtype)
    y, pm_out = get(dtype).extract_finetune(d_src, d_pm)
    assert torch.equal(pm_out.cpu(), pm_ref)
    y = y.float().cpu()
    assert torch.isfinite(y).all()
    if dtype == torch.float32:
        assert rel_err(y, y_ref) < FP32_TOL
        for i, n in enumerate(lengths):
            assert rel_err(y[i, :n], y_ref[i, :n]) < FP32_TOL, (i, n)
    else:
        assert cosine(y, y_ref) > BF16_COS
        for i, n in enumerate(lengths):
            assert cosine(y[i, :n], y_ref[i, :n]) > BF16_COS, (i, n)


def _clip(i, n):
    """Clip i of the config-3 batch: its own seeded video / audio features of n frames."""
    src, _ = ao.synthetic_inputs(1, n, seed=3000 + i)
    return src["video"][0], src["audio"][0]          # [1,n,88,88], [104,n]


@pytest.mark.parametrize("world", [2, 8])
def test_config3_literal_ragged_batch_sharded_vs_oracle(large, world):
    """BASELINE config 3: Large, 64 clips with lengths U{25..600} (seed 7), key-padding masks, clips dealt to `world`
    ranks by balanced_shards and run as length_buckets sub-batches; the valid frames of 8 sampled clips (incl. the
    shortest and the longest) are compared with the oracle run on each clip alone (bf16 gate)."""
    oracle, get = large
    m = get(torch.bfloat16)
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(25, 601, (64,), generator=g).tolist()
    shards = sharding.balanced_shards(lengths, world)
    assert sorted(i for s in shards for i in s) == list(range(64))
    order = sorted(range(64), key=lambda i: lengths[i])
    sample = {order[0], order[-1], *order[5:60:10]}
    assert len(sample) == 8
    got = {}
    ranks = range(world) if world == 2 else [r for r in range(world) if sample & set(shards[r])]
    for r in ranks:
        for bucket in sharding.length_buckets(shards[r], lengths, max_pad_frac=0.15, max_clips=8):
            Tb = max(lengths[i] for i in bucket)
            v = torch.zeros(len(bucket), 1, Tb, 88, 88)
            a = torch.zeros(len(bucket), 104, Tb)
            pm = torch.ones(len(bucket), Tb, dtype=torch.bool)
            for j, i in enumerate(bucket):
                cv, ca = _clip(i, lengths[i])
                v[j, :, :lengths[i]] = cv
                a[j, :, :lengths[i]] = ca
                pm[j, :lengths[i]] = False
            y, pm_out = m.extract_finetune({"audio": a.cuda().bfloat16(), "video": v.cuda().bfloat16()}, pm.cuda())
            assert torch.equal(pm_out.cpu(), pm)
            assert torch.isfinite(y).all()
            for j, i in enumerate(bucket):
                if i in sample:
                    got[i] = y[j, :lengths[i]].float().cpu()
    assert set(got) == sample
    for i in sorted(sample):
        cv, ca = _clip(i, lengths[i])
        with torch.no_grad():
            y_ref, _ = oracle.extract_finetune({"audio": ca[None], "video": cv[None]}, None)
        assert cosine(got[i], y_ref[0]) > BF16_COS, (i, lengths[i])


@pytest.mark.parametrize("snr", [-5, 0, 5])
def test_config4_noisy_24s_segments_large_encoder_vs_oracle(large, snr):
    """BASELINE config 4 on the Large encoder: babble noise mixed at `snr` dB (device) -> log-fbank/stack/LN
    (device) -> AV-HuBERT Large on 24 s segments (T = 600), against the oracle chain fed the same mixed waveform
    (the int16 mix itself is gated at +-1 LSB in test_gpu_configs)."""
    oracle, get = large
    m = get(torch.bfloat16)
    B, T = 2, 600
    clean = [fo.synthetic_wave(T * 640, 400 + 10 * (snr + 5) + i) for i in range(B)]
    noise = fo.synthetic_babble(100000, 3)
    mixed = audio.add_noise([torch.from_numpy(c) for c in clean], torch.from_numpy(noise), snr)
    a_dev, pm = audio.logfbank_stack_collate(mixed, video_lens=[T] * B)
    assert a_dev.shape == (B, 104, T) and not pm.any()
    a_ref = np.stack([fo.featurize_clip(mx.cpu().numpy(), n_video=T) for mx in mixed])        # [B,T,104]
    assert np.abs(a_dev.transpose(1, 2).cpu().numpy() - a_ref).max() < 1e-4
    src, _ = ao.synthetic_inputs(B, T, seed=90 + snr, audio=False)
    ref_src = {"audio": torch.from_numpy(a_ref).float().transpose(1, 2), "video": src["video"]}
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(ref_src, None)
    y, _ = m.extract_finetune({"audio": a_dev.bfloat16(), "video": src["video"].cuda().bfloat16()}, pm)
    y = y.float().cpu()
    for i in range(B):
        assert cosine(y[i], y_ref[i]) > BF16_COS, (snr, i)


def test_raw_uint8_video_ragged_pad_frames_are_zero_in_normalised_space():
    """the reference collater zero-pads AFTER the per-sample Normalize (hubert_dataset.py:430-456), so pad
    frames are 0.0 in normalised space whatever bytes sit in the uint8 buffer.  Oracle input = normalise-then-zero-pad."""
    from oracle import video_oracle as vo
    oracle = ao.build_oracle("tiny", seed=1234)
    m = make_device_model(oracle, {}, "tiny", torch.float32)
    lengths = [12, 7]
    clips = [vo.synthetic_frames(n, 96, 96, seed=50 + i) for i, n in enumerate(lengths)]
    v_ref, _ = vo.collater_video([vo.load_video_feats(c) for c in clips], 12)               # [B,1,T,88,88] fp32
    raw = torch.full((2, 1, 12, 96, 96), 173, dtype=torch.uint8)                              # garbage in the pad frames
    for i, c in enumerate(clips):
        raw[i, 0, :lengths[i]] = torch.from_numpy(np.asarray(c)).view(lengths[i], 96, 96)
    src, pm = ao.synthetic_inputs(2, 12, lengths=lengths, seed=3, video=False)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune({"audio": src["audio"], "video": torch.as_tensor(v_ref)}, pm)
    y, _ = m.extract_finetune({"audio": src["audio"].cuda(), "video": raw.cuda()}, pm.cuda())
    assert rel_err(y.cpu(), y_ref) < FP32_TOL
    host = m.extract_finetune_host(raw, src["audio"], pm)
    assert rel_err(host, y_ref) < FP32_TOL


def test_two_devices_in_one_process():
    """dynamic-smem opt-in is per (device, function); a second handle on cuda:1 must work after cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    oracle = ao.build_oracle("tiny", seed=1234)
    src, pm = ao.synthetic_inputs(2, 20, lengths=[20, 14], seed=11)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, pm)
    for dev in ("cuda:0", "cuda:1"):
        for dtype in (torch.float32, torch.bfloat16):
            m = make_device_model(oracle, {}, "tiny", dtype, device=dev)
            d_src, d_pm = to_dev(src, pm, device=dev, dtype=dtype)
            y, _ = m.extract_finetune(d_src, d_pm)
            assert cosine(y.float().cpu(), y_ref) > BF16_COS


def test_packed_ragged_mode_equals_dense_mode_on_valid_frames_tiny():
    """cfg.ragged='packed' (no work on pad frames): valid positions carry the dense path's values (same kernels, same
    per-row arithmetic; GEMM rows are independent), pad positions are zeros; also vs the oracle."""
    oracle = ao.build_oracle("tiny", seed=1234)
    lengths = [40, 17, 1, 33, 40, 8]
    src, pm = ao.synthetic_inputs(6, 40, lengths=lengths, seed=31)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, pm)
    dense = make_device_model(oracle, {}, "tiny", torch.bfloat16)
    packed = make_device_model(oracle, {}, "tiny", torch.bfloat16, ragged="packed")
    d_src, d_pm = to_dev(src, pm, dtype=torch.bfloat16)
    y_d, _ = dense.extract_finetune(d_src, d_pm)
    for trial in range(3):                                  # 2nd call captures the CUDA graph, 3rd replays it
        y_p, pm_out = packed.extract_finetune(d_src, d_pm)
        assert torch.equal(pm_out, d_pm)
        for i, n in enumerate(lengths):
            assert cosine(y_p[i, :n].float().cpu(), y_d[i, :n].float().cpu()) > 0.9999, (trial, i)
            assert cosine(y_p[i, :n].float().cpu(), y_ref[i, :n]) > BF16_COS, (trial, i)
            assert not y_p[i, n:].any()
    # explicit lengths (no mask read-back), video only / audio only, early exit
    y_l, _ = packed.extract_finetune(d_src, d_pm, lengths=lengths)
    assert torch.equal(y_l, y_p)
    for drop in ("audio", "video"):
        one = dict(d_src)
        one[drop] = None
        y1, _ = packed.extract_finetune(one, d_pm, output_layer=1)
        y2, _ = dense.extract_finetune(one, d_pm, output_layer=1)
        for i, n in enumerate(lengths):
            assert cosine(y1[i, :n].float().cpu(), y2[i, :n].float().cpu()) > 0.9999, (drop, i)


def test_packed_ragged_mode_large_long_clips_vs_oracle(large):
    """Large, ragged 25..600-frame clips in packed mode (streamed attention per clip through cu_rows, ragged stem work
    list, positional conv through the row map) against the oracle on each clip alone."""
    oracle, _ = large
    m = make_device_model(oracle, {}, "large", torch.bfloat16, ragged="packed")
    lengths = [600, 25, 161, 310, 97]
    T = max(lengths)
    v = torch.zeros(len(lengths), 1, T, 88, 88)
    a = torch.zeros(len(lengths), 104, T)
    pm = torch.ones(len(lengths), T, dtype=torch.bool)
    for j, n in enumerate(lengths):
        cv, ca = _clip(100 + j, n)
        v[j, :, :n] = cv
        a[j, :, :n] = ca
        pm[j, :n] = False
    y, _ = m.extract_finetune({"audio": a.cuda().bfloat16(), "video": v.cuda().bfloat16()}, pm.cuda())
    y2, _ = m.extract_finetune({"audio": a.cuda().bfloat16(), "video": v.cuda().bfloat16()}, pm.cuda())
    assert torch.equal(y, y2)
    for j, n in enumerate(lengths):
        cv, ca = _clip(100 + j, n)
        with torch.no_grad():
            y_ref, _ = oracle.extract_finetune({"audio": ca[None], "video": cv[None]}, None)
        assert cosine(y[j, :n].float().cpu(), y_ref[0]) > BF16_COS, (j, n)
        assert not y[j, n:].any()
