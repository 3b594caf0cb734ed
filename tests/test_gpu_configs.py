"""BASELINE configs 3 (ragged batch, key-padding masks, sharded) and 4 (babble noise at -5/0/5 dB in front of the
log-fbank frontend) at reduced and full sizes: CUDA path vs the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from multimodalvc_b200 import audio, sharding
from oracle import avhubert_oracle as ao
from oracle import fbank_oracle as fo

from helpers import cosine, make_device_model, rel_err, to_dev

pytestmark = pytest.mark.gpu


def test_config3_ragged_batch_sharded_over_two_ranks_matches_oracle():
    """Ragged clips, sorted/bucketed per rank like a 2-GPU run; every clip's valid frames match the oracle run on
    the whole padded batch (fp32 mode gate 2e-3)."""
    oracle = ao.build_oracle("tiny", seed=1234)
    m = make_device_model(oracle, {}, "tiny", torch.float32)
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(5, 41, (12,), generator=g).tolist()
    T = max(lengths)
    src, pm = ao.synthetic_inputs(12, T, lengths=lengths, seed=13)
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, pm)
    shards = sharding.balanced_shards(lengths, 2)
    seen = set()
    for shard in shards:
        for bucket in sharding.length_buckets(shard, lengths, max_pad_frac=0.25, max_clips=4):
            Tb = max(lengths[i] for i in bucket)
            idx = torch.tensor(bucket)
            sub = {"audio": src["audio"][idx][:, :, :Tb], "video": src["video"][idx][:, :, :Tb]}
            sub_pm = pm[idx][:, :Tb]
            d_src, d_pm = to_dev(sub, sub_pm)
            y, pm_out = m.extract_finetune(d_src, d_pm)
            assert torch.equal(pm_out.cpu(), sub_pm)                       # masks bit-exact
            for j, i in enumerate(bucket):
                n = lengths[i]
                assert rel_err(y[j, :n].cpu(), y_ref[i, :n]) < 2e-3, (i, n)
                seen.add(i)
    assert seen == set(range(12))


def test_config3_large_ragged_bf16_valid_frames_independent_of_batching():
    oracle = ao.build_oracle("large", seed=1234)
    m = make_device_model(oracle, {}, "large", torch.bfloat16)
    lengths = [150, 97, 25, 60]
    src, pm = ao.synthetic_inputs(4, 150, lengths=lengths, seed=5)
    d_src, d_pm = to_dev(src, pm, dtype=torch.bfloat16)
    y, _ = m.extract_finetune(d_src, d_pm)
    for i, n in enumerate(lengths):
        one = {"audio": d_src["audio"][i:i + 1, :, :n], "video": d_src["video"][i:i + 1, :, :n]}
        y1, _ = m.extract_finetune(one, None)
        assert cosine(y1.float().cpu(), y[i:i + 1, :n].float().cpu()) > 0.9995


@pytest.mark.parametrize("snr", [-5, 0, 5])
def test_config4_noisy_frontend_matches_oracle(snr):
    """add_noise -> logfbank -> stack -> LN on the device vs the float64 oracle chain; the device mix may differ
    from numpy's by 1 LSB on a few samples (fp32 RMS summation order), so features are compared on the device's
    own mixed waveform (1e-4 gate) and the mix itself at +-1 LSB."""
    clean = [fo.synthetic_wave(n, 30 + i) for i, n in enumerate([38400, 25600, 31000])]
    noise = fo.synthetic_babble(20000, 9)
    mixed = audio.add_noise([torch.from_numpy(c) for c in clean], torch.from_numpy(noise), snr)
    vlen = [60, 40, 48]
    feats, pm = audio.logfbank_stack_collate(mixed, video_lens=vlen)
    torch.cuda.synchronize()
    feats = feats.transpose(1, 2).cpu().numpy()
    for i, c in enumerate(clean):
        ref_mix = fo.add_noise(c, noise, snr)
        dev_mix = mixed[i].cpu().numpy()
        assert np.abs(dev_mix.astype(np.int32) - ref_mix.astype(np.int32)).max() <= 1
        ref = fo.featurize_clip(dev_mix, n_video=vlen[i])
        assert np.abs(feats[i, :vlen[i]] - ref).max() < 1e-4
        assert not feats[i, vlen[i]:].any()
        assert np.array_equal(pm[i].cpu().numpy(), np.arange(60) >= vlen[i])


def test_config4_full_size_noisy_batch_runs_end_to_end():
    """32 x 24 s segments (T = 600): noise mix + fbank + encoder forward in chunks of 8 clips; properties only."""
    oracle = ao.build_oracle("tiny", seed=1234)           # tiny encoder: the audio side is what is at full size
    m = make_device_model(oracle, {}, "tiny", torch.bfloat16)
    clean = [torch.from_numpy(fo.synthetic_wave(384000, 200 + i)) for i in range(8)]
    noise = torch.from_numpy(fo.synthetic_babble(100000, 3))
    mixed = audio.add_noise(clean, noise, 0)
    a, pm = audio.logfbank_stack_collate(mixed, video_lens=[600] * 8)
    assert a.shape == (8, 104, 600) and not pm.any()
    v = torch.randn(8, 1, 600, 88, 88, device="cuda", dtype=torch.bfloat16)
    y, _ = m.extract_finetune({"audio": a.bfloat16(), "video": v}, pm)
    assert y.shape == (8, 600, 128) and torch.isfinite(y).all()
    y2, _ = m.extract_finetune({"audio": a.bfloat16()[:2], "video": v[:2]}, pm[:2])
    assert cosine(y2.float().cpu(), y[:2].float().cpu()) > 0.9999
