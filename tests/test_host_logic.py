"""Host-side logic that needs no GPU: frame arithmetic, sharding, and the N>1 partition under gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalvc_b200 import audio, sharding
from oracle import fbank_oracle as fo


@pytest.mark.parametrize("n", [1, 399, 400, 401, 559, 560, 561, 32000, 96000, 384000, 12345])
def test_num_frames_matches_oracle(n):
    assert audio.num_frames(n) == fo.num_frames(n)
    assert audio.stacked_len(n) == len(fo.stacker(__import__("numpy").zeros((fo.num_frames(n), 26)), 4))


def test_contiguous_shard_partitions_exactly():
    for n in (0, 1, 7, 16, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += list(sharding.contiguous_shard(n, r, world))
            assert seen == list(range(n))
            sizes = [len(sharding.contiguous_shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_shards_cover_and_balance():
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(25, 601, (64,), generator=g).tolist()          # BASELINE config 3
    for world in (2, 4, 8):
        shards = sharding.balanced_shards(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(64))
        loads = [sum(sharding.clip_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.10
        for s in shards:
            assert [lengths[i] for i in s] == sorted((lengths[i] for i in s), reverse=True)


def test_length_buckets_bound_padding():
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(25, 601, (64,), generator=g).tolist()
    buckets = sharding.length_buckets(range(64), lengths, max_pad_frac=0.15)
    assert sorted(i for b in buckets for i in b) == list(range(64))
    for b in buckets:
        tmax = max(lengths[i] for i in b)
        assert 1 - sum(lengths[i] for i in b) / (tmax * len(b)) <= 0.15 + 1e-9


def test_clip_cost_matches_baseline_numbers():
    # one 6 s Large clip = 95 495 731 200 MAC
    assert sharding.clip_cost(150) == pytest.approx(95_495_731_200, rel=1e-9)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.balanced_shards(lengths, world)[rank]
    # every rank "processes" its clips: here, a per-clip checksum; no data-path collective is needed.
    part = torch.zeros(len(lengths), dtype=torch.int64)
    for i in mine:
        part[i] = lengths[i] * 31 + 7
    # bench.py-style bookkeeping: units processed and max-over-ranks time
    units = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(units)
    dist.all_reduce(part)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((units.item(), part.tolist(), t.item()))
    dist.destroy_process_group()


def test_two_rank_sharded_run_under_gloo():
    lengths = [150, 30, 600, 75, 320, 25, 410, 90]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    units, part, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert units == len(lengths)
    assert part == [n * 31 + 7 for n in lengths]          # every clip processed exactly once
    assert tmax == 2.0


def test_token_buckets_partition_and_respect_the_budget():
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(25, 601, (64,), generator=g).tolist()          # BASELINE config 3
    for world in (1, 2, 8):
        for shard in sharding.balanced_shards(lengths, world):
            buckets = sharding.token_buckets(shard, lengths, max_tokens=4800, max_clips=32)
            assert sorted(i for b in buckets for i in b) == sorted(shard)
            for b in buckets:
                assert len(b) <= 32
                assert sum(lengths[i] for i in b) <= 4800 or len(b) == 1
                assert [lengths[i] for i in b] == sorted((lengths[i] for i in b), reverse=True)
    assert sharding.token_buckets([], lengths) == []
    assert sharding.token_buckets([3], [10, 10, 10, 9000]) == [[3]]          # a clip longer than the budget runs alone


@pytest.mark.parametrize("fuse", ["concat", "add"])
def test_full_parameters_spec_maps_the_flat_gradient_onto_every_parameter(fuse):
    """AVHubertModel.full_parameters: the (buffer shape, map) pairs that cut the library's flat gradient buffer
    (include/avh_b200.h, avh_full_train_forward) must cover every trainable parameter once, and each map must land on
    the parameter's own shape with the element it names — checked with index-valued buffers, no device needed."""
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    m = AVHubertModel(AVHubertConfig.named("tiny", modality_fuse=fuse, trainable=True))
    m.remove_pretraining_modules()
    params, spec = m.full_parameters(True, True)
    assert len(params) == len(spec)
    named = {id(p): n for n, p in m.named_parameters()}
    seen = set()
    for p, (shape, to_param) in zip(params, spec):
        assert id(p) in named and id(p) not in seen, "unknown or repeated parameter"
        seen.add(id(p))
        n = 1
        for d in shape:
            n *= d
        buf = torch.arange(n, dtype=torch.float64).view(shape)
        g = to_param(buf)
        assert tuple(g.shape) == tuple(p.shape), (named[id(p)], tuple(g.shape), tuple(p.shape))
        assert g.numel() <= n and torch.unique(g).numel() == g.numel()          # a selection / permutation, nothing repeated
    missing = [n for n, p in m.named_parameters() if id(p) not in seen and n != "mask_emb"]
    assert not missing, missing
    # spot checks of the layouts the header documents
    name_of = {named[id(p)]: (p, s) for p, s in zip(params, spec)}
    p, (shape, f) = name_of["feature_extractor_video.resnet.trunk.layer2.0.conv1.weight"]
    assert shape == (128, 3, 3, 64)                                              # [C, kh, kw, Cin] in the buffer
    buf = torch.arange(128 * 9 * 64, dtype=torch.float64).view(shape)
    assert f(buf)[5, 7, 2, 1].item() == buf[5, 2, 1, 7].item()                   # -> [C, Cin, kh, kw]
    p, (shape, f) = name_of["feature_extractor_video.resnet.frontend3D.0.weight"]
    assert shape == (64, 5, 8, 8)                                                # (dt, kh, kw) with the 8th row / column zero
    buf = torch.arange(64 * 320, dtype=torch.float64).view(shape)
    assert f(buf)[3, 0, 4, 6, 5].item() == buf[3, 4, 6, 5].item()
    p, (shape, f) = name_of["feature_extractor_audio.proj.weight"]
    assert shape == (128, 128) and tuple(f(torch.zeros(shape)).shape) == (128, 104)   # K padded to 64 in the buffer
