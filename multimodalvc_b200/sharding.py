"""Batch sharding of inference across the GPUs of one node — one process per GPU, no collective on the data
path (the reference shards its eval iterator the same way: get_batch_iterator(num_shards, shard_id),
src/eval.py:168-169).  Pure host logic; exercised with world_size-2 gloo tests on CPU.

The unit is a clip.  Cost model per clip of T frames: lip frontend + encoder GEMMs are linear in
T, attention is quadratic: flops(T) = T * (FRONTEND + ENCODER_LIN) + T^2 * ATT.
"""
from typing import List, Sequence

FRONTEND_MAC_PER_FRAME = 316_158_976


def clip_cost(T: int, D: int = 1024, F: int = 4096, L: int = 24, audio: bool = True) -> float:
    """Algorithmic MACs of one clip of T frames."""
    lin = 512 * D + (104 * D if audio else 0) + 2 * D * D + D * (D // 16) * 128 + L * (4 * D * D + 2 * D * F)
    return float(T) * (FRONTEND_MAC_PER_FRAME + lin) + float(T) * float(T) * (2 * L * D)


def contiguous_shard(n: int, rank: int, world: int) -> range:
    """Fixed-length batches: contiguous B/G slices; the first n % world ranks get one extra clip."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def balanced_shards(lengths: Sequence[int], world: int, **cost_kw) -> List[List[int]]:
    """Ragged batches: longest-processing-time-first greedy assignment so every rank gets about the same work.
    Deterministic (ties broken by index); returns one index list per rank, each sorted by decreasing length so
    that a rank can run length-bucketed sub-batches with little padding."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += clip_cost(lengths[i], **cost_kw)
    return shards


def length_buckets(indices: Sequence[int], lengths: Sequence[int], max_pad_frac: float = 0.15,
                   max_clips: int = 64) -> List[List[int]]:
    """Split a rank's (length-sorted) clips into sub-batches whose padding waste stays below max_pad_frac."""
    idx = sorted(indices, key=lambda i: (-lengths[i], i))
    out: List[List[int]] = []
    cur: List[int] = []
    for i in idx:
        if cur:
            tmax = lengths[cur[0]]
            tot = sum(lengths[j] for j in cur) + lengths[i]
            if len(cur) >= max_clips or 1.0 - tot / float(tmax * (len(cur) + 1)) > max_pad_frac:
                out.append(cur)
                cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def token_buckets(indices: Sequence[int], lengths: Sequence[int], max_tokens: int = 4800, max_clips: int = 64) -> List[List[int]]:
    """Sub-batches for the PACKED ragged path (cfg.ragged='packed': no work on pad frames, so padding waste is not a
    constraint): clips in decreasing-length order, a new sub-batch whenever the packed frame count would pass
    `max_tokens` (one forward ~ one or two config-2-sized steps keeps the GEMMs at their efficient M)."""
    idx = sorted(indices, key=lambda i: (-lengths[i], i))
    out: List[List[int]] = []
    cur: List[int] = []
    tot = 0
    for i in idx:
        if cur and (tot + lengths[i] > max_tokens or len(cur) >= max_clips):
            out.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += lengths[i]
    if cur:
        out.append(cur)
    return out
