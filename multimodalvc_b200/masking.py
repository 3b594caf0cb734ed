"""Host side of AV-HuBERT's span masking: where the spans fall.

``compute_mask_indices`` plays the role of ``avhubert/utils.py:142-270`` (the AV-HuBERT variant of fairseq's function:
it also returns the runs of every clip).  The reference draws from numpy's GLOBAL generator; this version draws the
same variates in the same order — one ``rand()`` for the batch-wide span count, one per clip when a padding mask is
given, the span lengths of the chosen distribution, one ``choice`` of span starts without replacement, and one more
``choice`` per clip that has to be thinned to the batch minimum — so ``np.random.seed(s)`` gives the masks the
reference gives.  The tensor work (substituting frames / embeddings at the masked positions) is a device kernel
behind ``avh_mask_substitute``; ``span_codes_*`` below build its per-frame source codes.
"""
import numpy as np

KEEP, ZERO, EMB = -1, -2, -3          # codes of avh_mask_substitute (include/avh_b200.h)


def _span_lengths(kind, count, mask_length, mask_other):
    if kind == "static":
        return [int(mask_length)] * count
    if kind == "uniform":
        return [int(v) for v in np.random.randint(mask_other, mask_length * 2 + 1, size=count)]
    if kind == "normal":
        return [max(1, int(round(v))) for v in np.random.normal(mask_length, mask_other, size=count)]
    if kind == "poisson":
        return [int(round(v)) for v in np.random.poisson(mask_length, size=count)]
    raise Exception("unknown mask selection " + str(kind))


def _place_without_overlap(lengths, n, min_space):
    """Longest span first into a free interval picked with probability proportional to its size; what is left of the
    interval on either side (minus ``min_space``) stays available if the shortest span still fits."""
    taken = []
    free = [(0, n)]
    shortest = min(lengths)
    for length in sorted(lengths, reverse=True):
        room = np.array([hi - lo if hi - lo >= length + min_space else 0 for lo, hi in free], dtype=np.int64)
        total = int(room.sum())
        if total == 0:
            break
        pick = np.random.choice(len(free), p=room / total)
        lo, hi = free.pop(pick)
        start = np.random.randint(lo, hi - length)
        taken.extend(range(start, start + length))
        if start - lo - min_space >= shortest:
            free.append((lo, start - min_space + 1))
        if hi - start - shortest - min_space > shortest:
            free.append((start + length + min_space, hi))
    return np.asarray(taken, dtype=np.int64)


def mask_runs(row):
    """(starts, ends) of the runs of True in a 1-D bool array."""
    row = np.asarray(row, dtype=bool)
    if row.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    edge = np.diff(np.concatenate(([0], row.astype(np.int8), [0])))
    return np.nonzero(edge == 1)[0].astype(np.int64), np.nonzero(edge == -1)[0].astype(np.int64)


def compute_mask_indices(shape, padding_mask, mask_prob, mask_length, mask_type="static", mask_other=0.0, min_masks=0,
                         no_overlap=False, min_space=0):
    """Returns (mask bool [B,T], starts, ends, batch_indexes) — the runs of True of every clip in clip order, as the
    reference's variant does.  ``padding_mask``: None, a numpy array or a (CPU / CUDA) bool tensor [B,T]."""
    B, T = shape
    mask = np.zeros((B, T), dtype=bool)
    if padding_mask is not None and hasattr(padding_mask, "detach"):
        padding_mask = padding_mask.detach().to("cpu").numpy()

    def span_count(n):          # probabilistic rounding: one uniform variate per call
        return max(min_masks, int(mask_prob * n / float(mask_length) + np.random.rand()))

    batch_wide = span_count(T)
    chosen = []
    for b in range(B):
        if padding_mask is not None:
            n = T - int(np.asarray(padding_mask[b]).astype(np.int64).sum())
            count = span_count(n)
        else:
            n, count = T, batch_wide
        lengths = _span_lengths(mask_type, count, mask_length, mask_other)
        if sum(lengths) == 0:
            lengths[0] = min(mask_length, n - 1)
        if no_overlap:
            idc = _place_without_overlap(lengths, n, min_space)
        else:
            shortest = min(lengths)
            if n - shortest <= count:
                shortest = n - count - 1
            starts = np.random.choice(n - shortest, count, replace=False)
            idc = np.asarray([starts[j] + k for j in range(count) for k in range(lengths[j])], dtype=np.int64)
        chosen.append(np.unique(idc[idc < n]))
    fewest = min(len(c) for c in chosen)
    starts, ends, owners = [], [], []
    for b, idc in enumerate(chosen):
        if len(idc) > fewest:           # every clip masks the same number of frames
            idc = np.random.choice(idc, fewest, replace=False)
        mask[b, idc] = True
        s, e = mask_runs(mask[b])
        starts.append(s)
        ends.append(e)
        owners.append(np.full(len(s), b, dtype=np.int64))
    return mask, np.concatenate(starts), np.concatenate(ends), np.concatenate(owners)


def span_codes_constant(mask, code):
    """Every masked frame takes ``code`` (ZERO or EMB)."""
    codes = np.full(mask.shape, KEEP, dtype=np.int32)
    codes[mask] = code
    return codes


def span_codes_other_clip(mask, perm):
    """selection_type 'same_other_seq' (hubert.py:469-472): masked frame (b, t) takes frame t of clip perm[b]."""
    B, T = mask.shape
    src = (np.asarray(perm, dtype=np.int64)[:, None] * T + np.arange(T, dtype=np.int64)[None, :]).astype(np.int32)
    return np.where(mask, src, np.int32(KEEP)).astype(np.int32)


def span_codes_same_clip(mask, starts, ends, owners):
    """selection_type 'same_seq' (hubert.py:473-486): every masked run takes a run of the same length from elsewhere in
    its clip — the start drawn uniformly (one ``np.random.choice`` per run, in run order) from the frames outside
    [start - length, end), 0 when there are none; source indices clipped to the last frame."""
    B, T = mask.shape
    codes = np.full((B, T), KEEP, dtype=np.int32)
    for b, s, e in zip(owners, starts, ends):
        length = int(e - s)
        allowed = np.setdiff1d(np.arange(T), np.arange(max(0, s - length), e))
        first = int(np.random.choice(allowed, size=1)[0]) if len(allowed) > 0 else 0
        src = np.minimum(np.arange(first, first + length), T - 1)
        codes[b, s:e] = (b * T + src).astype(np.int32)
    return codes
