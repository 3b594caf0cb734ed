"""Gradient all-reduce for the fine-tuning configuration (BASELINE config 5): the B200 form of
fairseq's ``LegacyDistributedDataParallel.all_reduce_grads``
(fairseq/fairseq/distributed/legacy_distributed_data_parallel.py:76-165): gradients are averaged over the ranks of a
process group — pre-divided by the world size, summed, written back in place; parameters without a gradient contribute
zeros and receive the average; parameters tagged ``expert`` are skipped; ``no_sync()`` / ``accumulate_grads`` postpone
the reduction.  One process per GPU, NCCL over NVLink 5 / NVSwitch (gloo on CPU for the tests).

What differs from the reference's loop (one 2**28-element buffer, one blocking all-reduce per fill):
  * buckets of ``bucket_bytes`` (default 32 MiB) carved out of ONE persistent flat buffer, filled in REVERSE parameter
    order (the order backward produces gradients) and reduced asynchronously: bucket k is on the wire while bucket
    k+1 is packed, and every result is copied back only at the end (NVSwitch reduces in the switch, so the cost is
    launch latency + bytes / 725 GB/s, not a per-link ring: many medium buckets overlap better than one 1.3 GB call);
  * with NCCL the 1/world scaling rides inside the collective (PreMulSum) instead of a separate pass over the buffer;
  * ``reduce_bucket_when_ready(param)`` lets a backward implementation hand over gradients as they are produced.
The numerical result is the reference's: mean over ranks in the gradient dtype (sum order is the collective's).
"""
from contextlib import contextmanager
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, buffer_size: int = 2 ** 28,
                 bucket_bytes: int = 32 << 20):
        if isinstance(params, torch.nn.Module):
            params = params.parameters()
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("no parameters to reduce")
        self.process_group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        # never use a bigger buffer than the number of model params (reference :48-49)
        self.buffer_size = min(int(buffer_size), total)
        self.bucket_bytes = int(bucket_bytes)
        self.buffer: Optional[torch.Tensor] = None
        self.accumulate_grads = False
        self._side = None                 # stream the overlapped bucket reductions are issued from
        self._presynced = set()           # ids of parameters whose gradients the backward already averaged this step

    @contextmanager
    def no_sync(self):
        """Disable gradient synchronisation inside the context (reference :64-70)."""
        old = self.accumulate_grads
        self.accumulate_grads = True
        try:
            yield
        finally:
            self.accumulate_grads = old

    # ------------------------------------------------------------------------------- overlap with the device backward
    def attach(self, module):
        """Overlap the reduction with `module`'s device backward (TransformerEncoder / AVHubertModel with trainable
        weights): the library hands the flat gradient buffer over bucket by bucket (avh_encoder_backward_buckets) and every
        bucket's all-reduce is issued from a side stream as soon as its event is recorded.  all_reduce_grads() then only
        reduces what the backward did not cover (other modules' parameters).  Returns self."""
        module._grad_sync = self
        return self

    def reduce_buckets(self, lib, handle, flat) -> bool:
        """Called by the backward (multimodalvc_b200/_train.py) right after the library call returned (all launches are
        enqueued, none needs to have run).  Averages `flat` over the group in place, bucket by bucket; the caller's stream
        waits for the last collective before anything reads the gradients.  False = nothing done (single rank, no_sync,
        or not an NCCL group): all_reduce_grads() will do the work."""
        import ctypes
        if self.accumulate_grads or self.world_size == 1 or not flat.is_cuda:
            return False
        if dist.get_backend(self.process_group) != "nccl":
            return False
        from . import _lib
        n = ctypes.c_int32()
        _lib.check(lib.avh_grad_bucket_count(handle, ctypes.byref(n)))
        if n.value == 0:
            return False
        dev = flat.device
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        flat.record_stream(self._side)
        works, covered = [], 0
        for k in range(n.value):
            b, e = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(lib.avh_grad_bucket_range(handle, k, ctypes.byref(b), ctypes.byref(e)))
            _lib.check(lib.avh_grad_bucket_wait(handle, k, ctypes.c_void_p(self._side.cuda_stream)))
            with torch.cuda.stream(self._side):        # the collective is ordered after the side stream, i.e. after the event
                works.append(self._reduce_async(flat[b.value:e.value], True))
            covered += e.value - b.value
        if covered != flat.numel():
            raise RuntimeError("gradient buckets do not cover the flat buffer")
        for w in works:
            if w is not None:
                w.wait()                               # the current stream waits; the host does not
        return True

    def mark_reduced(self, params):
        self._presynced.update(id(p) for p in params)

    # ------------------------------------------------------------------------------------------------ helpers
    def _reduce_async(self, flat: torch.Tensor, nonzero: bool):
        """Average `flat` over the group in place; returns a work handle (or None for a single rank)."""
        if self.world_size == 1:
            return None
        backend = dist.get_backend(self.process_group)
        if backend == "nccl" and nonzero and hasattr(dist, "_make_nccl_premul_sum"):
            op = dist._make_nccl_premul_sum(1.0 / self.world_size)       # scaling folded into the collective
            return dist.all_reduce(flat, op=op, group=self.process_group, async_op=True)
        if nonzero:
            flat.div_(self.world_size)
        return dist.all_reduce(flat, group=self.process_group, async_op=True)

    def _buckets(self, params):
        """Parameter lists whose element counts fit the bucket size, in reverse registration order."""
        first = params[0]
        cap = max(1, min(self.buffer_size, self.bucket_bytes // max(1, first.element_size())))
        cur, n = [], 0
        for p in reversed(params):
            sz = p.numel()
            if sz > cap:                      # big parameter: reduced on its own (reference :143-145)
                if cur:
                    yield cur
                    cur, n = [], 0
                yield [p]
                continue
            if n + sz > cap:
                yield cur
                cur, n = [], 0
            cur.append(p)
            n += sz
        if cur:
            yield cur

    # ------------------------------------------------------------------------------------------------ the reduction
    def all_reduce_grads(self):
        """Call after backward (reference :76-165).  Returns the number of collectives issued."""
        if self.accumulate_grads:
            return 0
        by_device = {}
        presynced, self._presynced = self._presynced, set()
        for p in self.params:
            if not p.requires_grad or hasattr(p, "expert") or id(p) in presynced:
                continue
            if p.grad is not None and p.grad.requires_grad:
                raise RuntimeError("gradient all-reduce only works with gradients that don't require grad")
            by_device.setdefault((p.device, p.dtype), []).append(p)
        issued = 0
        for (device, dtype), params in by_device.items():
            total = sum(p.numel() for p in params)
            if self.buffer is None or self.buffer.device != device or self.buffer.dtype != dtype or self.buffer.numel() < total:
                self.buffer = torch.empty(total, device=device, dtype=dtype)
            offset = 0
            pending = []
            for bucket in self._buckets(params):
                n = sum(p.numel() for p in bucket)
                flat = self.buffer[offset:offset + n]
                offset += n
                nonzero = False
                o = 0
                dsts, srcs = [], []
                for p in bucket:
                    sz = p.numel()
                    if p.grad is not None:
                        dsts.append(flat[o:o + sz].view_as(p))
                        srcs.append(p.grad.detach())
                        nonzero = True
                    else:
                        flat[o:o + sz].zero_()
                    o += sz
                if dsts:
                    torch._foreach_copy_(dsts, srcs)          # one multi-tensor copy per bucket, not one kernel per parameter
                pending.append((self._reduce_async(flat, nonzero), flat, bucket))
                issued += 1
            # copy the averaged gradients back into their original place (reference :114-121)
            for work, flat, bucket in pending:
                if work is not None:
                    work.wait()
                o = 0
                dsts, srcs = [], []
                for p in bucket:
                    sz = p.numel()
                    if p.grad is not None:
                        dsts.append(p.grad.detach())
                        srcs.append(flat[o:o + sz].view_as(p))
                    else:
                        p.grad = flat[o:o + sz].view_as(p).clone()
                    o += sz
                if dsts:
                    torch._foreach_copy_(dsts, srcs)
        return issued
