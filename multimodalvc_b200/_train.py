"""Host side of the device backward shared by the three autograd wrappers (TransformerEncoder, the fusion tail, the whole
AVHubertModel): one call into the library that fills a flat gradient buffer in the parameters' dtype bucket by bucket,
and — when a GradientAllReducer is attached to the module — the all-reduce of every bucket issued on a side stream as
soon as the library has recorded the bucket's event, so that the collective overlaps the rest of the backward
(fairseq legacy_distributed_data_parallel.py:76-165 does the reduction after the backward; same result)."""
import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.AVH_F32, torch.float16: _lib.AVH_F16, torch.bfloat16: _lib.AVH_BF16}


def backward_flat(handle, dout, dx, n_floats, dtypes, fwd_stream, owner):
    """Runs the library backward of the last training forward on `handle`.  Returns {dtype: flat gradient buffer}: when
    every parameter has the same dtype the buffer is written in that dtype directly (and all-reduced bucket by bucket if
    `owner._grad_sync` is an active GradientAllReducer); otherwise fp32, converted once per dtype on demand."""
    dev = dout.device
    lib = _lib.load()
    vp = ctypes.c_void_p
    single = dtypes[0] if dtypes and all(d == dtypes[0] for d in dtypes) and dtypes[0] in _DTYPES else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        if stream != fwd_stream:
            raise RuntimeError("the backward must run on the CUDA stream of its forward")
        dxp = vp(dx.data_ptr()) if dx is not None else None
        dxd = _DTYPES[dx.dtype] if dx is not None else 0
        if single is None:
            flat = torch.empty(n_floats, device=dev, dtype=torch.float32)
            _lib.check(lib.avh_encoder_backward(handle, vp(dout.data_ptr()), _DTYPES[dout.dtype], dxp, dxd, vp(flat.data_ptr()),
                                                n_floats, vp(stream)))
            return _FlatGrads(flat)
        flat = torch.empty(n_floats, device=dev, dtype=single)
        _lib.check(lib.avh_encoder_backward_buckets(handle, vp(dout.data_ptr()), _DTYPES[dout.dtype], dxp, dxd,
                                                    vp(flat.data_ptr()), _DTYPES[single], n_floats, vp(stream)))
        sync = getattr(owner, "_grad_sync", None)
        reduced = sync is not None and sync.reduce_buckets(lib, handle, flat)
    out = _FlatGrads(flat)
    out.reduced = bool(reduced)
    return out


class _FlatGrads:
    """The flat buffer in one dtype, other dtypes converted once on demand."""

    def __init__(self, flat):
        self.base = flat
        self.by_dtype = {flat.dtype: flat}
        self.reduced = False

    def of(self, dt):
        if dt not in self.by_dtype:
            self.by_dtype[dt] = self.base.to(dt)
        return self.by_dtype[dt]


def refresh_weights_device(handle, module, prefix=""):
    """After an optimizer step: rewrite the packed weights the training plans read from `module`'s device-resident
    parameters, in place and on the device (avh_refresh_weights_device) — plans and captured graphs stay valid, nothing
    crosses the host.  `prefix` maps the module's parameter names onto the state-dict keys the handle was loaded with."""
    lib = _lib.load()
    names, tensors = [], []
    for name, p in module.named_parameters():
        t = p.detach()
        if not t.is_floating_point():
            continue
        if t.dtype not in _DTYPES:
            t = t.float()
        names.append((prefix + name).encode())
        tensors.append(t.contiguous())
    n = len(names)
    c_names = (ctypes.c_char_p * n)(*names)
    c_ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    c_dts = (ctypes.c_int32 * n)(*[_DTYPES[t.dtype] for t in tensors])
    c_num = (ctypes.c_int64 * n)(*[t.numel() for t in tensors])
    dev = tensors[0].device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.avh_refresh_weights_device(handle, c_names, c_ptrs, c_dts, c_num, n, ctypes.c_void_p(stream)))
