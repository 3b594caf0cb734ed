"""Standalone launches of the tcgen05 GEMM at the encoder's shapes (for ncu / event timing)."""
import ctypes
import sys
import torch
OCC = int(__import__("os").environ.get("PROBE_OCC", "0"))
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from multimodalvc_b200 import _lib

SHAPES = [  # name, M, N, K, block_n, gelu, residual(fp32), c_fp32
    ("fc1", 2400, 4096, 1024, 0, 1, 0, 0),
    ("fc2", 2400, 1024, 4096, 0, 0, 1, 1),
    ("qkv", 2400, 3072, 1024, 0, 0, 0, 0),
    ("out", 2400, 1024, 1024, 0, 0, 1, 1),
    ("big", 8192, 8192, 8192, 0, 0, 0, 0),
    ("n64", 300000, 64, 576, 0, 0, 0, 0),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    pair = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib = _lib.load()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    for name, M, N, K, bn, gelu, res, cf in SHAPES:
        A = torch.randn(M, K, device="cuda").bfloat16()
        B = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda") if res else None
        C = torch.empty(M, N, device="cuda", dtype=torch.float32 if cf else torch.bfloat16)
        def run():
            _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(B.data_ptr()), M, N, K, vp(bias.data_ptr()), gelu,
                                         vp(R.data_ptr()) if res else None, 1, vp(C.data_ptr()), cf, bn, pair, OCC, vp(st)))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name}: M={M} N={N} K={K} {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
