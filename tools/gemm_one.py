"""One encoder-shaped tcgen05 GEMM launch, repeated (ncu target): python gemm_one.py <name> [pair] [reps]"""
import ctypes, sys, os
import torch
OCC = int(__import__("os").environ.get("PROBE_OCC", "0"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib
from gemm_probe import SHAPES
name = sys.argv[1]; pair = int(sys.argv[2]) if len(sys.argv) > 2 else 0; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lib = _lib.load(); vp = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
for nm, M, N, K, bn, gelu, res, cf in SHAPES:
    if nm != name: continue
    A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda") if res else None
    C = torch.empty(M, N, device="cuda", dtype=torch.float32 if cf else torch.bfloat16)
    for _ in range(reps):
        _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(B.data_ptr()), M, N, K, vp(bias.data_ptr()), gelu,
                                     vp(R.data_ptr()) if res else None, 1, vp(C.data_ptr()), cf, bn, pair, OCC, vp(st)))
    torch.cuda.synchronize()
    print("done", nm)
