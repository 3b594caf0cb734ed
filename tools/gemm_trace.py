"""Device-side timeline of one tcgen05 GEMM launch (globaltimer stamps per CTA)."""
import ctypes, sys, os
import torch
OCC = int(__import__("os").environ.get("PROBE_OCC", "0"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib
from gemm_probe import SHAPES

NAMES = ["entry", "prologue_done", "first_tma", "first_full", "tile0_mma_done", "last_mma_done", "epi0_start",
         "epi0_end", "epiL_start", "epiL_end", "final_sync", "dealloc", "x12", "x13", "x14", "x15"]

def main():
    pair = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    lib = _lib.load(); vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    var = os.environ.get("TRACE_VARIANT", "")
    for name, M, N, K, bn, gelu, res, cf in SHAPES[:4]:
        use_bias = var not in ("none", "res")
        if var in ("none", "bias"): res = 0
        if var == "nogelu": gelu = 0
        A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda") if res else None
        C = torch.empty(M, N, device="cuda", dtype=torch.float32 if cf else torch.bfloat16)
        if res and os.environ.get("TRACE_INPLACE", "1") == "1":
            C = R              # x += A B^T + bias: the TMA reduce-add epilogue, as in the model
        def run():
            _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(B.data_ptr()), M, N, K, vp(bias.data_ptr()) if use_bias else None, gelu,
                                         vp(R.data_ptr()) if res else None, 1, vp(C.data_ptr()), cf, bn, pair, OCC, vp(st)))
        for _ in range(3): run()
        torch.cuda.synchronize()
        lib.avh_gemm_set_trace(vp(trace.data_ptr()))
        trace.zero_(); run(); torch.cuda.synchronize()
        lib.avh_gemm_set_trace(None)
        t = trace.view(148, 16).cpu()
        used = t[:, 0] > 0
        print(f"== {name} M={M} N={N} K={K} ctas={int(used.sum())}  (SM clocks / 1000, per-CTA origin)")
        for cta in (0, 1, int(used.sum()) - 1):
            row = t[cta]
            t0 = row[0].item()
            print(f"  cta {cta}: " + " ".join(f"{NAMES[i]}={(row[i].item()-t0)/1e3:.2f}" for i in range(16) if row[i] > 0))

if __name__ == "__main__":
    main()
