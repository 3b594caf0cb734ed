"""Steady-state main-loop time per k-block of the tcgen05 GEMM for (pair, BN): 8 tiles per work unit, K = 1024, operands L2-resident."""
import ctypes, sys, os, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib

def clk():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                              capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"

def main():
    lib = _lib.load(); vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    K = 1024
    NT = 1
    REP = 24
    os.environ["AVH_GEMM_KREPEAT"] = str(REP)
    for pair in (1, 2):
        for bn in (32, 64, 96, 128, 160, 192, 224, 256):
            M = 128 * 148
            N = bn * NT
            A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
            C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            def run():
                _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(B.data_ptr()), M, N, K, None, 0, None, 0,
                                             vp(C.data_ptr()), 0, bn, pair, 0, vp(st)))
            for _ in range(3): run()
            torch.cuda.synchronize()
            lib.avh_gemm_set_trace(vp(trace.data_ptr())); trace.zero_(); run(); torch.cuda.synchronize()
            c = clk()
            lib.avh_gemm_set_trace(None)
            t = trace.view(148, 16).cpu()
            lead = t[:, 3] > 0
            ml = (t[lead, 5] - t[lead, 3]).double().mean().item() / (REP * K // 64)      # SM clocks per k-block
            ghz = float(c) / 1e3 if c.replace(".", "").isdigit() else 1.9
            fl = 2.0 * 128 * pair * bn * 64
            print(f"pair={pair} BN={bn}: {ml:.0f} clk/k-block ({ml/ghz:.0f} ns at {c} MHz) -> "
                  f"{fl/(ml/ghz)/1e3*148/pair:.0f} TFLOP/s chip, {100*2.0*bn/ml:.0f}% of the tensor pipe", flush=True)

if __name__ == "__main__":
    main()
