"""Stall accounting of the encoder GEMM shapes (M = 2400): per CTA, SM clocks from the first MMA to the last commit
(main loop), cycles the MMA warp waited for operands (full barriers) and for a free accumulator, cycles the TMA
producer waited for a free stage, and the epilogue tail after the last MMA.  Weights rotate (DRAM-resident as in
the pipeline); the traced launch follows 6 untraced ones."""
import ctypes
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib  # noqa: E402


def clk():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                              capture_output=True, text=True).stdout.strip().split("\n")[0]
    except Exception:
        return "?"


def main():
    M = 2400
    lib = _lib.load()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    cfgs = [("qkv", 3072, 1024, 0, 256, 1), ("qkv", 3072, 1024, 0, 256, 2), ("out_proj", 1024, 1024, 1, 160, 1),
            ("fc1", 4096, 1024, 0, 192, 1), ("fc1", 4096, 1024, 0, 256, 2), ("fc2", 1024, 4096, 1, 160, 1),
            ("fc2", 1024, 4096, 1, 256, 2), ("fc2", 1024, 4096, 1, 256, 1)]
    for name, N, K, res, bn, pair in cfgs:
        A = torch.randn(M, K, device="cuda").bfloat16()
        Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(8)]
        bias = torch.zeros(N, device="cuda")
        C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if res else torch.bfloat16)

        def run(i):
            _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(Ws[i % 8].data_ptr()), M, N, K, vp(bias.data_ptr()), 0,
                                         vp(C.data_ptr()) if res else None, res, vp(C.data_ptr()), res, bn, pair, 1, vp(st)))
        for i in range(6):
            run(i)
        torch.cuda.synchronize()
        lib.avh_gemm_set_trace(vp(trace.data_ptr()))
        trace.zero_()
        run(6)
        torch.cuda.synchronize()
        c = clk()
        lib.avh_gemm_set_trace(None)
        t = trace.view(148, 16).cpu().double()
        lead = t[:, 3] > 0
        n = int(lead.sum())
        main_loop = (t[lead, 5] - t[lead, 3])
        total = (t[lead, 10] - t[lead, 0])
        tail = (t[lead, 10] - t[lead, 5])
        fill = (t[lead, 3] - t[lead, 0])
        kb = K // 64
        mt = (M + 128 * pair - 1) // (128 * pair)
        nt = (N + bn - 1) // bn
        tiles = mt * nt
        units = 148 // pair
        rounds = (tiles + units - 1) // units
        drain = (t[lead, 8] - t[lead, 5])        # last MMA issued -> last epilogue starts (MMA drain)
        epi = (t[lead, 9] - t[lead, 8])          # last epilogue
        fin = (t[lead, 10] - t[lead, 9])         # store drain + final sync
        print(f"   tail split: mma drain {drain.mean():6.0f}  last epilogue {epi.mean():6.0f}  store drain + sync {fin.mean():6.0f}"
              f" | fill split: prologue+wait {(t[lead, 1] - t[lead, 0]).mean():6.0f} first TMA->first full "
              f"{(t[lead, 3] - t[lead, 1]).mean():6.0f}")
        print(f"{name:9s} pair={pair} BN={bn}: leaders={n} tiles={tiles} rounds={rounds} clk={c} MHz | total {total.mean():7.0f} "
              f"fill {fill.mean():6.0f} main {main_loop.mean():7.0f} (max {main_loop.max():7.0f}) tail {tail.mean():6.0f} | "
              f"mma wait operands {t[lead, 12].mean():7.0f} wait tmem {t[lead, 13].mean():6.0f} | producer wait "
              f"{t[t[:, 14] > 0, 14].mean() if (t[:, 14] > 0).any() else 0:7.0f} | per k-block of a full-round CTA "
              f"{main_loop.max() / (rounds * kb):5.0f} clk (tensor {2 * bn})", flush=True)


if __name__ == "__main__":
    main()
