"""Attention core alone: mma.sync kernel (impl 0) vs tcgen05 / TMEM kernel (impl 1), Large head count (16 x 64),
back-to-back launches timed with CUDA events.  FLOPs = 4 B H T^2 64 (QK^T + PV, dense)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    H, D = 16, 1024
    for B, T in [(16, 150), (8, 300), (4, 600), (2, 1200), (64, 38)]:
        qkv = [(torch.randn(B * T, 3 * D, device="cuda") * 1.2).bfloat16() for _ in range(4)]
        out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
        res = {}
        for impl in (0, 1):
            def run(i):
                _lib.check(lib.avh_attention_bf16(vp(qkv[i % 4].data_ptr()), None, None, B * T, B, T, D, H, impl,
                                                  vp(out.data_ptr()), vp(st)))
            for i in range(5):
                run(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            res[impl] = e0.elapsed_time(e1) * 10.0      # us per launch
        fl = 4.0 * B * H * T * T * 64
        print(f"B={B:3d} T={T:4d}: mma.sync {res[0]:7.2f} us ({fl / res[0] / 1e6:6.1f} TFLOP/s)   tcgen05 {res[1]:7.2f} us "
              f"({fl / res[1] / 1e6:6.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
