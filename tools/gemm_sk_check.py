"""Stream-K GEMM check (AVH_GEMM_SK=1 forces the stream-K partition where it applies, 0 switches it off): encoder
shapes at M = 2400 in single-CTA and CTA-pair form, bf16 and fp32 reduce-add outputs, against an fp32 torch reference;
prints one line per case with the max-normalised error and a checksum of the output bits (reruns must agree)."""
import ctypes
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    cases = [(2400, 3072, 1024, 0, 0, 256, 1), (2400, 3072, 1024, 0, 0, 256, 2), (2400, 1024, 1024, 1, 0, 160, 1),
             (2400, 4096, 1024, 0, 1, 192, 1), (2400, 4096, 1024, 0, 1, 256, 2), (2400, 1024, 4096, 1, 0, 160, 1),
             (2400, 1024, 4096, 1, 0, 256, 2), (2400, 1024, 4096, 1, 0, 256, 1), (300, 1024, 1024, 1, 0, 0, 1),
             (4000, 2048, 512, 0, 0, 0, 1)]
    ok = True
    for M, N, K, res, gelu, bn, pair in cases:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).bfloat16().cuda()
        W = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
        bias = torch.randn(N, generator=g).cuda()
        X = torch.randn(M, N, generator=g).cuda()
        ref = A.float() @ W.float().t() + bias
        if gelu:
            ref = torch.nn.functional.gelu(ref)
        if res:
            ref = ref + X
        sums = []
        for rep in range(3):
            C = X.clone() if res else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            _lib.check(lib.avh_gemm_bf16(vp(A.data_ptr()), vp(W.data_ptr()), M, N, K, vp(bias.data_ptr()), gelu,
                                         vp(C.data_ptr()) if res else None, res, vp(C.data_ptr()), res, bn, pair, 1, vp(st)))
            torch.cuda.synchronize()
            sums.append(hashlib.sha1(C.cpu().view(torch.uint8).numpy().tobytes()).hexdigest()[:12])
        err = ((C.float() - ref).abs().max() / ref.abs().max()).item()
        good = err < (2e-5 if res else 1e-2) and len(set(sums)) == 1
        ok &= good
        print(f"M={M} N={N} K={K} res={res} gelu={gelu} bn={bn} pair={pair}: err={err:.2e} bits={sums[0]} "
              f"reruns_identical={len(set(sums)) == 1} {'ok' if good else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
