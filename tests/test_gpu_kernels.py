"""Kernel-level parity on a B200, through the C ABI: the tcgen05 GEMM (all tile widths, ragged M/K, fused
epilogues) and the audio kernels against the float64 oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from multimodalvc_b200 import _lib, audio
from oracle import fbank_oracle as fo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gemm(A, B, bias=None, gelu=False, R=None, c_fp32=True, block_n=0, pair=0, occ=0, inplace=False):
    M, K = A.shape
    N = B.shape[0]
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if c_fp32 else torch.bfloat16)
    if inplace:          # x += A B^T + bias: residual and output are the same buffer (TMA reduce-add epilogue)
        C = R.clone()
        R = C
    vp = ctypes.c_void_p
    _lib.check(_lib.load().avh_gemm_bf16(
        vp(A.data_ptr()), vp(B.data_ptr()), M, N, K, vp(bias.data_ptr()) if bias is not None else None, int(gelu),
        vp(R.data_ptr()) if R is not None else None, int(R is not None and R.dtype == torch.float32),
        vp(C.data_ptr()), int(c_fp32), block_n, pair, occ, vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("pair", [1, 2])
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256), (256, 32, 64, 32),   # single tile, single K block
    (256, 128, 512, 128), (300, 192, 320, 64), (1000, 384, 1024, 160),     # ragged M/N tiles, several K blocks
    (2400, 1024, 1024, 0), (2400, 3072, 1024, 0), (2400, 4096, 1024, 0), (2400, 1024, 4096, 0),   # c2 shapes
    (2400, 4096, 1024, 224), (2400, 1024, 4096, 96),
    (77, 64, 104, 64), (20000, 64, 256, 64),                                # K tail (OOB zero fill), many tiles
])
def test_gemm_matches_fp32_reference(M, N, K, bn, pair):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    C = gemm(A, B, block_n=bn, pair=pair)
    ref = A.float() @ B.float().t()
    assert torch.isfinite(C).all()
    tol = 1e-4 * (K ** 0.5) * 0.05 * 4 + 1e-5          # fp32 accumulation error only
    assert (C - ref).abs().max().item() < max(tol, 2e-3 * ref.abs().max().item() * 1e-2 + tol)


def test_gemm_fused_epilogue_bias_gelu_residual_bf16_out():
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 515, 256, 384
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g)
    ref = torch.nn.functional.gelu(A.float() @ B.float().t() + bias) + R
    C = gemm(A, B, bias=bias, gelu=True, R=R, c_fp32=True)
    assert (C - ref).abs().max().item() < 2e-4
    C1 = gemm(A, B, bias=bias, gelu=True, R=R, c_fp32=True, pair=1)
    assert torch.equal(C, C1)                                 # pair and single-CTA tiles: same arithmetic
    Rb = R.bfloat16()
    Cb = gemm(A, B, bias=bias, gelu=True, R=Rb, c_fp32=False)
    refb = (torch.nn.functional.gelu(A.float() @ B.float().t() + bias) + Rb.float())
    assert (Cb.float() - refb).abs().max().item() < 3e-2      # bf16 output rounding


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (1000, 128, 576, 64), (40000, 64, 576, 64), (30000, 128, 1152, 128),
                                      (2400, 1024, 1024, 128), (500, 96, 320, 32)])
def test_gemm_two_ctas_per_sm(M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ B.float().t() + bias
    C1 = gemm(A, B, bias=bias, block_n=bn, pair=1, occ=1)
    C2 = gemm(A, B, bias=bias, block_n=bn, pair=1, occ=2)
    assert (C1 - ref).abs().max().item() < 2e-3
    assert torch.equal(C1, C2)
    Cb = gemm(A, B, bias=bias, block_n=bn, pair=1, occ=2, c_fp32=False)
    assert (Cb.float() - ref).abs().max().item() < ref.abs().max().item() * 2 ** -8 + 1e-3      # bf16 output rounding


def test_gemm_inplace_residual_uses_reduce_add():
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 2400, 1024, 512
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + A.float() @ B.float().t() + bias
    for occ in (1, 2):
        out = gemm(A, B, bias=bias, R=x, c_fp32=True, inplace=True, occ=occ, block_n=128 if occ == 2 else 0)
        assert (out - ref).abs().max().item() < 2e-4
    assert (gemm(A, B, bias=bias, R=x, c_fp32=True) - ref).abs().max().item() < 2e-4      # out-of-place residual


def test_gemm_many_tiles_per_cta_exercises_both_accumulator_stages():
    g = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 148 * 128 * 3 + 50, 128, 128            # > 3 tiles per CTA: TMEM double buffering + phase flips
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    ref = A.float() @ B.float().t()
    for pair in (1, 2):
        C = gemm(A, B, block_n=128, pair=pair)
        assert (C - ref).abs().max().item() < 1e-3


# ------------------------------------------------------------------------------------------------- audio
def _fbank_dev(wavs, video_lens=None, normalize=True, max_sample_size=None):
    a, pm = audio.logfbank_stack_collate([torch.from_numpy(w) for w in wavs], video_lens=video_lens,
                                         normalize=normalize, max_sample_size=max_sample_size)
    torch.cuda.synchronize()
    return a.transpose(1, 2).cpu().numpy(), pm.cpu().numpy()      # [B,T,104]


def test_logfbank_within_1e4_of_oracle_and_stacking_bit_exact():
    z = np.load(os.path.join(GOLDEN, "audio_fbank.npz"))
    names = [k[4:] for k in z.files if k.startswith("wav_")]
    wavs = [z["wav_" + n] for n in names]
    feats, pm = _fbank_dev(wavs, normalize=False)
    T = feats.shape[1]
    for i, n in enumerate(names):
        ref = fo.stacker(z["fbank_" + n].astype(np.float32), 4)          # [T_i, 104]
        Ti = ref.shape[0]
        assert np.abs(feats[i, :Ti] - ref).max() < 1e-4, n              # gate: 1e-4 abs vs the restatement
        # stacking / zero rows / collation padding are copies: exact zeros where the oracle has zeros
        assert np.array_equal(feats[i, :Ti] == 0, ref == 0), n
        assert not feats[i, Ti:].any()
        assert np.array_equal(pm[i], np.arange(T) >= Ti)               # padding mask bit-exact


def test_logfbank_layernorm_alignment_and_crop():
    z = np.load(os.path.join(GOLDEN, "audio_fbank.npz"))
    wavs = [z["wav_noise_6s"], z["wav_noise_ragged"], z["wav_tone_1k"]]
    vlen = [148, 25, 30]                 # trim by 2, pad by several, pad by 5
    feats, pm = _fbank_dev(wavs, video_lens=vlen, normalize=True, max_sample_size=100)
    assert feats.shape == (3, 100, 104)
    for i, w in enumerate(wavs):
        ref = fo.featurize_clip(w, n_video=vlen[i], normalize=True)
        out, mask = fo.collater_audio([ref], 100)
        assert np.abs(feats[i] - out[0]).max() < 1e-4
        assert np.array_equal(pm[i], mask[0])


def test_logfbank_empty_and_single_frame_inputs():
    feats, pm = _fbank_dev([np.zeros(1, dtype=np.int16), fo.synthetic_wave(400, 1), fo.synthetic_wave(401, 2)],
                           normalize=False)
    assert feats.shape == (3, 1, 104)
    assert np.allclose(feats[0, 0, :26], np.log(np.finfo(float).eps), atol=1e-4)      # all-zero frame -> log(eps)
    assert not feats[0, 0, 26:].any() and not pm.any()
    for i, w in [(1, fo.synthetic_wave(400, 1)), (2, fo.synthetic_wave(401, 2))]:
        ref = fo.stacker(fo.logfbank(w).astype(np.float32), 4)
        assert np.abs(feats[i, :1] - ref).max() < 1e-4


def test_logfbank_full_size_batch_properties():
    # BASELINE config sizes: 32 x 24 s clips; checked through size-independent properties
    wavs = [fo.synthetic_wave(384000, 100 + i) for i in range(4)] * 8
    feats, pm = _fbank_dev(wavs, video_lens=[600] * 32, normalize=True)
    assert feats.shape == (32, 600, 104) and not pm.any()
    assert np.abs(feats.mean(-1)).max() < 1e-4 and np.abs(feats.var(-1) - 1).max() < 1e-3     # per-row LN
    assert np.array_equal(feats[:4], feats[4:8])                                              # deterministic
    ref = fo.featurize_clip(wavs[0], n_video=600)
    assert np.abs(feats[0] - ref).max() < 1e-4


def test_add_noise_matches_reference_within_one_lsb():
    z = np.load(os.path.join(GOLDEN, "audio_reference.npz"))
    clean, noise = torch.from_numpy(z["clean"]), torch.from_numpy(z["noise"])
    for snr, key in [(-5, "mix_snr_m5"), (0, "mix_snr_0"), (5, "mix_snr_5"), (40, "mix_snr_40")]:
        out = audio.add_noise([clean, clean[:20000]], noise, snr)
        torch.cuda.synchronize()
        d = np.abs(out[0].cpu().numpy().astype(np.int32) - z[key].astype(np.int32))
        assert d.max() <= 1, (snr, d.max())                 # stated tolerance: +-1 LSB (fp32 RMS summation order)
        assert (d != 0).mean() < 0.02
        ref2 = fo.add_noise(z["clean"][:20000], z["noise"], snr)
        d2 = np.abs(out[1].cpu().numpy().astype(np.int32) - ref2.astype(np.int32))
        assert d2.max() <= 1
    loud = torch.from_numpy(z["loud"])
    out = audio.add_noise([loud], noise, -5)[0].cpu().numpy().astype(np.int32)
    assert np.abs(out - z["mix_loud"].astype(np.int32)).max() <= 1


def test_logfbank_packed_entry_point_matches_list_entry_point():
    wavs = [fo.synthetic_wave(96000, 11), fo.synthetic_wave(96000, 12), fo.synthetic_wave(96000, 13)]
    a_ref, pm_ref = audio.logfbank_stack_collate([torch.from_numpy(w) for w in wavs], video_lens=[150] * 3)
    flat = torch.from_numpy(np.concatenate(wavs)).cuda()
    off = (torch.arange(4, dtype=torch.int64) * 96000).cuda()
    vl = torch.full((3,), 150, dtype=torch.int32, device="cuda")
    a, pm = audio.logfbank_stack_collate_packed(flat, off, 150, vl)
    assert a.shape == (3, 104, 150) and torch.equal(a, a_ref) and torch.equal(pm, pm_ref)
    with pytest.raises(ValueError):
        audio.logfbank_stack_collate_packed(flat.cpu(), off, 150, vl)


def _attention_ref(qkv, lens, T, D, H):
    """fp32 softmax(q k^T) v per clip and head on the bf16-rounded inputs (q is pre-scaled)."""
    B = len(lens)
    x = qkv.float().view(B, T, 3, H, 64)
    out = torch.zeros(B, T, H, 64)
    for b, n in enumerate(lens):
        q, k, v = x[b, :, 0].transpose(0, 1), x[b, :n, 1].transpose(0, 1), x[b, :n, 2].transpose(0, 1)   # [H,T,64]
        out[b] = (torch.softmax(q @ k.transpose(1, 2), dim=-1) @ v).transpose(0, 1)
    return out.view(B * T, D)


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("T,lens", [(1, [1, 1]), (20, [20, 7]), (75, [75, 1]), (150, [150, 97, 150]), (160, [160, 33]),
                                    (161, [161, 130]), (300, [300, 129, 128]), (600, [600, 417]), (1000, [1000, 513])])
def test_attention_kernels_vs_fp32_reference(impl, T, lens):
    """Both attention kernels (mma.sync, tcgen05 / TMEM) on dense padded batches with key-padding masks: single key
    block (T <= 160), streamed 128-key blocks with the two-pass maxima (T > 160), ragged tails, 1-frame clips."""
    import ctypes
    from multimodalvc_b200 import _lib
    H, D, B = 4, 256, len(lens)
    g = torch.Generator().manual_seed(T)
    qkv = (torch.randn(B * T, 3 * D, generator=g) * 1.5).bfloat16()
    kpm = torch.zeros(B, T, dtype=torch.uint8)
    for b, n in enumerate(lens):
        kpm[b, n:] = 1
    ref = _attention_ref(qkv, lens, T, D, H)
    d_qkv, d_kpm = qkv.cuda(), kpm.cuda()
    out = torch.full((B * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.load().avh_attention_bf16(vp(d_qkv.data_ptr()), vp(d_kpm.data_ptr()), None, B * T, B, T, D, H, impl,
                                              vp(out.data_ptr()), vp(st)))
    torch.cuda.synchronize()
    o = out.float().cpu()
    assert torch.isfinite(o).all()
    assert (o - ref).abs().max() < 0.03 * ref.abs().max() + 1e-3        # bf16 P and bf16 output
    assert torch.nn.functional.cosine_similarity(o.flatten(), ref.flatten(), dim=0) > 0.9999


def test_attention_tc_packed_ragged_batch_equals_dense():
    """cu_rows form (clips packed back to back, no pad rows) gives the same bits as the dense padded form."""
    import ctypes
    from multimodalvc_b200 import _lib
    H, D = 4, 256
    lens = [300, 17, 150, 129, 1]
    T = max(lens)
    B = len(lens)
    g = torch.Generator().manual_seed(5)
    dense = torch.zeros(B, T, 3 * D, dtype=torch.bfloat16)
    packed = []
    for b, n in enumerate(lens):
        x = (torch.randn(n, 3 * D, generator=g) * 1.5).bfloat16()
        dense[b, :n] = x
        packed.append(x)
    packed = torch.cat(packed).cuda()
    kpm = torch.zeros(B, T, dtype=torch.uint8)
    for b, n in enumerate(lens):
        kpm[b, n:] = 1
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    lib = _lib.load()
    o_d = torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)
    o_p = torch.zeros(sum(lens), D, device="cuda", dtype=torch.bfloat16)
    d = dense.view(B * T, 3 * D).cuda()
    k = kpm.cuda()
    _lib.check(lib.avh_attention_bf16(vp(d.data_ptr()), vp(k.data_ptr()), None, B * T, B, T, D, H, 1, vp(o_d.data_ptr()), vp(st)))
    _lib.check(lib.avh_attention_bf16(vp(packed.data_ptr()), None, vp(cu.data_ptr()), sum(lens), B, T, D, H, 1,
                                      vp(o_p.data_ptr()), vp(st)))
    torch.cuda.synchronize()
    r = 0
    for b, n in enumerate(lens):
        assert torch.equal(o_p[r:r + n], o_d.view(B, T, D)[b, :n]), b
        r += n


def test_gemm_stream_k_matches_reference_and_is_deterministic():
    """Stream-K partition (equal k-block ranges per SM, partials added by the tile's owner in unit order) forced on for
    the encoder shapes, single CTAs and CTA pairs: same tolerance as the persistent form, bit-identical reruns."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sk in ("1", "0"):
        env = dict(os.environ, AVH_GEMM_SK=sk)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "gemm_sk_check.py")], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, (sk, r.stdout[-3000:], r.stderr[-2000:])
        assert "FAIL" not in r.stdout
