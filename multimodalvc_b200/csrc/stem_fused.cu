// Lip-frontend stem in ONE kernel: Conv3d(1->64, 5x7x7, stride 1x2x2, pad 2x3x3) + BatchNorm3d + PReLU +
// MaxPool3d(1x3x3, stride 1x2x2, pad 0x1x1)   (avhubert/resnet.py:136-141), video [B,1,T,88,88] -> pooled maps in the
// shared-zero-padded NHWC layout [frames, 23*23, 64] that the layer1 convolutions read (conv_window.cu).
//
// The unfused chain (patch matrix -> tcgen05 GEMM -> pool) moved 4 x 600 MB through HBM per 2400 frames and was bound
// by it.  Here the only HBM traffic is the video (15.5 KB per frame) and the pooled maps (66 KB per frame).
//
//   * work item = (clip, band of 2 pooled rows, time segment).  A band needs 5 stem rows x 44 columns of stem pixels
//     (one halo row shared with the band above) and 15 input rows.
//   * builder warps turn the input rows of ONE frame into the band's patch tile (pixel rows x 64 bf16, K = kh*8 + kw
//     so that a 16-byte chunk is 8 consecutive input pixels; 128-byte-swizzled K-major, generic stores +
//     fence.proxy.async) in a ring of 7 frame slots.  The CTA walks time: every patch tile is built once and used by
//     the temporal taps of 5 consecutive output frames.
//   * the GEMM is TRANSPOSED: D[channel, pixel] = W[channel, k] * P[pixel, k]^T with the WEIGHTS as the A operand
//     resident in TENSOR MEMORY (tcgen05.mma with A in TMEM) and the patch tile as B.  A 64-channel A would use half
//     the 128-row datapath, so TWO consecutive output frames are stacked: for the pair (t, t+1) and input frame
//     f = t-2+i, A_i = [W_i ; W_(i-1)] (W_-1 = W_5 = 0) — 6 variants x 32 TMEM columns.  Per pair: <= 6 frames x 4
//     K-steps per pixel half, every MMA 128 x N x 16 at the full tensor rate, and the only shared-memory operand
//     traffic is the patch tile (the first version, pixels as M and N = 64, re-read its A tile for 64 output columns
//     and was shared-memory-bandwidth bound at 3300 clk per frame).
//   * pixels are split into two halves by column (x <= 21 | x >= 21, column 21 in both) so that every pooling window
//     lies inside one half; the halves are the two accumulator stages (N = 112 and 128): the epilogue of one overlaps
//     the MMAs of the other.
//   * 8 epilogue warps, thread = (frame of the pair, channel): tcgen05.ld of the three stem rows of one pooled row ->
//     BN scale/bias + PReLU with per-thread constants -> 3x3/stride-2 max entirely in registers -> bf16 stores
//     (a warp writes 64 contiguous bytes per pooled pixel).  Out-of-image pool taps are replaced by a duplicate of
//     an in-window tap (nn.MaxPool3d pads with -inf).
#include "common.cuh"
#include "gemm.h"
#include "kernels.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace avh {
namespace {

constexpr int NT = 512;                     // 16 warps: 0, 2 MMA issuers, 1 TMEM alloc, 3 spare, 4-11 epilogue, 12-15 builders
constexpr int RING = 7;
constexpr int H0_ROWS = 112;                // half 0: 5 x 22 pixels (x 0..21) + 2 zero rows
constexpr int H1_ROWS = 128;                // half 1: 5 x 23 pixels (x 21..43) + 13 zero rows
constexpr int SLOT_BYTES = (H0_ROWS + H1_ROWS) * 128;     // 30 KB
constexpr int H1_OFF = H0_ROWS * 128;
constexpr int NBUILD = 110 + 115;           // patch rows built per frame
constexpr int SIN_PITCH = 96;               // 3 zero columns + 88 + 5 zero columns
constexpr int SIN_ROWS = 15;
constexpr int SIN_BYTES = 3072;
constexpr size_t SMEM_BYTES = 1024 + RING * SLOT_BYTES + SIN_BYTES + 256;
constexpr uint32_t TM_A = 0, TM_D0 = 192, TM_D1 = 320;      // TMEM columns: 6 x 32 weights | D half 0 | D half 1

struct StemParams {
  const void* video;
  int in_dt;
  int T, b0, nb;
  int nseg, seglen, num_items;
  const int4* items;           // ragged batches: device item list {clip, band, t0, t1} (num_items read from n_items), else null
  const int* cu;               // ragged batches: [clips + 1] first output frame of every clip; T is then the input clip pitch
  const int* n_items;          // ragged batches: device int holding the number of items of this call
  __nv_bfloat16* out;          // [nb*T, 529, 64]
  const __nv_bfloat16* w;      // [64, 5*64], K = dt*64 + kh*8 + kw
  const float* scale;
  const float* bias;
  const float* slope;
  unsigned long long* dbg;     // debug (AVH_STEM_DBG=1): per CTA 8 stall counters in SM clocks, else null
};

// mbarrier wait that charges the stall to a per-role counter when the debug buffer is present
__device__ __forceinline__ void wait_acc(uint64_t* bar, uint32_t parity, unsigned long long* dbg, long long& acc) {
  if (dbg == nullptr) { mbar_wait(bar, parity); return; }
  const long long c0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - c0;
}

struct Item {
  int j, t0, t1, f_first, f_last;
  int T;             // frames of the item's clip
  long long in0;     // index of the clip's first frame in the input video
  long long out0;    // ... and in the output maps
};
__device__ __forceinline__ Item decode_item(const StemParams& p, int item) {
  Item it;
  if (p.items != nullptr) {
    // ragged batch: (clip, band, segment) list written by the host for this call; clips sit Tstride frames apart in
    // the (padded) input and back to back in the output
    const int4 e = __ldg(p.items + item);           // {clip, band, t0, t1}
    const int c0 = __ldg(p.cu + e.x);
    it.j = e.y; it.t0 = e.z; it.t1 = e.w;
    it.T = __ldg(p.cu + e.x + 1) - c0;
    it.in0 = (long long)e.x * p.T;
    it.out0 = c0;
  } else {
    const int sg = item % p.nseg;
    const int rest = item / p.nseg;
    it.j = rest % 11;
    const int bl = rest / 11;
    it.t0 = sg * p.seglen;
    it.T = p.T;
    it.t1 = min(p.T, it.t0 + p.seglen);
    it.in0 = (long long)(p.b0 + bl) * p.T;
    it.out0 = (long long)bl * p.T;
  }
  it.f_first = max(0, it.t0 - 2);
  it.f_last = min(it.T - 1, it.t1 + 1);
  return it;
}

// One pixel half of one frame pair for this thread's (frame, channel): three stem rows of pooled row `prow`.
template <int HALF>
__device__ __forceinline__ void epilogue_half(uint32_t taddr, int prow, bool skip_top, float sc, float bi, float sl,
                                              __nv_bfloat16* obase, bool do_store, uint64_t* empty_bar, int lane) {
  constexpr int W = HALF ? 23 : 22;
  const uint32_t ta = taddr + 2 * prow * W;
  uint32_t r0[32], r1[32], r2[8];
  tmem_ld_32x32(ta, r0);
  tmem_ld_32x32(ta + 32, r1);
  tmem_ld_32x8(ta + 64, r2);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);       // accumulator half is free for the next pair
  float v[3 * W];
#pragma unroll
  for (int i = 0; i < 3 * W; ++i) {
    const uint32_t u = i < 32 ? r0[i] : (i < 64 ? r1[i - 32] : r2[i - 64]);
    const float x = fmaf(__uint_as_float(u), sc, bi);
    v[i] = x > 0.f ? x : x * sl;
  }
#pragma unroll
  for (int pl = 0; pl < 11; ++pl) {
    float m = -INFINITY;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int lc = 2 * pl + d + (HALF ? 0 : -1);
      if (lc < 0) lc = 0;                      // column -1: duplicate of column 0
      const float top = skip_top ? v[W + lc] : v[lc];      // stem row -1 (band 0): duplicate of the row below
      m = fmaxf(m, fmaxf(top, fmaxf(v[W + lc], v[2 * W + lc])));
    }
    if (do_store) obase[pl * 64] = __float2bfloat16_rn(m);
  }
}

__global__ void __launch_bounds__(NT, 1)
stem_fused_kernel(const StemParams p_in) {
  StemParams p = p_in;
  if (p.n_items != nullptr) p.num_items = __ldg(p.n_items);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* ring = smem;
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(ring + RING * SLOT_BYTES);
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sIn) + SIN_BYTES);
  uint64_t* slot_free = patch_full + RING;
  uint64_t* d_full = slot_free + RING;
  uint64_t* d_empty = d_full + 2;
  uint64_t* a_ready = d_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const long long k_start = clock64();
  long long st0 = 0, st1 = 0;

  // launch-constant setup: zero the ring (K columns 56..63 and the pad rows stay zero for good) and the input pads
  for (int i = threadIdx.x; i < RING * SLOT_BYTES / 16; i += NT) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < SIN_BYTES / 4; i += NT) reinterpret_cast<uint32_t*>(sIn)[i] = 0u;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(&patch_full[s], 128);
      mbar_init(&slot_free[s], 2);        // one commit per MMA-issuing warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d_full[s], 1);
      mbar_init(&d_empty[s], 8);
    }
    mbar_init(a_ready, 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();          // the zero-filled ring is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 0 -> pixel half 0, warp 2 ->
    // half 1.  One elected lane needs ~430 clk per block of 4 MMAs (issue + barrier handling); with one issuer the 12
    // blocks of a frame pair took 5100 clk against 3400 clk of tensor-pipe time.
    const int h = warp >> 1;
    const uint32_t idesc = h ? umma_idesc_bf16(128, H1_ROWS) : umma_idesc_bf16(128, H0_ROWS);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    int n0 = 0, waited = 0, pc = 0;
    const uint32_t ring_base = smem_u32(ring);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const Item im = decode_item(p, item);
      int next_free = im.f_first;
      for (int t = im.t0; t < im.t1; t += 2, ++pc) {
        const bool two = t + 1 < im.t1;
        const int lo = max(t - 2, 0), hi = min(t + (two ? 3 : 2), im.T - 1);
        {
          wait_acc(&d_empty[h], (pc & 1) ^ 1, p.dbg, st1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (h ? TM_D1 : TM_D0);
          for (int f = lo; f <= hi; ++f) {        // oldest frame first: the newest one may still be under construction
            const int m = n0 + f - im.f_first;
            while (waited <= m) {
              wait_acc(&patch_full[waited % RING], (waited / RING) & 1, p.dbg, st0);
              ++waited;
            }
            tc_fence_after();
            const uint32_t b_addr = ring_base + (m % RING) * SLOT_BYTES + h * H1_OFF;
            const uint32_t a_addr = tmem_base + TM_A + 32 * (f - t + 2);
            const uint64_t bdesc = umma_desc_sw128(b_addr);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts(tmem_d, a_addr + 8 * k, bdesc + 2 * k, idesc, (f != lo) || (k != 0));
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&d_full[h]);
          __syncwarp();
        }
        // frames below the next pair's window are dead; at the end of the item everything is
        const int upto = (t + 2 >= im.t1) ? im.f_last + 1 : t;
        while (next_free < upto) {
          if (elect_one()) umma_commit(&slot_free[(n0 + next_free - im.f_first) % RING]);
          __syncwarp();
          ++next_free;
        }
      }
      n0 += im.f_last - im.f_first + 1;
    }
    if (p.dbg != nullptr && lane == 0 && h == 0) { p.dbg[blockIdx.x * 8 + 2] = st0; p.dbg[blockIdx.x * 8 + 3] = st1; }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ weights -> TMEM (once), then the epilogue
    const int q = warp & 3;                                // TMEM lane quarter
    const int prow = (warp - 4) >> 2;                      // pooled row of the band handled by this warp
    const int lane_m = q * 32 + lane;                      // accumulator row: frame of the pair * 64 + channel
    const int fp = lane_m >> 6, ch = lane_m & 63;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    if (prow == 0) {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int dt = i - fp;
        uint32_t r[32];
        if (dt >= 0 && dt <= 4) {
          const uint4* src = reinterpret_cast<const uint4*>(p.w + ch * 320 + dt * 64);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(src + c);
            r[4 * c] = u.x; r[4 * c + 1] = u.y; r[4 * c + 2] = u.z; r[4 * c + 3] = u.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = 0u;
        }
        tmem_st_32x32(lane_base + TM_A + 32 * i, r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(a_ready);
    }
    const float sc = __ldg(p.scale + ch), bi = __ldg(p.bias + ch), sl = __ldg(p.slope + ch);
    int pc = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const Item im = decode_item(p, item);
      const bool skip_top = im.j == 0 && prow == 0;
      for (int t = im.t0; t < im.t1; t += 2, ++pc) {
        const bool do_store = t + fp < im.t1;
        __nv_bfloat16* orow = p.out + (((im.out0 + t + fp) * 23 + (2 * im.j + prow)) * 23) * 64 + ch;
        wait_acc(&d_full[0], pc & 1, p.dbg, st0);
        tc_fence_after();
        epilogue_half<0>(lane_base + TM_D0, prow, skip_top, sc, bi, sl, orow, do_store, &d_empty[0], lane);
        wait_acc(&d_full[1], pc & 1, p.dbg, st1);
        tc_fence_after();
        epilogue_half<1>(lane_base + TM_D1, prow, skip_top, sc, bi, sl, orow + 11 * 64, do_store, &d_empty[1], lane);
      }
    }
    if (p.dbg != nullptr && threadIdx.x == 128) { p.dbg[blockIdx.x * 8 + 4] = st0; p.dbg[blockIdx.x * 8 + 5] = st1; }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ builders: input rows -> patch tiles
    const int bt = threadIdx.x - 384;                      // 0..127
    const bool fast = p.in_dt == DT_BF16;
    int item = blockIdx.x;
    Item im;
    int f = 0;
    bool valid = item < p.num_items;
    if (valid) { im = decode_item(p, item); f = im.f_first; }
    uint4 rv[2];                                           // bf16 fast path: 16-byte pieces bt and bt + 128 of the 165
    float rf[11];                                          // other dtypes: elements bt + 128 k of the 1320
    auto prefetch = [&](const Item& m, int ff) {
      const long long src = (m.in0 + ff) * 7744;
      const int h0 = 8 * m.j - 5;
      if (fast) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = bt + 128 * k;
          const int r = i / 11, c8 = i - r * 11;
          const int hh = h0 + r;
          rv[k] = make_uint4(0, 0, 0, 0);
          if (i < 165 && hh >= 0 && hh < 88)
            rv[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.video) + src + hh * 88) + c8);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const int i = bt + 128 * k;
          const int r = i / 88, cc = i - r * 88;
          const int hh = h0 + r;
          rf[k] = 0.f;
          if (i < SIN_ROWS * 88 && hh >= 0 && hh < 88) {
            const long long idx = src + hh * 88 + cc;
            rf[k] = p.in_dt == DT_F16 ? __half2float(reinterpret_cast<const __half*>(p.video)[idx])
                                      : reinterpret_cast<const float*>(p.video)[idx];
          }
        }
      }
    };
    // patch rows of this thread: row slot rs -> (stem row of the band, stem column, row of the tile)
    int b_ry[2], b_x[2], b_row[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rs = bt + 128 * k;
      if (rs < 110) { b_ry[k] = rs / 22; b_x[k] = rs - b_ry[k] * 22; b_row[k] = rs; }
      else { const int u = rs - 110; b_ry[k] = u / 23; b_x[k] = 21 + u - b_ry[k] * 23; b_row[k] = H0_ROWS + u; }
    }
    const bool second = bt + 128 < NBUILD;
    pdl_wait();                                            // the video may be produced by the previous kernel
    if (valid) prefetch(im, f);
    int n = 0;
    while (valid) {
      // staged input rows of frame f -> smem
      if (fast) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = bt + 128 * k;
          if (i < 165) {
            const int r = i / 11, c8 = i - r * 11;
            __nv_bfloat16* d = sIn + r * SIN_PITCH + 3 + c8 * 8;
            const __nv_bfloat16* sv = reinterpret_cast<const __nv_bfloat16*>(&rv[k]);
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2) d[e2] = sv[e2];
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const int i = bt + 128 * k;
          if (i < SIN_ROWS * 88) {
            const int r = i / 88, cc = i - r * 88;
            sIn[r * SIN_PITCH + 3 + cc] = __float2bfloat16_rn(rf[k]);
          }
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      // next frame of the walk (possibly the first of the next item) into registers while this one is built
      Item nim = im;
      int nf = f + 1, nitem = item;
      bool nvalid = true;
      if (nf > im.f_last) {
        nitem = item + gridDim.x;
        nvalid = nitem < p.num_items;
        if (nvalid) { nim = decode_item(p, nitem); nf = nim.f_first; }
      }
      if (nvalid) prefetch(nim, nf);
      const long long cb0 = clock64();
      const int slot = n % RING;
      wait_acc(&slot_free[slot], ((n / RING) & 1) ^ 1, p.dbg, st0);
      uint8_t* tile = ring + slot * SLOT_BYTES;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 0 || second) {
          uint8_t* rowp = tile + b_row[k] * 128;
          const int sw = b_row[k] & 7;
          const uint32_t* s0 = reinterpret_cast<const uint32_t*>(sIn + (2 * b_ry[k]) * SIN_PITCH + 2 * b_x[k]);
#pragma unroll
          for (int kh = 0; kh < 7; ++kh) {
            const uint32_t* s = s0 + kh * (SIN_PITCH / 2);
            *reinterpret_cast<uint4*>(rowp + ((kh ^ sw) << 4)) = make_uint4(s[0], s[1], s[2], s[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&patch_full[slot]);
      asm volatile("bar.sync 2, 128;" ::: "memory");       // sIn is rewritten next
      st1 += clock64() - cb0;
      ++n;
      im = nim; f = nf; item = nitem; valid = nvalid;
    }
    if (p.dbg != nullptr && bt == 0) { p.dbg[blockIdx.x * 8 + 6] = st0; p.dbg[blockIdx.x * 8 + 7] = st1; p.dbg[blockIdx.x * 8 + 1] = n; }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8] = clock64() - k_start;
}

}  // namespace

int stem_fused_plan(const void* w_packed, StemFusedPlan* plan) {
  AVH_CHECK(w_packed != nullptr && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, "stem weights must be 16-byte aligned");
  plan->w = w_packed;
  return 0;
}

int stem_fused_launch(const StemFusedPlan& plan, const void* video, int in_dt, int T, int b0, int nb, const float* scale,
                      const float* bias, const float* slope, void* out, cudaStream_t stream) {
  if (nb <= 0 || T <= 0) return 0;
  AVH_CHECK(in_dt == DT_BF16 || in_dt == DT_F16 || in_dt == DT_F32, "unsupported video dtype");
  AVH_CHECK(in_dt != DT_BF16 || (reinterpret_cast<uintptr_t>(video) & 15) == 0, "video must be 16-byte aligned");
  const int sms = device_sm_count();
  // time segments (even lengths: frames are processed in pairs): enough items to fill whole waves; every segment
  // re-builds 2 + 2 boundary frames
  int best_seg = 1;
  double best = 1e30;
  for (int ns = 1; ns <= T && ns <= 32; ++ns) {
    int len = (T + ns - 1) / ns;
    len += len & 1;
    const int real = (T + len - 1) / len;               // segments that actually hold frames
    if (real != ns) continue;
    const long long items = (long long)nb * 11 * ns;
    const long long rounds = (items + sms - 1) / sms;
    const double cost = (double)rounds * (len + 3.0);
    if (cost < best) { best = cost; best_seg = ns; }
  }
  StemParams p;
  p.video = video; p.in_dt = in_dt; p.T = T; p.b0 = b0; p.nb = nb;
  p.items = nullptr; p.cu = nullptr; p.n_items = nullptr;
  p.nseg = best_seg;
  p.seglen = (T + best_seg - 1) / best_seg;
  p.seglen += p.seglen & 1;
  p.num_items = nb * 11 * best_seg;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.w = reinterpret_cast<const __nv_bfloat16*>(plan.w);
  p.scale = scale; p.bias = bias; p.slope = slope;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(stem_fused_kernel), (int)SMEM_BYTES)) return 1;
  const int grid = p.num_items < sms ? p.num_items : sms;
  p.dbg = nullptr;
  static int dbg_env = -1;
  if (dbg_env < 0) { const char* ev = std::getenv("AVH_STEM_DBG"); dbg_env = ev != nullptr ? std::atoi(ev) : 0; }
  if (dbg_env > 0) {
    // debug only: stall accounting of one launch (synchronises the stream), printed to stderr
    --dbg_env;
    unsigned long long* d = nullptr;
    AVH_CUDA_OK(cudaMalloc(&d, (size_t)grid * 64));
    AVH_CUDA_OK(cudaMemset(d, 0, (size_t)grid * 64));
    p.dbg = d;
    AVH_CUDA_OK(launch_pdl(stem_fused_kernel, dim3(grid), dim3(NT), SMEM_BYTES, stream, p));
    AVH_CUDA_OK(cudaStreamSynchronize(stream));
    std::vector<unsigned long long> hst((size_t)grid * 8);
    AVH_CUDA_OK(cudaMemcpy(hst.data(), d, (size_t)grid * 64, cudaMemcpyDeviceToHost));
    cudaFree(d);
    static const char* names[8] = {"total", "frames_built", "mma_wait_patch", "mma_wait_d_empty", "epi_wait_d0_full",
                                   "epi_wait_d1_full", "build_wait_slot", "build_busy"};
    for (int c : {0, grid / 2, grid - 1}) {
      std::fprintf(stderr, "[stem_fused dbg] cta %d (nseg %d seglen %d items %d):", c, p.nseg, p.seglen, p.num_items);
      for (int i = 0; i < 8; ++i) std::fprintf(stderr, " %s=%llu", names[i], hst[(size_t)c * 8 + i]);
      std::fprintf(stderr, "\n");
    }
    count_launch(1);
    return 0;
  }
  AVH_CUDA_OK(launch_pdl(stem_fused_kernel, dim3(grid), dim3(NT), SMEM_BYTES, stream, p));
  count_launch(1);
  return 0;
}

// ---- ragged batches: the (clip, band, time segment) work list of one call, built on the host from the clip lengths.
// Segments are even-length (frames are processed in pairs); every segment re-builds 2 + 2 boundary frames, so the
// segment length trades boundary work against whole waves over the SMs.
int stem_fused_ragged_items(const int* lengths, int n_clips, int sms, int4* items, int max_items) {
  long long total = 0;
  int tmax = 0;
  for (int b = 0; b < n_clips; ++b) { total += lengths[b]; tmax = lengths[b] > tmax ? lengths[b] : tmax; }
  int best_len = 2;
  double best = 1e30;
  for (int len = 2; len <= ((tmax + 1) & ~1) || len == 2; len += 2) {
    long long segs = 0;
    for (int b = 0; b < n_clips; ++b) segs += (lengths[b] + len - 1) / len;
    const long long n = segs * 11;
    if (n > max_items) continue;
    const long long rounds = (n + sms - 1) / sms;
    const double cost = (double)rounds * (len + 3.0);
    if (cost < best) { best = cost; best_len = len; }
  }
  int n = 0;
  for (int b = 0; b < n_clips; ++b)
    for (int j = 0; j < 11; ++j)
      for (int t0 = 0; t0 < lengths[b]; t0 += best_len) {
        if (n >= max_items) return -1;
        const int t1 = t0 + best_len < lengths[b] ? t0 + best_len : lengths[b];
        items[n++] = make_int4(b, j, t0, t1);
      }
  return n;
}

int stem_fused_launch_ragged(const StemFusedPlan& plan, const void* video, int in_dt, int Tpitch, const int4* items,
                             const int* n_items, const int* cu, const float* scale, const float* bias, const float* slope,
                             void* out, cudaStream_t stream) {
  AVH_CHECK(in_dt == DT_BF16 || in_dt == DT_F16 || in_dt == DT_F32, "unsupported video dtype");
  AVH_CHECK(in_dt != DT_BF16 || (reinterpret_cast<uintptr_t>(video) & 15) == 0, "video must be 16-byte aligned");
  StemParams p;
  p.video = video; p.in_dt = in_dt; p.T = Tpitch; p.b0 = 0; p.nb = 0;
  p.nseg = 1; p.seglen = 2; p.num_items = 0;
  p.items = items; p.cu = cu; p.n_items = n_items;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.w = reinterpret_cast<const __nv_bfloat16*>(plan.w);
  p.scale = scale; p.bias = bias; p.slope = slope;
  p.dbg = nullptr;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(stem_fused_kernel), (int)SMEM_BYTES)) return 1;
  AVH_CUDA_OK(launch_pdl(stem_fused_kernel, dim3(device_sm_count()), dim3(NT), SMEM_BYTES, stream, p));
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace avh
