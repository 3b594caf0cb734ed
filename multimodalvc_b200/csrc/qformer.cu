// Kernels of the Q-Former block that consumes the AV-HuBERT features in MMS-LLaMA (SURVEY 8(f) rank 3;
// src/model.py:584-619 -> src/sub_model/Qformer.py:678-968): everything that is not a Linear layer (those run on the
// tcgen05 GEMM).  A handful of queries (<= 120) per clip attend to themselves and to up to 1200 resized AV frames; the
// attention core is 1-2 % of the block's FLOPs (the K / V projections of the frames dominate), so it is a plain fp32
// CUDA-core kernel: exact softmax, any query / key length, separate Q / K / V pointers (cross-attention), head dim 64.
//
//  * attention_x_kernel: one CTA per (clip, head, 32 queries); warp w owns keys w, w+8, ... of every 64-key tile, lane
//    = query: the query row and the output accumulator live in registers, K / V tiles in shared memory are read as
//    warp-wide broadcasts; running (max, sum, acc) per thread, the 8 partial states of a query are merged through
//    shared memory at the end.  Keys flagged in key_pad are skipped — the reference adds -10000 to their scores
//    (Qformer.py:797-801), which exp() turns into exactly 0 in fp32 whenever one key is valid.
//  * rows_broadcast_kernel: hidden states of every clip start as the same LayerNorm(query_tokens) rows
//    (BertEmbeddings with query_embeds only, Qformer.py:98-110).
#include "attn_tile.cuh"
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace avh {
namespace {

constexpr int AX_QT = 32;      // queries per CTA
constexpr int AX_KT = 64;      // keys per shared-memory tile
constexpr int AX_HD = 64;      // head dim

__device__ __forceinline__ float ldx(const void* p, int dt, long long i) {
  if (dt == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  if (dt == DT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}

__global__ void __launch_bounds__(256)
attention_x_kernel(const void* __restrict__ Q, long long ldq, const void* __restrict__ K, long long ldk,
                   const void* __restrict__ V, long long ldv, int dt, const unsigned char* __restrict__ key_pad,
                   void* __restrict__ O, long long ldo, int o_dt, int Lq, int Lk, float scale, float* __restrict__ lse) {
  __shared__ __align__(16) float Ks[AX_KT][AX_HD];
  __shared__ __align__(16) float Vs[AX_KT][AX_HD];
  __shared__ unsigned char dead[AX_KT];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AX_QT;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int qi = q0 + lane;
  const bool q_ok = qi < Lq;
  float q[AX_HD], acc[AX_HD];
  {   // the CTA's 32 query rows: coalesced load into shared memory (padded rows), then one row per thread
    float* stage = &Ks[0][0];                                   // [32][65] fits the 64 x 64 key tile
    attn_load_tile<AX_QT, AX_HD + 1, 256>(Q, dt, ldq, (long long)b * Lq, q0, Lq, h * AX_HD, stage);
    __syncthreads();
#pragma unroll
    for (int d = 0; d < AX_HD; ++d) {
      q[d] = stage[lane * (AX_HD + 1) + d] * scale;
      acc[d] = 0.f;
    }
  }
  (void)q_ok;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += AX_KT) {
    __syncthreads();
    attn_load_tile<AX_KT, AX_HD, 256>(K, dt, ldk, (long long)b * Lk, k0, Lk, h * AX_HD, &Ks[0][0]);
    attn_load_tile<AX_KT, AX_HD, 256>(V, dt, ldv, (long long)b * Lk, k0, Lk, h * AX_HD, &Vs[0][0]);
    if (threadIdx.x < AX_KT) {
      const int key = k0 + threadIdx.x;
      dead[threadIdx.x] = (key >= Lk || (key_pad != nullptr && key_pad[(long long)b * Lk + key] != 0)) ? 1 : 0;
    }
    __syncthreads();
    float s[AX_KT / 8];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < AX_KT / 8; ++j) {
      const int r = w + 8 * j;
      float a = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(Ks[r]);
#pragma unroll
      for (int d4 = 0; d4 < AX_HD / 4; ++d4) {
        const float4 kv = kr[d4];
        a = fmaf(q[4 * d4], kv.x, a); a = fmaf(q[4 * d4 + 1], kv.y, a);
        a = fmaf(q[4 * d4 + 2], kv.z, a); a = fmaf(q[4 * d4 + 3], kv.w, a);
      }
      s[j] = dead[r] ? -INFINITY : a;
      tmax = fmaxf(tmax, s[j]);
    }
    if (tmax == -INFINITY) continue;             // no live key of this tile belongs to the warp (uniform per warp)
    const float m_new = fmaxf(m, tmax);
    const float resc = __expf(m - m_new);        // m = -inf on the first live tile: exp(-inf) = 0
    l *= resc;
#pragma unroll
    for (int d = 0; d < AX_HD; ++d) acc[d] *= resc;
    m = m_new;
#pragma unroll
    for (int j = 0; j < AX_KT / 8; ++j) {
      if (s[j] == -INFINITY) continue;
      const int r = w + 8 * j;
      const float p = __expf(s[j] - m);
      l += p;
      const float4* vr = reinterpret_cast<const float4*>(Vs[r]);
#pragma unroll
      for (int d4 = 0; d4 < AX_HD / 4; ++d4) {
        const float4 vv = vr[d4];
        acc[4 * d4] = fmaf(p, vv.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(p, vv.y, acc[4 * d4 + 1]);
        acc[4 * d4 + 2] = fmaf(p, vv.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(p, vv.w, acc[4 * d4 + 3]);
      }
    }
  }
  // merge the 8 per-warp states of every query: warp 0 deposits, warps 1..7 fold theirs in, in warp order
  __syncthreads();
  float* Ms = &Ks[0][0];                 // [32] running max, then [32] sums
  float* Ls = Ms + 32;
  float (*As)[AX_HD + 1] = reinterpret_cast<float (*)[AX_HD + 1]>(&Vs[0][0]);     // [32][65] accumulators (8.3 KB of 16)
  for (int turn = 0; turn < 8; ++turn) {
    if (w == turn) {
      if (turn == 0) {
        Ms[lane] = m; Ls[lane] = l;
#pragma unroll
        for (int d = 0; d < AX_HD; ++d) As[lane][d] = acc[d];
      } else if (m != -INFINITY) {
        const float mo = Ms[lane];
        const float mn = fmaxf(mo, m);
        const float so = mo == -INFINITY ? 0.f : __expf(mo - mn), sn = __expf(m - mn);
        Ms[lane] = mn;
        Ls[lane] = Ls[lane] * so + l * sn;
#pragma unroll
        for (int d = 0; d < AX_HD; ++d) As[lane][d] = As[lane][d] * so + acc[d] * sn;
      }
    }
    __syncthreads();
  }
  // log-sum-exp of every query row for the backward pass ([B, H, Lq]; +inf for a row without a live key: P = 0)
  if (lse != nullptr && threadIdx.x < AX_QT && q0 + threadIdx.x < Lq)
    lse[((long long)b * gridDim.y + h) * Lq + q0 + threadIdx.x] =
        Ls[threadIdx.x] > 0.f ? Ms[threadIdx.x] + __logf(Ls[threadIdx.x]) : INFINITY;
  for (int e = threadIdx.x; e < AX_QT * AX_HD; e += 256) {
    const int r = e / AX_HD, d = e % AX_HD;
    if (q0 + r >= Lq) continue;
    const float den = Ls[r];
    const float v = den > 0.f ? As[r][d] / den : 0.f;
    const long long o = ((long long)b * Lq + q0 + r) * ldo + h * AX_HD + d;
    if (o_dt == DT_BF16) reinterpret_cast<__nv_bfloat16*>(O)[o] = __float2bfloat16_rn(v);
    else if (o_dt == DT_F16) reinterpret_cast<__half*>(O)[o] = __float2half_rn(v);
    else reinterpret_cast<float*>(O)[o] = v;
  }
}

// dst [B*L, C] (fp32) = src [L, C] repeated for every clip
__global__ void __launch_bounds__(256)
rows_broadcast_kernel(const float* __restrict__ src, float* __restrict__ dst, long long per_clip, int B) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= per_clip) return;
  const float v = src[i];
  for (int b = 0; b < B; ++b) dst[(long long)b * per_clip + i] = v;
}

}  // namespace

int launch_attention_x(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, int dt,
                       const unsigned char* key_pad, void* O, long long ldo, int o_dt, int B, int H, int Lq, int Lk,
                       float scale, cudaStream_t stream, float* lse) {
  AVH_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0, "bad attention shape");
  AVH_CHECK(B <= 65535 && H <= 65535, "batch / head count exceeds the grid limits");
  dim3 grid((Lq + AX_QT - 1) / AX_QT, H, B);
  attention_x_kernel<<<grid, 256, 0, stream>>>(Q, ldq, K, ldk, V, ldv, dt, key_pad, O, ldo, o_dt, Lq, Lk, scale, lse);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

int launch_rows_broadcast(const float* src, float* dst, long long per_clip, int B, cudaStream_t stream) {
  if (per_clip <= 0 || B <= 0) return 0;
  rows_broadcast_kernel<<<(unsigned)((per_clip + 255) / 256), 256, 0, stream>>>(src, dst, per_clip, B);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace avh
