// Flash-style multi-head self-attention with key-padding mask, head_dim 64, bf16 in / fp32 softmax.
//
// Replaces the attention core inside torch's F.multi_head_attention_forward as called by
// fairseq/fairseq/modules/multihead_attention.py:170-192 (baddbmm + softmax + bmm + head-averaged
// [B,T,T] weights that the encoder discards, wav2vec2.py:889).  Nothing of size T x T is materialised.
// Input is the fused QKV projection [B*T, 3*D] (q pre-scaled by head_dim^-0.5 at weight-fold time),
// output is the per-head context [B*T, D] ready for out_proj.
//
// Warp-level mma.sync (m16n8k16) tiles: attention is ~1.2 % of the path's FLOPs at T=150 and the (b,h) problems
// are tiny (150x150x64), so one CTA per head with Q/K/V resident in smem beats a tcgen05 pipeline's fixed costs.
#include "common.cuh"
#include "kernels.h"

namespace avh {
namespace {

constexpr int HD = 64;       // head dim
constexpr int PITCH = 72;    // smem row pitch in bf16 (144 B) -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;      // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous copy of a [ROWS x 64] bf16 tile (rows row0.., global row stride ld) into smem [ROWS][PITCH];
// rows >= nrows are zero-filled
template <int ROWS, int NT>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int row0,
                                                int nrows) {
  for (int i = threadIdx.x; i < ROWS * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const bool ok = row0 + r < nrows;
    cp_async16(dst + r * PITCH + c * 8, src + (long long)(ok ? row0 + r : 0) * ld + c * 8, ok);
  }
}

// One CTA = NW warps x 16 query rows of one (batch, head); keys/values go through smem in chunks of BKV keys with an
// online softmax.  NRES = 1: chunks stream through one buffer (any T).  NRES = 2 (T <= 2 * BKV, the 6 s clip): both
// chunks are requested up front into their own buffers, so the second chunk's L2 latency hides behind the first
// chunk's math instead of sitting between two __syncthreads.
template <int NW, int BKV, int MINB, int NRES>
__global__ void __launch_bounds__(NW * 32, MINB)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ kpm,
                 __nv_bfloat16* __restrict__ out, int T, int D, float* __restrict__ lse) {
  constexpr int BQ = NW * 16;
  constexpr int NT = NW * 32;
  constexpr int NJ = BKV / 8;        // 8-key score tiles per chunk
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem_att[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sK = sQ + BQ * PITCH;
  __nv_bfloat16* sV = sK + NRES * BKV * PITCH;
  float* sMask = reinterpret_cast<float*>(sV + NRES * BKV * PITCH);

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ld = 3ll * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  const int q0 = qt * BQ;

  load_tile_async<BQ, NT>(sQ, base, ld, q0, T);
  load_tile_async<BKV, NT>(sK, base + D, ld, 0, T);
  load_tile_async<BKV, NT>(sV, base + 2 * D, ld, 0, T);
  cp_async_commit();
  if (NRES == 2) {
    load_tile_async<BKV, NT>(sK + BKV * PITCH, base + D, ld, BKV, T);
    load_tile_async<BKV, NT>(sV + BKV * PITCH, base + 2 * D, ld, BKV, T);
    cp_async_commit();
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  constexpr float LOG2E = 1.4426950408889634f;
  uint32_t qf[4][4];

  for (int k0 = 0; k0 < T; k0 += BKV) {
    if (NRES == 1 && k0 > 0) {
      __syncthreads();   // previous chunk fully consumed
      load_tile_async<BKV, NT>(sK, base + D, ld, k0, T);
      load_tile_async<BKV, NT>(sV, base + 2 * D, ld, k0, T);
      cp_async_commit();
    }
    if (NRES == 2 && k0 > 0) {
      __syncthreads();   // mask of the previous chunk consumed
      sK += BKV * PITCH;
      sV += BKV * PITCH;
    }
    for (int i = threadIdx.x; i < BKV; i += NT) {
      const int k = k0 + i;
      const bool dead = (k >= T) || (kpm != nullptr && kpm[(long long)b * T + k] != 0);
      sMask[i] = dead ? -INFINITY : 0.f;
    }
    if (NRES == 2 && k0 == 0) cp_async_wait_group<1>();
    else cp_async_wait_all();
    __syncthreads();
    if (k0 == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(qf[kk], sQ + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + kk * 16 + 8 * (lane >> 4));
    }

    float s[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {          // 8-key n-tiles
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {     // two pairs of k-steps (dims 0-31, 32-63)
        uint32_t kf[4];
        ldsm_x4(kf, sK + (j * 8 + (lane & 7)) * PITCH + kp * 32 + 8 * (lane >> 3));
        mma_bf16(s[j], qf[2 * kp], kf[0], kf[1]);
        mma_bf16(s[j], qf[2 * kp + 1], kf[2], kf[3]);
      }
    }
    // mask + online softmax (rows: lane/4 and lane/4+8; this thread's keys: j*8 + 2*(lane%4) + {0,1})
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float2 mk = *reinterpret_cast<const float2*>(sMask + j * 8 + 2 * (lane & 3));
      s[j][0] += mk.x; s[j][1] += mk.y; s[j][2] += mk.x; s[j][3] += mk.y;
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float scale[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      scale[r] = exp2f((m_run[r] - msafe) * LOG2E);     // m_run=-inf -> 0
      m_run[r] = mnew[r];
      mnew[r] = msafe * LOG2E;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], LOG2E, -mnew[0]));
      s[j][1] = exp2f(fmaf(s[j][1], LOG2E, -mnew[0]));
      s[j][2] = exp2f(fmaf(s[j][2], LOG2E, -mnew[1]));
      s[j][3] = exp2f(fmaf(s[j][3], LOG2E, -mnew[1]));
      rs[0] += s[j][0] + s[j][1];
      rs[1] += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= scale[0]; o[j][1] *= scale[0]; o[j][2] *= scale[1]; o[j][3] *= scale[1];
    }
    l_run[0] = l_run[0] * scale[0] + rs[0];
    l_run[1] = l_run[1] * scale[1] + rs[1];
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {       // 16 keys per step
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {     // pairs of 8-wide dim tiles
        uint32_t vf[4];
        ldsm_x4_t(vf, sV + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + dp * 16 + 8 * (lane >> 4));
        mma_bf16(o[2 * dp], pf, vf[0], vf[1]);
        mma_bf16(o[2 * dp + 1], pf, vf[2], vf[3]);
      }
    }
  }
  // finalise: divide by the row sums (full sum across the 4 lanes of a row)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int row0 = q0 + warp * 16 + (lane >> 2);
  if (lse != nullptr && (lane & 3) == 0) {     // log-sum-exp of every query row (training: the backward recomputes P from it)
    float* lr = lse + ((long long)b * gridDim.y + h) * T;
    if (row0 < T) lr[row0] = l_run[0] > 0.f ? m_run[0] + logf(l_run[0]) : INFINITY;
    if (row0 + 8 < T) lr[row0 + 8] = l_run[1] > 0.f ? m_run[1] + logf(l_run[1]) : INFINITY;
  }
  __nv_bfloat16* ob = out + (long long)b * T * D + h * HD + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (row0 < T)
      *reinterpret_cast<uint32_t*>(ob + (long long)row0 * D + j * 8) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    if (row0 + 8 < T)
      *reinterpret_cast<uint32_t*>(ob + (long long)(row0 + 8) * D + j * 8) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
}

template <int NW, int BKV, int MINB, int NRES>
int launch_att_t(const void* qkv, const unsigned char* kpm, void* out, int B, int T, int D, int H, cudaStream_t stream,
                 float* lse) {
  constexpr int BQ = NW * 16;
  const size_t smem = (size_t)(BQ + 2 * NRES * BKV) * PITCH * 2 + BKV * 4;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(attention_kernel<NW, BKV, MINB, NRES>), (int)smem)) return 1;
  dim3 grid((T + BQ - 1) / BQ, H, B);
  AVH_CUDA_OK(launch_pdl(attention_kernel<NW, BKV, MINB, NRES>, grid, dim3(NW * 32), smem, stream,
                         reinterpret_cast<const __nv_bfloat16*>(qkv), kpm, reinterpret_cast<__nv_bfloat16*>(out), T, D, lse));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- backward (bf16 mode)
// Same tiles and fragments as the forward.  P is recomputed from the saved log-sum-exp of every query row:
// P_ij = exp(q_i . k_j - lse_i), D_i = dO_i . O_i, dS = P o (dO V^T - D), dQ = dS K, dK = dS^T Q, dV = P^T dO.
// Two kernels, no atomics: one accumulates dQ over the key chunks (CTA = NW x 16 queries), one accumulates dK / dV over
// the query chunks (CTA = NW x 16 keys, the scores computed TRANSPOSED: K Q^T, V dO^T, so that the keys are the MMA rows).
// P and dS go through bf16 for the second product, exactly as P does in the forward.
template <int NW, int BKV>
__global__ void __launch_bounds__(NW * 32)
attention_bwd_dq_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dO,
                           const __nv_bfloat16* __restrict__ O, const float* __restrict__ lse,
                           const unsigned char* __restrict__ kpm, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ Dbuf,
                           int T, int D) {
  constexpr int BQ = NW * 16, NT = NW * 32, NJ = BKV / 8;
  extern __shared__ __align__(16) uint8_t smem_att[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sG = sQ + BQ * PITCH;
  __nv_bfloat16* sK = sG + BQ * PITCH;
  __nv_bfloat16* sV = sK + BKV * PITCH;
  float* sMask = reinterpret_cast<float*>(sV + BKV * PITCH);
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ld = 3ll * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  const __nv_bfloat16* gbase = dO + (long long)b * T * D + h * HD;
  const __nv_bfloat16* obase = O + (long long)b * T * D + h * HD;
  const int q0 = qt * BQ;
  load_tile_async<BQ, NT>(sQ, base, ld, q0, T);
  load_tile_async<BQ, NT>(sG, gbase, D, q0, T);
  load_tile_async<BKV, NT>(sK, base + D, ld, 0, T);
  load_tile_async<BKV, NT>(sV, base + 2 * D, ld, 0, T);
  cp_async_commit();
  // this thread's two query rows, their log-sum-exp and D = dO . O (each of a row's 4 lanes sums 16 channels)
  const int ra = q0 + warp * 16 + (lane >> 2), rb = ra + 8;
  float Dr[2] = {0.f, 0.f}, L2[2];
  {
    const int rows[2] = {ra, rb};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float acc = 0.f;
      if (rows[r] < T) {
        const __nv_bfloat16* gp = gbase + (long long)rows[r] * D + (lane & 3) * 16;
        const __nv_bfloat16* op = obase + (long long)rows[r] * D + (lane & 3) * 16;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc = fmaf(__bfloat162float(gp[d]), __bfloat162float(op[d]), acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      Dr[r] = acc;
      const float l = rows[r] < T ? lse[((long long)b * H + h) * T + rows[r]] : INFINITY;
      L2[r] = l * 1.4426950408889634f;
      if ((lane & 3) == 0 && rows[r] < T) Dbuf[((long long)b * H + h) * T + rows[r]] = acc;
    }
  }
  constexpr float LOG2E = 1.4426950408889634f;
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  uint32_t qf[4][4], gf[4][4];
  for (int k0 = 0; k0 < T; k0 += BKV) {
    if (k0 > 0) {
      __syncthreads();
      load_tile_async<BKV, NT>(sK, base + D, ld, k0, T);
      load_tile_async<BKV, NT>(sV, base + 2 * D, ld, k0, T);
      cp_async_commit();
    }
    for (int i = threadIdx.x; i < BKV; i += NT) {
      const int k = k0 + i;
      sMask[i] = ((k >= T) || (kpm != nullptr && kpm[(long long)b * T + k] != 0)) ? -INFINITY : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    if (k0 == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        ldsm_x4(qf[kk], sQ + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + kk * 16 + 8 * (lane >> 4));
        ldsm_x4(gf[kk], sG + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + kk * 16 + 8 * (lane >> 4));
      }
    }
    float s[NJ][4], dp[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, sK + (j * 8 + (lane & 7)) * PITCH + kp * 32 + 8 * (lane >> 3));
        ldsm_x4(vf, sV + (j * 8 + (lane & 7)) * PITCH + kp * 32 + 8 * (lane >> 3));
        mma_bf16(s[j], qf[2 * kp], kf[0], kf[1]);
        mma_bf16(s[j], qf[2 * kp + 1], kf[2], kf[3]);
        mma_bf16(dp[j], gf[2 * kp], vf[0], vf[1]);
        mma_bf16(dp[j], gf[2 * kp + 1], vf[2], vf[3]);
      }
    }
    // dS = P o (dP - D) in the score fragment layout (rows ra / rb, keys j*8 + 2*(lane%4) + {0,1})
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float2 mk = *reinterpret_cast<const float2*>(sMask + j * 8 + 2 * (lane & 3));
      s[j][0] = exp2f(fmaf(s[j][0] + mk.x, LOG2E, -L2[0])) * (dp[j][0] - Dr[0]);
      s[j][1] = exp2f(fmaf(s[j][1] + mk.y, LOG2E, -L2[0])) * (dp[j][1] - Dr[0]);
      s[j][2] = exp2f(fmaf(s[j][2] + mk.x, LOG2E, -L2[1])) * (dp[j][2] - Dr[1]);
      s[j][3] = exp2f(fmaf(s[j][3] + mk.y, LOG2E, -L2[1])) * (dp[j][3] - Dr[1]);
    }
    // dQ += dS K  (K as the [keys, dims] B operand: transposed ldmatrix, as V in the forward)
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t kt[4];
        ldsm_x4_t(kt, sK + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + dpair * 16 + 8 * (lane >> 4));
        mma_bf16(dq[2 * dpair], pf, kt[0], kt[1]);
        mma_bf16(dq[2 * dpair + 1], pf, kt[2], kt[3]);
      }
    }
  }
  __nv_bfloat16* ob = dqkv + (long long)b * T * ld + h * HD + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (ra < T) *reinterpret_cast<uint32_t*>(ob + (long long)ra * ld + j * 8) = pack_bf16(dq[j][0], dq[j][1]);
    if (rb < T) *reinterpret_cast<uint32_t*>(ob + (long long)rb * ld + j * 8) = pack_bf16(dq[j][2], dq[j][3]);
  }
}

template <int NW, int BQC>
__global__ void __launch_bounds__(NW * 32)
attention_bwd_dkv_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dO,
                            const float* __restrict__ lse, const float* __restrict__ Dbuf,
                            const unsigned char* __restrict__ kpm, __nv_bfloat16* __restrict__ dqkv, int T, int D) {
  constexpr int BKB = NW * 16, NT = NW * 32, NJ = BQC / 8;
  extern __shared__ __align__(16) uint8_t smem_att[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sV = sK + BKB * PITCH;
  __nv_bfloat16* sQ = sV + BKB * PITCH;
  __nv_bfloat16* sG = sQ + BQC * PITCH;
  float* sL = reinterpret_cast<float*>(sG + BQC * PITCH);      // log2(e) * lse of the chunk's queries (+inf past the clip)
  float* sD = sL + BQC;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ld = 3ll * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  const __nv_bfloat16* gbase = dO + (long long)b * T * D + h * HD;
  const int j0 = kt * BKB;
  load_tile_async<BKB, NT>(sK, base + D, ld, j0, T);
  load_tile_async<BKB, NT>(sV, base + 2 * D, ld, j0, T);
  load_tile_async<BQC, NT>(sQ, base, ld, 0, T);
  load_tile_async<BQC, NT>(sG, gbase, D, 0, T);
  cp_async_commit();
  const int ka = j0 + warp * 16 + (lane >> 2), kb = ka + 8;      // this thread's two key rows
  const bool live_a = ka < T && !(kpm != nullptr && kpm[(long long)b * T + ka] != 0);
  const bool live_b = kb < T && !(kpm != nullptr && kpm[(long long)b * T + kb] != 0);
  constexpr float LOG2E = 1.4426950408889634f;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  uint32_t kf[4][4], vf[4][4];
  for (int i0 = 0; i0 < T; i0 += BQC) {
    if (i0 > 0) {
      __syncthreads();
      load_tile_async<BQC, NT>(sQ, base, ld, i0, T);
      load_tile_async<BQC, NT>(sG, gbase, D, i0, T);
      cp_async_commit();
    }
    for (int i = threadIdx.x; i < BQC; i += NT) {
      const int q = i0 + i;
      sL[i] = q < T ? lse[((long long)b * H + h) * T + q] * LOG2E : INFINITY;
      sD[i] = q < T ? Dbuf[((long long)b * H + h) * T + q] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    if (i0 == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        ldsm_x4(kf[kk], sK + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + kk * 16 + 8 * (lane >> 4));
        ldsm_x4(vf[kk], sV + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + kk * 16 + 8 * (lane >> 4));
      }
    }
    // transposed scores: rows = keys, columns = the chunk's queries
    float st[NJ][4], dpt[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t qb[4], gb[4];
        ldsm_x4(qb, sQ + (j * 8 + (lane & 7)) * PITCH + kp * 32 + 8 * (lane >> 3));
        ldsm_x4(gb, sG + (j * 8 + (lane & 7)) * PITCH + kp * 32 + 8 * (lane >> 3));
        mma_bf16(st[j], kf[2 * kp], qb[0], qb[1]);
        mma_bf16(st[j], kf[2 * kp + 1], qb[2], qb[3]);
        mma_bf16(dpt[j], vf[2 * kp], gb[0], gb[1]);
        mma_bf16(dpt[j], vf[2 * kp + 1], gb[2], gb[3]);
      }
    }
    // P^T and dS^T (this thread: key rows ka / kb, query columns j*8 + 2*(lane%4) + {0,1})
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float2 l2 = *reinterpret_cast<const float2*>(sL + j * 8 + 2 * (lane & 3));
      const float2 dd = *reinterpret_cast<const float2*>(sD + j * 8 + 2 * (lane & 3));
      const float p0 = live_a ? exp2f(fmaf(st[j][0], LOG2E, -l2.x)) : 0.f;
      const float p1 = live_a ? exp2f(fmaf(st[j][1], LOG2E, -l2.y)) : 0.f;
      const float p2 = live_b ? exp2f(fmaf(st[j][2], LOG2E, -l2.x)) : 0.f;
      const float p3 = live_b ? exp2f(fmaf(st[j][3], LOG2E, -l2.y)) : 0.f;
      st[j][0] = p0; st[j][1] = p1; st[j][2] = p2; st[j][3] = p3;
      dpt[j][0] = p0 * (dpt[j][0] - dd.x); dpt[j][1] = p1 * (dpt[j][1] - dd.y);
      dpt[j][2] = p2 * (dpt[j][2] - dd.x); dpt[j][3] = p3 * (dpt[j][3] - dd.y);
    }
    // dV += P^T dO, dK += dS^T Q  (dO / Q as [queries, dims] B operands: transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < BQC / 16; ++kk) {
      uint32_t pf[4], sf[4];
      pf[0] = pack_bf16(st[2 * kk][0], st[2 * kk][1]);
      pf[1] = pack_bf16(st[2 * kk][2], st[2 * kk][3]);
      pf[2] = pack_bf16(st[2 * kk + 1][0], st[2 * kk + 1][1]);
      pf[3] = pack_bf16(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      sf[0] = pack_bf16(dpt[2 * kk][0], dpt[2 * kk][1]);
      sf[1] = pack_bf16(dpt[2 * kk][2], dpt[2 * kk][3]);
      sf[2] = pack_bf16(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
      sf[3] = pack_bf16(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t gt[4], qt4[4];
        ldsm_x4_t(gt, sG + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + dpair * 16 + 8 * (lane >> 4));
        ldsm_x4_t(qt4, sQ + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + dpair * 16 + 8 * (lane >> 4));
        mma_bf16(dv[2 * dpair], pf, gt[0], gt[1]);
        mma_bf16(dv[2 * dpair + 1], pf, gt[2], gt[3]);
        mma_bf16(dk[2 * dpair], sf, qt4[0], qt4[1]);
        mma_bf16(dk[2 * dpair + 1], sf, qt4[2], qt4[3]);
      }
    }
  }
  __nv_bfloat16* okb = dqkv + (long long)b * T * ld + D + h * HD + 2 * (lane & 3);
  __nv_bfloat16* ovb = dqkv + (long long)b * T * ld + 2 * D + h * HD + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (ka < T) {
      *reinterpret_cast<uint32_t*>(okb + (long long)ka * ld + j * 8) = pack_bf16(dk[j][0], dk[j][1]);
      *reinterpret_cast<uint32_t*>(ovb + (long long)ka * ld + j * 8) = pack_bf16(dv[j][0], dv[j][1]);
    }
    if (kb < T) {
      *reinterpret_cast<uint32_t*>(okb + (long long)kb * ld + j * 8) = pack_bf16(dk[j][2], dk[j][3]);
      *reinterpret_cast<uint32_t*>(ovb + (long long)kb * ld + j * 8) = pack_bf16(dv[j][2], dv[j][3]);
    }
  }
}

// fp32 reference-precision attention for the split-precision (fp32) mode: one warp per query row.
__global__ void attention_f32_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ kpm,
                                     float* __restrict__ out, int T, int D, int H, float* __restrict__ lse) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sp[];                 // [warps][T] probabilities
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * nw + warp;
  if (t >= T) return;
  const long long ld = 3ll * D;
  const float* base = qkv + (long long)b * T * ld + h * HD;
  const float q0 = base[(long long)t * ld + lane], q1 = base[(long long)t * ld + 32 + lane];
  float* p = sp + warp * T;
  float mx = -INFINITY;
  for (int k = 0; k < T; ++k) {
    float d = q0 * base[(long long)k * ld + D + lane] + q1 * base[(long long)k * ld + D + 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (kpm != nullptr && kpm[(long long)b * T + k]) d = -INFINITY;
    if (lane == 0) p[k] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float sum = 0.f;
  for (int k = lane; k < T; k += 32) {
    const float e = (mx == -INFINITY) ? 0.f : expf(p[k] - mx);
    p[k] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  if (lse != nullptr && lane == 0) lse[((long long)b * H + h) * T + t] = sum > 0.f ? mx + logf(sum) : INFINITY;
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < T; ++k) {
    const float w = p[k];
    a0 += w * base[(long long)k * ld + 2 * D + lane];
    a1 += w * base[(long long)k * ld + 2 * D + 32 + lane];
  }
  float* ob = out + ((long long)b * T + t) * D + h * HD;
  ob[lane] = a0 * inv;
  ob[32 + lane] = a1 * inv;
}

}  // namespace

int launch_attention_bwd_tc(const void* qkv, const void* dO, const void* O, const float* lse, const unsigned char* kpm,
                            void* dqkv, float* Dbuf, int B, int T, int D, int H, cudaStream_t stream) {
  AVH_CHECK(D == H * HD, "attention kernel requires head_dim 64");
  constexpr int NW = 4, BC = 64;
  const size_t smem_q = (size_t)(2 * NW * 16 + 2 * BC) * PITCH * 2 + BC * 4;
  const size_t smem_kv = (size_t)(2 * NW * 16 + 2 * BC) * PITCH * 2 + 2 * BC * 4;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(attention_bwd_dq_tc_kernel<NW, BC>), (int)smem_q)) return 1;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(attention_bwd_dkv_tc_kernel<NW, BC>), (int)smem_kv)) return 1;
  dim3 grid((T + NW * 16 - 1) / (NW * 16), H, B);
  attention_bwd_dq_tc_kernel<NW, BC><<<grid, NW * 32, smem_q, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dO),
      reinterpret_cast<const __nv_bfloat16*>(O), lse, kpm, reinterpret_cast<__nv_bfloat16*>(dqkv), Dbuf, T, D);
  AVH_CUDA_OK(cudaGetLastError());
  attention_bwd_dkv_tc_kernel<NW, BC><<<grid, NW * 32, smem_kv, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dO), lse, Dbuf, kpm,
      reinterpret_cast<__nv_bfloat16*>(dqkv), T, D);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return 0;
}

int launch_attention(const void* qkv, const unsigned char* kpm, void* out, int B, int T, int D, int H, int fp32,
                     cudaStream_t stream, float* lse) {
  AVH_CHECK(D == H * HD, "attention kernel requires head_dim 64");
  if (fp32) {
    const int nw = 4;
    dim3 grid((T + nw - 1) / nw, H, B);
    const size_t smem = (size_t)nw * T * sizeof(float);
    AVH_CHECK(smem <= 48 * 1024, "sequence too long for the fp32 attention kernel");
    AVH_CUDA_OK(launch_pdl(attention_f32_kernel, grid, dim3(nw * 32), smem, stream,
                           reinterpret_cast<const float*>(qkv), kpm, reinterpret_cast<float*>(out), T, D, H, lse));
  } else {
    // tile shapes: whole-clip tiles for short clips (T <= 160: one CTA per (batch, head)), else 128 x 128
    int rc;
    // (a 160-key chunk needs 80 score registers per thread and drops to one CTA per SM: 256 heads on 148 SMs
    //  = two waves; 80-key chunks with the online softmax fit two CTAs per SM and finish in one wave)
    if (T <= 64) rc = launch_att_t<4, 64, 2, 1>(qkv, kpm, out, B, T, D, H, stream, lse);
    else if (T <= 96) rc = launch_att_t<6, 96, 2, 1>(qkv, kpm, out, B, T, D, H, stream, lse);
    else if (T <= 160) rc = launch_att_t<10, 80, 2, 2>(qkv, kpm, out, B, T, D, H, stream, lse);
    else rc = launch_att_t<8, 128, 2, 1>(qkv, kpm, out, B, T, D, H, stream, lse);
    if (rc) return rc;
  }
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace avh
