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
                 __nv_bfloat16* __restrict__ out, int T, int D) {
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
int launch_att_t(const void* qkv, const unsigned char* kpm, void* out, int B, int T, int D, int H, cudaStream_t stream) {
  constexpr int BQ = NW * 16;
  const size_t smem = (size_t)(BQ + 2 * NRES * BKV) * PITCH * 2 + BKV * 4;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(attention_kernel<NW, BKV, MINB, NRES>), (int)smem)) return 1;
  dim3 grid((T + BQ - 1) / BQ, H, B);
  AVH_CUDA_OK(launch_pdl(attention_kernel<NW, BKV, MINB, NRES>, grid, dim3(NW * 32), smem, stream,
                         reinterpret_cast<const __nv_bfloat16*>(qkv), kpm, reinterpret_cast<__nv_bfloat16*>(out), T, D));
  return 0;
}

// fp32 reference-precision attention for the split-precision (fp32) mode: one warp per query row.
__global__ void attention_f32_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ kpm,
                                     float* __restrict__ out, int T, int D, int H) {
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

int launch_attention(const void* qkv, const unsigned char* kpm, void* out, int B, int T, int D, int H, int fp32,
                     cudaStream_t stream) {
  AVH_CHECK(D == H * HD, "attention kernel requires head_dim 64");
  if (fp32) {
    const int nw = 4;
    dim3 grid((T + nw - 1) / nw, H, B);
    const size_t smem = (size_t)nw * T * sizeof(float);
    AVH_CHECK(smem <= 48 * 1024, "sequence too long for the fp32 attention kernel");
    AVH_CUDA_OK(launch_pdl(attention_f32_kernel, grid, dim3(nw * 32), smem, stream,
                           reinterpret_cast<const float*>(qkv), kpm, reinterpret_cast<float*>(out), T, D, H));
  } else {
    // tile shapes: whole-clip tiles for short clips (T <= 160: one CTA per (batch, head)), else 128 x 128
    int rc;
    // (a 160-key chunk needs 80 score registers per thread and drops to one CTA per SM: 256 heads on 148 SMs
    //  = two waves; 80-key chunks with the online softmax fit two CTAs per SM and finish in one wave)
    if (T <= 64) rc = launch_att_t<4, 64, 2, 1>(qkv, kpm, out, B, T, D, H, stream);
    else if (T <= 96) rc = launch_att_t<6, 96, 2, 1>(qkv, kpm, out, B, T, D, H, stream);
    else if (T <= 160) rc = launch_att_t<10, 80, 2, 2>(qkv, kpm, out, B, T, D, H, stream);
    else rc = launch_att_t<8, 128, 2, 1>(qkv, kpm, out, B, T, D, H, stream);
    if (rc) return rc;
  }
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace avh
