// Training step of the lip ResNet (`feature_grad_mult > 0`): the element-wise / gather / transpose
// kernels of a path in which every convolution is "explicit patches -> one GEMM" on DENSE NHWC maps:
//   forward   col = im2col(x);  raw = col W^T (tcgen05 GEMM);  BatchNorm with batch statistics + PReLU (+ residual)
//   backward  dW = d_raw^T col (GEMM over K = pixels, operands written transposed: im2colT / transposeT);
//             d_col = d_raw W (GEMM);  dx = col2im(d_col) (gather, no atomics)
// avhubert/resnet.py:35-74 (BasicBlock), :131-169 (ResEncoder: Conv3d stem, BatchNorm3d, PReLU, MaxPool3d, trunk, avgpool).
// Two sets of kernels: generic scalar ones (any dtype, hi / mid operand planes: the fp32-faithful mode the parity tests
// run) and the bf16 fast paths (16-byte vectors along the channel axis, 32-bit index arithmetic, 64 x 64 smem-tile
// transposes; ncu record: profiles/r2_full_train_ncu_full.txt).  The inference path keeps its fused implicit-GEMM kernels;
// this file only serves avh_full_train_forward / avh_encoder_backward.
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace avh {
namespace {

__device__ __forceinline__ float ldf(const void* p, int dt, long long i) {
  if (dt == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  if (dt == DT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stf(void* out, int dt, long long i, float v) {
  if (dt == DT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  else if (dt == DT_F16) reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
  else reinterpret_cast<float*>(out)[i] = v;
}
// operand form: plane 0 = bf16 round, plane 1 (planes == 2) = bf16 of the remainder, `ps` elements apart
__device__ __forceinline__ void st_op(__nv_bfloat16* out, long long i, long long ps, int planes, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  out[i] = hi;
  if (planes > 1) out[i + ps] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---- stem patches: video [B,1,T,88,88] -> col [B*T*1936, planes*320]; column dt*64 + kh*8 + kw (kw = 7 and kh = 7 zero:
// the K order of the fused inference stem, whose packed weights the training GEMM shares)
// = video[b, t+dt-2, 2*oy+kh-3, 2*ox+kw-3] (Conv3d(1,64,(5,7,7),stride (1,2,2),pad (2,3,3)), resnet.py:137)
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const void* __restrict__ video, int dt_in, int B, int T, __nv_bfloat16* __restrict__ col, int planes) {
  const long long rows = (long long)B * T * 1936;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;      // (row, dt, j)
  if (idx >= rows * 320) return;
  const long long row = idx / 320;
  const int c = (int)(idx % 320), d = c / 64, j = c % 64;
  float v = 0.f;
  const int kh = j / 8, kw = j % 8;
  if (kh < 7 && kw < 7) {
    const int pix = (int)(row % 1936), oy = pix / 44, ox = pix % 44;
    const long long f = row / 1936;
    const int t = (int)(f % T) + d - 2, y = 2 * oy + kh - 3, x = 2 * ox + kw - 3;
    if (t >= 0 && t < T && y >= 0 && y < 88 && x >= 0 && x < 88)
      v = ldf(video, dt_in, ((f / T) * T + t) * 7744 + y * 88 + x);
  }
  st_op(col, row * planes * 320 + c, 320, planes, v);
}

// ---- 2-D patches of a dense NHWC map: x [n,H,H,C] -> col [n*Ho*Ho, planes*ks*ks*C], column (kh*ks+kw)*C + c
__global__ void __launch_bounds__(256)
im2col2d_kernel(const void* __restrict__ x, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho,
                __nv_bfloat16* __restrict__ col, int planes) {
  const int K = ks * ks * C;
  const long long total = n * Ho * Ho * K;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long row = idx / K;
    const int k = (int)(idx % K), tap = k / C, c = k % C, kh = tap / ks, kw = tap % ks;
    const int pix = (int)(row % (Ho * Ho)), oy = pix / Ho, ox = pix % Ho;
    const long long f = row / (Ho * Ho);
    const int y = oy * stride + kh - pad, xx = ox * stride + kw - pad;
    float v = 0.f;
    if (y >= 0 && y < H && xx >= 0 && xx < H) v = ldf(x, dt, ((f * H + y) * H + xx) * C + c);
    st_op(col, row * planes * K + k, K, planes, v);
  }
}
// gradient of the above as a gather: dx[f,y,x,c] (+)= sum over the taps whose window covers (y,x) of dcol
__global__ void __launch_bounds__(256)
col2im2d_kernel(const void* __restrict__ dcol, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho,
                float* __restrict__ dx, int accumulate) {
  const int K = ks * ks * C;
  const long long total = n * H * H * C;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xx = (int)(p % H), y = (int)((p / H) % H);
    const long long f = p / ((long long)H * H);
    float s = 0.f;
    for (int kh = 0; kh < ks; ++kh) {
      const int ny = y + pad - kh;
      if (ny < 0 || ny % stride != 0) continue;
      const int oy = ny / stride;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < ks; ++kw) {
        const int nx = xx + pad - kw;
        if (nx < 0 || nx % stride != 0) continue;
        const int ox = nx / stride;
        if (ox >= Ho) continue;
        s += ldf(dcol, dt, ((f * Ho + oy) * Ho + ox) * K + (kh * ks + kw) * C + c);
      }
    }
    dx[idx] = accumulate ? dx[idx] + s : s;
  }
}

// ---- MaxPool (3x3, stride 2, pad 1) on dense NHWC maps and its backward (recomputes the window maxima: the gradient goes to
// the FIRST maximal element in (kh,kw) order, as ATen's max_pool does)
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int dt, long long n, int H, int C, int Ho) {
  const long long total = n * Ho * Ho * C;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int ox = (int)(p % Ho), oy = (int)((p / Ho) % Ho);
    const long long f = p / ((long long)Ho * Ho);
    float m = -INFINITY;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int yy = 2 * oy + kh - 1, xx = 2 * ox + kw - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < H) m = fmaxf(m, ldf(x, dt, ((f * H + yy) * H + xx) * C + c));
      }
    stf(y, dt, idx, m);
  }
}
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const void* __restrict__ x, int dt, const float* __restrict__ dy, float* __restrict__ dx, long long n, int H,
                   int C, int Ho) {
  const long long total = n * H * H * C;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xx = (int)(p % H), y = (int)((p / H) % H);
    const long long f = p / ((long long)H * H);
    float s = 0.f;
    for (int oy = (y + 1) / 2 - 1; oy <= (y + 1) / 2; ++oy) {            // windows that may contain row y: 2*oy-1 <= y <= 2*oy+1
      if (oy < 0 || oy >= Ho || 2 * oy - 1 > y || 2 * oy + 1 < y) continue;
      for (int ox = (xx + 1) / 2 - 1; ox <= (xx + 1) / 2; ++ox) {
        if (ox < 0 || ox >= Ho || 2 * ox - 1 > xx || 2 * ox + 1 < xx) continue;
        // arg max of the window, first maximal element in scan order
        float m = -INFINITY;
        int ay = -1, ax = -1;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int yy = 2 * oy + kh - 1, x2 = 2 * ox + kw - 1;
            if (yy < 0 || yy >= H || x2 < 0 || x2 >= H) continue;
            const float v = ldf(x, dt, ((f * H + yy) * H + x2) * C + c);
            if (v > m) { m = v; ay = yy; ax = x2; }
          }
        if (ay == y && ax == xx) s += dy[((f * Ho + oy) * Ho + ox) * C + c];
      }
    }
    dx[idx] = s;
  }
}

// ---- AdaptiveAvgPool2d(1) on dense maps [n, HW, C] -> [n, C] and its backward
__global__ void __launch_bounds__(256)
avgpool_dense_kernel(const void* __restrict__ x, void* __restrict__ y, int dt, long long n, int HW, int C) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n * C) return;
  const int c = (int)(idx % C);
  const long long f = idx / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += ldf(x, dt, (f * HW + p) * C + c);
  stf(y, dt, idx, s / HW);
}
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long n, int HW, int C) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n * HW * C) return;
  const int c = (int)(idx % C);
  const long long f = idx / ((long long)HW * C);
  dx[idx] = dy[f * C + c] / HW;
}

// ---- BatchNorm (batch statistics) + optional residual + optional PReLU, backward.  Forward was
// u = gamma (raw - mean) rstd + beta; v = u + res; out = PReLU(v).  With dz = dL/d(out):
// dv = dz * (v > 0 ? 1 : slope), dslope = sum dz v [v <= 0], d_res = dv, dbeta = sum dv, dgamma = sum dv xhat,
// d_raw = gamma rstd (dv - dbeta / R - xhat dgamma / R).
// reduce: per-channel sums (double, BN_SLOTS copies like bn_stats); apply: the element-wise part.
constexpr int BWD_ROWS_PER_CTA_MAX = 2048;
__device__ __forceinline__ float bn_dv(float raw, float mean, float rstd, float gamma, float beta, float res, bool has_slope,
                                       float slope, float dz, float* v_out, float* xhat_out) {
  const float xh = (raw - mean) * rstd;
  const float v = fmaf(xh, gamma, beta) + res;
  *v_out = v;
  *xhat_out = xh;
  return has_slope ? (v > 0.f ? dz : dz * slope) : dz;
}
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const void* __restrict__ raw, int dt, const void* __restrict__ res, const float* __restrict__ dz,
                         const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ slope, long long rows, int C, int rows_per_cta, double* __restrict__ sums) {
  extern __shared__ double sh[];                 // [phases][C][3]
  const int phases = 256 / C > 0 ? 256 / C : 1;  // C <= 256: 256 / C rows at a time; C = 512: two channels per thread
  const int cpt = C > 256 ? C / 256 : 1;
  const int c0 = (threadIdx.x % (C / cpt)) * cpt, ph = threadIdx.x / (C / cpt);
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  double a[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0}, s3[2] = {0.0, 0.0};
  if (ph < phases)
    for (long long r = r0 + ph; r < r1; r += phases)
      for (int k = 0; k < cpt; ++k) {
        const int c = c0 + k;
        float v, xh;
        const float dv = bn_dv(ldf(raw, dt, r * C + c), stat[c], stat[C + c], gamma[c], beta[c],
                               res ? ldf(res, dt, r * C + c) : 0.f, slope != nullptr, slope ? slope[c] : 0.f, dz[r * C + c], &v, &xh);
        a[k] += (double)dv;
        b2[k] += (double)dv * (double)xh;
        if (slope != nullptr && v <= 0.f) s3[k] += (double)dz[r * C + c] * (double)v;
      }
  if (ph < phases)
    for (int k = 0; k < cpt; ++k) {
      sh[((size_t)ph * C + c0 + k) * 3] = a[k];
      sh[((size_t)ph * C + c0 + k) * 3 + 1] = b2[k];
      sh[((size_t)ph * C + c0 + k) * 3 + 2] = s3[k];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int p2 = 0; p2 < phases; ++p2) {
      t0 += sh[((size_t)p2 * C + c) * 3]; t1 += sh[((size_t)p2 * C + c) * 3 + 1]; t2 += sh[((size_t)p2 * C + c) * 3 + 2];
    }
    double* slot = sums + (size_t)(blockIdx.x % BN_SLOTS) * 3 * C;
    atomicAdd(&slot[c], t0);
    atomicAdd(&slot[C + c], t1);
    atomicAdd(&slot[2 * C + c], t2);
  }
}
// merges the slots: tot [3][C] floats (dbeta, dgamma, dslope), parameter gradients written, slots zeroed
__global__ void bn_act_bwd_finalize_kernel(double* __restrict__ sums, int C, float* __restrict__ tot, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta, float* __restrict__ dslope) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < BN_SLOTS; ++k)
    for (int j = 0; j < 3; ++j) {
      double* p = sums + (size_t)k * 3 * C + (size_t)j * C + c;
      t[j] += *p;
      *p = 0.0;
    }
  tot[c] = (float)t[0]; tot[C + c] = (float)t[1]; tot[2 * C + c] = (float)t[2];
  dbeta[c] = (float)t[0];
  dgamma[c] = (float)t[1];
  if (dslope != nullptr) dslope[c] = (float)t[2];
}
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const void* __restrict__ raw, int dt, const void* __restrict__ res, const float* __restrict__ dz,
                        const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ slope, const float* __restrict__ tot, long long rows, int C,
                        float* __restrict__ d_raw, float* __restrict__ d_res, int res_accumulate) {
  const long long total = rows * C;
  const float inv_rows = 1.0f / (float)rows;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx % C);
    float v, xh;
    const float dv = bn_dv(ldf(raw, dt, idx), stat[c], stat[C + c], gamma[c], beta[c], res ? ldf(res, dt, idx) : 0.f,
                           slope != nullptr, slope ? slope[c] : 0.f, dz[idx], &v, &xh);
    d_raw[idx] = gamma[c] * stat[C + c] * (dv - tot[c] * inv_rows - xh * tot[C + c] * inv_rows);
    if (d_res != nullptr) d_res[idx] = res_accumulate ? d_res[idx] + dv : dv;
  }
}
// batch mean / rstd of a BatchNorm from the (scale, bias) bn_finalize produced: rstd = scale / gamma, mean = (beta - bias) / scale
__global__ void bn_stat_from_affine_kernel(const float* __restrict__ scale, const float* __restrict__ bias,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, int C,
                                           float* __restrict__ stat) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  stat[c] = (beta[c] - bias[c]) / scale[c];
  stat[C + c] = scale[c] / gamma[c];
}

// shifted copies of the rows of a transposed operand: XS[(k*R + i), p*kp + r] = XT[i, p*kp + r + shift_k] (zero outside)
__global__ void __launch_bounds__(256)
add_f32_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) a[i] += b[i];
}


// ================= bf16-mode fast paths: 16-byte vectors along the channel axis, 64 x 64 smem-tile transposes =================
__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
// im2col2d: one thread per (patch row, 8-channel group), the taps in an unrolled loop — the row is decoded once (32-bit
// arithmetic) for KS * KS 16-byte copies.  (One thread per (row, tap, group) with 64-bit index arithmetic was instruction
// bound: 309 us for a layer1 map, 27 % of the HBM rate under ncu.)
template <int KS>
__global__ void __launch_bounds__(256)
im2col2d_vec_kernel(const uint4* __restrict__ x, unsigned rows, int H, int c8_shift, int stride, int pad, int Ho,
                    uint4* __restrict__ col) {
  const unsigned C8 = 1u << c8_shift, HoHo = (unsigned)(Ho * Ho);
  const unsigned total = rows << c8_shift;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned row = idx >> c8_shift, c8 = idx & (C8 - 1);
    const unsigned f = row / HoHo, pix = row - f * HoHo, oy = pix / (unsigned)Ho, ox = pix - oy * (unsigned)Ho;
    const uint4* img = x + ((size_t)f * H * H << c8_shift) + c8;
    uint4* out = col + ((size_t)row * (KS * KS) << c8_shift) + c8;
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const int y = (int)oy * stride + kh - pad;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const int xx = (int)ox * stride + kw - pad;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < H && xx >= 0 && xx < H) v = __ldg(img + ((size_t)(y * H + xx) << c8_shift));
        out[(size_t)(kh * KS + kw) << c8_shift] = v;
      }
    }
  }
}
// the 64 x 64 tile store shared by the transposing kernels: tile[i][(j ^ (i/8 % 8)) * 8 + e] holds element (row i, column
// 8 j + e); output row (c0 + c) receives the 8 consecutive K indices r0 + 8 g .. + 7 as one 16-byte store
__device__ __forceinline__ void store_tile_T(const __nv_bfloat16 (*tile)[72], __nv_bfloat16* __restrict__ out, long long out_row0,
                                             int c_lim, long long out_ld, long long r0) {
  // one thread per (channel pair, 8-row group): eight 32-bit shared loads (conflict-free: bank = 4 q + 4 (vec ^ g) + pair)
  // give the 8 consecutive K indices of TWO output rows
  const int cp = threadIdx.x / 8, g = threadIdx.x % 8, c = 2 * cp;
  if (c >= c_lim) return;
  uint32_t w[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    w[q] = *reinterpret_cast<const uint32_t*>(&tile[g * 8 + q][(((c >> 3) ^ g) << 3) + (c & 7)]);
  uint4 lo, hi;
  lo.x = __byte_perm(w[0], w[1], 0x5410); hi.x = __byte_perm(w[0], w[1], 0x7632);
  lo.y = __byte_perm(w[2], w[3], 0x5410); hi.y = __byte_perm(w[2], w[3], 0x7632);
  lo.z = __byte_perm(w[4], w[5], 0x5410); hi.z = __byte_perm(w[4], w[5], 0x7632);
  lo.w = __byte_perm(w[6], w[7], 0x5410); hi.w = __byte_perm(w[6], w[7], 0x7632);
  *reinterpret_cast<uint4*>(out + (out_row0 + c) * out_ld + r0 + g * 8) = lo;
  if (c + 1 < c_lim) *reinterpret_cast<uint4*>(out + (out_row0 + c + 1) * out_ld + r0 + g * 8) = hi;
}
// patches written TRANSPOSED for the weight-gradient GEMM (K = pixels): colT[(tap * C + c), r] = x[patch r, tap, c], columns
// r in [rows, kp) zero.  grid (ks * ks * C / 64, kp / 64): the taps and channel blocks of one 64-row block are neighbours in
// the launch order, so the map is read from DRAM once (tap-major order re-read it per tap: 620 MB instead of 75 MB)
__global__ void __launch_bounds__(256)
im2colT_kernel(const __nv_bfloat16* __restrict__ x, unsigned rows, int H, int C, int ks, int stride, int pad, int Ho,
               __nv_bfloat16* __restrict__ colT, long long kp) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];
  const unsigned HoHo = (unsigned)(Ho * Ho);
  const unsigned r0 = blockIdx.y * 64u;
  const int cblocks = C / 64;
  const int tap = blockIdx.x / cblocks, c0 = (blockIdx.x - tap * cblocks) * 64, kh = tap / ks, kw = tap - kh * ks;
  for (int v = threadIdx.x; v < 512; v += 256) {
    const int i = v / 8, j = v % 8;
    const unsigned r = r0 + i;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows) {
      const unsigned f = r / HoHo, pix = r - f * HoHo, oy = pix / (unsigned)Ho, ox = pix - oy * (unsigned)Ho;
      const int y = (int)oy * stride + kh - pad, xx = (int)ox * stride + kw - pad;
      if (y >= 0 && y < H && xx >= 0 && xx < H)
        val = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)f * H + y) * H + xx) * C + c0) + j);
    }
    *reinterpret_cast<uint4*>(&tile[i][(j ^ ((i >> 3) & 7)) << 3]) = val;
  }
  __syncthreads();
  store_tile_T(tile, colT, (long long)tap * C + c0, 64, kp, r0);
}
// in [rows, C] (row stride ld; bf16 or fp32) -> out bf16 [C, out_ld], out[c, r] = in[r, c], columns r in [rows, kp) zero.
// grid (kp / 64, ceil(C / 64)); C and ld multiples of 8
template <bool F32>
__global__ void __launch_bounds__(256)
transposeT_kernel(const void* __restrict__ in, long long ld, long long rows, int C, __nv_bfloat16* __restrict__ out,
                  long long out_ld, float scale) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int v = threadIdx.x; v < 512; v += 256) {
    const int i = v / 8, j = v % 8;
    const long long r = r0 + i;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows && c0 + j * 8 < C) {
      if (F32) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + r * ld + c0 + j * 8);
        const float4 a = __ldg(src), b2 = __ldg(src + 1);
        __nv_bfloat162 h[4] = {__floats2bfloat162_rn(a.x * scale, a.y * scale), __floats2bfloat162_rn(a.z * scale, a.w * scale),
                               __floats2bfloat162_rn(b2.x * scale, b2.y * scale), __floats2bfloat162_rn(b2.z * scale, b2.w * scale)};
        val = *reinterpret_cast<const uint4*>(h);
      } else {
        val = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in) + r * ld + c0) + j);
        if (scale != 1.0f) {
          float t[8];
          bf16x8_to_float(val, t);
          __nv_bfloat162 h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(t[2 * e] * scale, t[2 * e + 1] * scale);
          val = *reinterpret_cast<const uint4*>(h);
        }
      }
    }
    *reinterpret_cast<uint4*>(&tile[i][(j ^ ((i >> 3) & 7)) << 3]) = val;
  }
  __syncthreads();
  store_tile_T(tile, out, c0, C - c0 < 64 ? C - c0 : 64, out_ld, r0);
}
// col2im as a gather, vectors along the channel axis (dcol bf16: 8 channels, fp32: 4 channels per thread); 32-bit
// index arithmetic, stride 1 or 2
template <bool F32>
__global__ void __launch_bounds__(256)
col2im2d_vec_kernel(const void* __restrict__ dcol, unsigned pixels, int H, int C, int cv_shift, int ks, int stride, int pad, int Ho,
                    float* __restrict__ dx, int accumulate) {
  constexpr int V = F32 ? 4 : 8;
  const unsigned CV = 1u << cv_shift, HH = (unsigned)(H * H);
  const size_t K = (size_t)ks * ks * C;
  const unsigned total = pixels << cv_shift;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned p = idx >> cv_shift, cv = idx & (CV - 1);
    const unsigned f = p / HH, rem = p - f * HH, y = rem / (unsigned)H, xx = rem - y * (unsigned)H;
    float s[V];
#pragma unroll
    for (int e = 0; e < V; ++e) s[e] = 0.f;
    for (int kh = 0; kh < ks; ++kh) {
      const int ny = (int)y + pad - kh;
      if (ny < 0 || (stride == 2 && (ny & 1))) continue;
      const int oy = stride == 2 ? ny >> 1 : ny;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < ks; ++kw) {
        const int nx = (int)xx + pad - kw;
        if (nx < 0 || (stride == 2 && (nx & 1))) continue;
        const int ox = stride == 2 ? nx >> 1 : nx;
        if (ox >= Ho) continue;
        const size_t off = (((size_t)f * Ho + oy) * Ho + ox) * K + (size_t)(kh * ks + kw) * C + cv * V;
        if (F32) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dcol) + off));
          s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
        } else {
          float t[8];
          bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dcol) + off)), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[e] += t[e];
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + (size_t)p * C + cv * V);
#pragma unroll
    for (int q = 0; q < V / 4; ++q) {
      float4 w = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
      if (accumulate) { const float4 old = o[q]; w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w; }
      o[q] = w;
    }
  }
}
// stem patches, bf16 video: one 16-byte vector = one (dt, kh) row of the patch per thread — the 7 pixels 2 ox - 3 .. 2 ox + 3
// come from four aligned 32-bit loads (pixels 2 ox - 4 .. 2 ox + 3; out-of-image pixels fall in whole words) shifted by one
__global__ void __launch_bounds__(256)
im2col_stem_vec_kernel(const __nv_bfloat16* __restrict__ video, int B, int T, __nv_bfloat16* __restrict__ col) {
  const unsigned total = (unsigned)B * T * 1936u * 40u;
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const unsigned row = idx / 40u, v = idx - row * 40u;
  const int d = (int)(v >> 3), kh = (int)(v & 7u);
  const unsigned f = row / 1936u, pix = row - f * 1936u;
  const int oy = (int)(pix / 44u), ox = (int)(pix - (pix / 44u) * 44u);
  const int t = (int)(f % (unsigned)T) + d - 2, y = 2 * oy + kh - 3;
  uint4 out = make_uint4(0u, 0u, 0u, 0u);
  if (kh < 7 && t >= 0 && t < T && y >= 0 && y < 88) {
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(video + ((size_t)(f / (unsigned)T) * T + t) * 7744 + y * 88);
    const int w0 = ox - 2;                       // word of pixels (2 ox - 4, 2 ox - 3)
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (w0 + k >= 0 && w0 + k < 44) ? __ldg(rowp + w0 + k) : 0u;
    out.x = __byte_perm(w[0], w[1], 0x5432);
    out.y = __byte_perm(w[1], w[2], 0x5432);
    out.z = __byte_perm(w[2], w[3], 0x5432);
    out.w = w[3] >> 16;
  }
  reinterpret_cast<uint4*>(col)[idx] = out;
}
// max-pool backward, bf16 maps, 8 channels per thread
__global__ void __launch_bounds__(256)
maxpool_bwd_vec_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n,
                       int H, int C, int Ho) {
  const int C8 = C / 8;
  const long long total = n * H * H * C8;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c8 = (int)(idx % C8);
    const long long p = idx / C8;
    const int xx = (int)(p % H), y = (int)((p / H) % H);
    const long long f = p / ((long long)H * H);
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int oy = (y + 1) / 2 - 1; oy <= (y + 1) / 2; ++oy) {
      if (oy < 0 || oy >= Ho || 2 * oy - 1 > y || 2 * oy + 1 < y) continue;
      for (int ox = (xx + 1) / 2 - 1; ox <= (xx + 1) / 2; ++ox) {
        if (ox < 0 || ox >= Ho || 2 * ox - 1 > xx || 2 * ox + 1 < xx) continue;
        float m[8];
        int arg[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; arg[e] = -1; }
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int yy = 2 * oy + kh - 1, x2 = 2 * ox + kw - 1;
            if (yy < 0 || yy >= H || x2 < 0 || x2 >= H) continue;
            float t[8];
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + ((f * H + yy) * H + x2) * C) + c8), t);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (t[e] > m[e]) { m[e] = t[e]; arg[e] = kh * 3 + kw; }
          }
        const int mine = (y - (2 * oy - 1)) * 3 + (xx - (2 * ox - 1));
        const float4* g = reinterpret_cast<const float4*>(dy + ((f * Ho + oy) * Ho + ox) * C + c8 * 8);
        const float4 g0 = __ldg(g), g1 = __ldg(g + 1);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (arg[e] == mine) s[e] += gv[e];
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + p * C + c8 * 8);
    o[0] = make_float4(s[0], s[1], s[2], s[3]);
    o[1] = make_float4(s[4], s[5], s[6], s[7]);
  }
}
// max-pool forward, bf16 maps, 8 channels per thread; `arg` (optional, [n,Ho,Ho,C] bytes) keeps the window position
// (kh*3+kw, first maximum in scan order) so that the backward is a pure gather
__global__ void __launch_bounds__(256)
maxpool_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ yout, unsigned char* __restrict__ arg_out,
                       long long n, int H, int C, int Ho) {
  const int C8 = C / 8;
  const long long total = n * Ho * Ho * C8;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c8 = (int)(idx % C8);
    const long long p = idx / C8;
    const int ox = (int)(p % Ho), oy = (int)((p / Ho) % Ho);
    const long long f = p / ((long long)Ho * Ho);
    float m[8];
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; arg[e] = 0; }
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int yy = 2 * oy + kh - 1, xx = 2 * ox + kw - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= H) continue;
        float t[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + ((f * H + yy) * H + xx) * C) + c8), t);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (t[e] > m[e]) { m[e] = t[e]; arg[e] = kh * 3 + kw; }
      }
    __align__(16) __nv_bfloat16 w[8];
    __align__(8) unsigned char a8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { w[e] = __float2bfloat16_rn(m[e]); a8[e] = (unsigned char)arg[e]; }
    reinterpret_cast<uint4*>(yout)[idx] = *reinterpret_cast<const uint4*>(w);
    if (arg_out != nullptr) reinterpret_cast<uint2*>(arg_out)[idx] = *reinterpret_cast<const uint2*>(a8);
  }
}
__global__ void __launch_bounds__(256)
maxpool_bwd_arg_kernel(const unsigned char* __restrict__ arg, const float* __restrict__ dy, float* __restrict__ dx, long long n,
                       int H, int C, int Ho) {
  const int C8 = C / 8;
  const long long total = n * H * H * C8;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c8 = (int)(idx % C8);
    const long long p = idx / C8;
    const int xx = (int)(p % H), y = (int)((p / H) % H);
    const long long f = p / ((long long)H * H);
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int oy = (y + 1) / 2 - 1; oy <= (y + 1) / 2; ++oy) {
      if (oy < 0 || oy >= Ho || 2 * oy - 1 > y || 2 * oy + 1 < y) continue;
      for (int ox = (xx + 1) / 2 - 1; ox <= (xx + 1) / 2; ++ox) {
        if (ox < 0 || ox >= Ho || 2 * ox - 1 > xx || 2 * ox + 1 < xx) continue;
        const long long o = ((f * Ho + oy) * Ho + ox) * C + c8 * 8;
        const uint2 au = __ldg(reinterpret_cast<const uint2*>(arg + o));
        const unsigned char* a8 = reinterpret_cast<const unsigned char*>(&au);
        const int mine = (y - (2 * oy - 1)) * 3 + (xx - (2 * ox - 1));
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy + o)), g1 = __ldg(reinterpret_cast<const float4*>(dy + o) + 1);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if ((int)a8[e] == mine) s[e] += gv[e];
      }
    }
    float4* o4 = reinterpret_cast<float4*>(dx + p * C + c8 * 8);
    o4[0] = make_float4(s[0], s[1], s[2], s[3]);
    o4[1] = make_float4(s[4], s[5], s[6], s[7]);
  }
}
// BatchNorm + residual + PReLU backward, bf16 maps: 8 channels per thread.  reduce: lane = channel group, 256 / (C / 8) row
// phases, fp32 partials per thread (<= BWD_ROWS_PER_CTA_MAX / phases rows), summed across phases and CTAs in double.
__device__ __forceinline__ void bn_dv8(const __nv_bfloat16* raw, const __nv_bfloat16* res, const float* dz, long long off,
                                       const float* mean, const float* rstd, const float* gamma, const float* beta,
                                       const float* slope, bool has_slope, float* dv, float* v, float* xh, float* dzv) {
  float r[8], rs[8];
  bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(raw + off)), r);
  if (res != nullptr) bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(res + off)), rs);
  const float4 d0 = __ldg(reinterpret_cast<const float4*>(dz + off)), d1 = __ldg(reinterpret_cast<const float4*>(dz + off) + 1);
  const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    xh[e] = (r[e] - mean[e]) * rstd[e];
    v[e] = fmaf(xh[e], gamma[e], beta[e]) + (res != nullptr ? rs[e] : 0.f);
    dzv[e] = d[e];
    dv[e] = has_slope ? (v[e] > 0.f ? d[e] : d[e] * slope[e]) : d[e];
  }
}
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_vec_kernel(const __nv_bfloat16* __restrict__ raw, const __nv_bfloat16* __restrict__ res,
                             const float* __restrict__ dz, const float* __restrict__ stat, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ slope, long long rows, int C,
                             int rows_per_cta, double* __restrict__ sums) {
  extern __shared__ float shf[];                 // [phases][C][3]
  const int lanes = C / 8, phases = 256 / lanes;
  const int l = threadIdx.x % lanes, ph = threadIdx.x / lanes, c0 = l * 8;
  float mean[8], rstd[8], gm[8], bt[8], sl[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mean[e] = stat[c0 + e]; rstd[e] = stat[C + c0 + e]; gm[e] = gamma[c0 + e]; bt[e] = beta[c0 + e];
    sl[e] = slope != nullptr ? slope[c0 + e] : 0.f;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float a[8], b2[8], s3[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = 0.f; b2[e] = 0.f; s3[e] = 0.f; }
  for (long long r = r0 + ph; r < r1; r += phases) {
    float dv[8], v[8], xh[8], d[8];
    bn_dv8(raw, res, dz, r * C + c0, mean, rstd, gm, bt, sl, slope != nullptr, dv, v, xh, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a[e] += dv[e];
      b2[e] = fmaf(dv[e], xh[e], b2[e]);
      if (slope != nullptr && v[e] <= 0.f) s3[e] = fmaf(d[e], v[e], s3[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    shf[((size_t)ph * C + c0 + e) * 3] = a[e];
    shf[((size_t)ph * C + c0 + e) * 3 + 1] = b2[e];
    shf[((size_t)ph * C + c0 + e) * 3 + 2] = s3[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int p2 = 0; p2 < phases; ++p2) {
      t0 += (double)shf[((size_t)p2 * C + c) * 3]; t1 += (double)shf[((size_t)p2 * C + c) * 3 + 1];
      t2 += (double)shf[((size_t)p2 * C + c) * 3 + 2];
    }
    double* slot = sums + (size_t)(blockIdx.x % BN_SLOTS) * 3 * C;
    atomicAdd(&slot[c], t0);
    atomicAdd(&slot[C + c], t1);
    atomicAdd(&slot[2 * C + c], t2);
  }
}
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_vec_kernel(const __nv_bfloat16* __restrict__ raw, const __nv_bfloat16* __restrict__ res,
                            const float* __restrict__ dz, const float* __restrict__ stat, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const float* __restrict__ slope, const float* __restrict__ tot,
                            long long rows, int C, void* __restrict__ d_raw, int d_raw_bf16, float* __restrict__ d_res,
                            int res_accumulate) {
  const int C8 = C / 8;
  const long long total = rows * C8;
  const float inv_rows = 1.0f / (float)rows;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int c0 = (int)(idx % C8) * 8;
    float mean[8], rstd[8], gm[8], bt[8], sl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mean[e] = stat[c0 + e]; rstd[e] = stat[C + c0 + e]; gm[e] = gamma[c0 + e]; bt[e] = beta[c0 + e];
      sl[e] = slope != nullptr ? slope[c0 + e] : 0.f;
    }
    float dv[8], v[8], xh[8], d[8], o[8];
    const long long off = idx * 8;
    bn_dv8(raw, res, dz, off, mean, rstd, gm, bt, sl, slope != nullptr, dv, v, xh, d);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      o[e] = gm[e] * rstd[e] * (dv[e] - tot[c0 + e] * inv_rows - xh[e] * tot[C + c0 + e] * inv_rows);
    if (d_raw_bf16) {       // bf16 mode: the only readers are GEMM operands (weight / input gradient), which are bf16 anyway
      __nv_bfloat162 h[4] = {__floats2bfloat162_rn(o[0], o[1]), __floats2bfloat162_rn(o[2], o[3]),
                             __floats2bfloat162_rn(o[4], o[5]), __floats2bfloat162_rn(o[6], o[7])};
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d_raw) + off) = *reinterpret_cast<const uint4*>(h);
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(d_raw) + off);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (d_res != nullptr) {
      float4* dr = reinterpret_cast<float4*>(d_res + off);
      if (res_accumulate) {
        const float4 p0 = dr[0], p1 = dr[1];
        dv[0] += p0.x; dv[1] += p0.y; dv[2] += p0.z; dv[3] += p0.w; dv[4] += p1.x; dv[5] += p1.y; dv[6] += p1.z; dv[7] += p1.w;
      }
      dr[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
      dr[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
    }
  }
}

inline unsigned grid_for(long long n) {
  const long long b = (n + 255) / 256;
  return (unsigned)(b < 148 * 32 ? (b > 0 ? b : 1) : 148 * 32);
}

}  // namespace

int launch_im2col_stem(const void* video, int dt, int B, int T, void* col, int planes, cudaStream_t stream) {
  const long long total = (long long)B * T * 1936 * 320;
  AVH_CHECK((total + 255) / 256 < (1ll << 31), "stem patch matrix too large");
  if (dt == DT_BF16 && planes == 1 && total / 8 < (1ll << 32) && (reinterpret_cast<uintptr_t>(video) & 3) == 0) {
    im2col_stem_vec_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(video), B, T, reinterpret_cast<__nv_bfloat16*>(col));
    AVH_CUDA_OK(cudaGetLastError());
    count_launch(1);
    return 0;
  }
  im2col_stem_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(video, dt, B, T, reinterpret_cast<__nv_bfloat16*>(col),
                                                                         planes);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_im2col2d(const void* x, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho, void* col, int planes,
                    cudaStream_t stream) {
  {
    const int C8 = C / 8;
    const bool pow2 = C % 8 == 0 && (C8 & (C8 - 1)) == 0;
    const long long rows = n * Ho * Ho;
    if (dt == DT_BF16 && planes == 1 && pow2 && (ks == 3 || ks == 1) && rows * C8 < (1ll << 32) && n * H * H < (1ll << 31)) {
      int sh = 0;
      while ((1 << sh) < C8) ++sh;
      const uint4* xi = reinterpret_cast<const uint4*>(x);
      uint4* co = reinterpret_cast<uint4*>(col);
      if (ks == 3)
        im2col2d_vec_kernel<3><<<grid_for(rows * C8), 256, 0, stream>>>(xi, (unsigned)rows, H, sh, stride, pad, Ho, co);
      else
        im2col2d_vec_kernel<1><<<grid_for(rows * C8), 256, 0, stream>>>(xi, (unsigned)rows, H, sh, stride, pad, Ho, co);
      AVH_CUDA_OK(cudaGetLastError());
      count_launch(1);
      return 0;
    }
  }
  im2col2d_kernel<<<grid_for(n * Ho * Ho * ks * ks * C), 256, 0, stream>>>(x, dt, n, H, C, ks, stride, pad, Ho,
                                                                          reinterpret_cast<__nv_bfloat16*>(col), planes);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_col2im2d(const void* dcol, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho, float* dx,
                    int accumulate, cudaStream_t stream) {
  {
    const int V = dt == DT_BF16 ? 8 : 4;
    const int CV = C / V;
    const bool pow2 = C % V == 0 && CV > 0 && (CV & (CV - 1)) == 0;
    const long long pixels = n * H * H;
    if ((dt == DT_BF16 || dt == DT_F32) && pow2 && (stride == 1 || stride == 2) && pixels * CV < (1ll << 32)) {
      int sh = 0;
      while ((1 << sh) < CV) ++sh;
      if (dt == DT_BF16)
        col2im2d_vec_kernel<false><<<grid_for(pixels * CV), 256, 0, stream>>>(dcol, (unsigned)pixels, H, C, sh, ks, stride, pad, Ho, dx,
                                                                             accumulate);
      else
        col2im2d_vec_kernel<true><<<grid_for(pixels * CV), 256, 0, stream>>>(dcol, (unsigned)pixels, H, C, sh, ks, stride, pad, Ho, dx,
                                                                            accumulate);
      AVH_CUDA_OK(cudaGetLastError());
      count_launch(1);
      return 0;
    }
  }
  col2im2d_kernel<<<grid_for(n * H * H * C), 256, 0, stream>>>(dcol, dt, n, H, C, ks, stride, pad, Ho, dx, accumulate);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_maxpool_dense(const void* x, void* y, int dt, long long n, int H, int C, int Ho, cudaStream_t stream,
                         unsigned char* arg) {
  if (dt == DT_BF16 && C % 8 == 0) {
    maxpool_fwd_vec_kernel<<<grid_for(n * Ho * Ho * (C / 8)), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), arg, n, H, C, Ho);
  } else {
    AVH_CHECK(arg == nullptr, "max-pool window positions are kept for bf16 maps only");
    maxpool_fwd_kernel<<<grid_for(n * Ho * Ho * C), 256, 0, stream>>>(x, y, dt, n, H, C, Ho);
  }
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_maxpool_bwd(const void* x, int dt, const float* dy, float* dx, long long n, int H, int C, int Ho, cudaStream_t stream,
                       const unsigned char* arg) {
  if (arg != nullptr && C % 8 == 0)
    maxpool_bwd_arg_kernel<<<grid_for(n * H * H * (C / 8)), 256, 0, stream>>>(arg, dy, dx, n, H, C, Ho);
  else if (dt == DT_BF16 && C % 8 == 0)
    maxpool_bwd_vec_kernel<<<grid_for(n * H * H * (C / 8)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), dy, dx, n, H,
                                                                             C, Ho);
  else
    maxpool_bwd_kernel<<<grid_for(n * H * H * C), 256, 0, stream>>>(x, dt, dy, dx, n, H, C, Ho);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_avgpool_dense(const void* x, void* y, int dt, long long n, int HW, int C, cudaStream_t stream) {
  avgpool_dense_kernel<<<(unsigned)((n * C + 255) / 256), 256, 0, stream>>>(x, y, dt, n, HW, C);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_avgpool_bwd(const float* dy, float* dx, long long n, int HW, int C, cudaStream_t stream) {
  avgpool_bwd_kernel<<<(unsigned)((n * HW * C + 255) / 256), 256, 0, stream>>>(dy, dx, n, HW, C);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_bn_stat_from_affine(const float* scale, const float* bias, const float* gamma, const float* beta, int C, float* stat,
                               cudaStream_t stream) {
  bn_stat_from_affine_kernel<<<(C + 127) / 128, 128, 0, stream>>>(scale, bias, gamma, beta, C, stat);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_bn_act_bwd(const void* raw, int dt, const void* res, const float* dz, const float* stat, const float* gamma,
                      const float* beta, const float* slope, long long rows, int C, double* sums, float* tot, void* d_raw,
                      float* d_res, int res_accumulate, float* dgamma, float* dbeta, float* dslope, cudaStream_t stream,
                      int d_raw_dt) {
  AVH_CHECK(C >= 32 && C <= 512 && (C <= 256 ? 256 % C == 0 : C % 256 == 0), "bn backward: channel count must divide or double 256");
  if (rows <= 0) return 0;
  if (dt == DT_BF16) {
    const int phases = 256 / (C / 8);
    long long rpc = (rows + 4 * 148 - 1) / (4 * 148);
    rpc = (rpc + phases - 1) / phases * phases;
    if (rpc < 4 * phases) rpc = 4 * phases;
    if (rpc > BWD_ROWS_PER_CTA_MAX) rpc = BWD_ROWS_PER_CTA_MAX;
    const size_t smem = (size_t)phases * C * 3 * sizeof(float);
    const __nv_bfloat16* rw = reinterpret_cast<const __nv_bfloat16*>(raw);
    const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(res);
    bn_act_bwd_reduce_vec_kernel<<<(unsigned)((rows + rpc - 1) / rpc), 256, smem, stream>>>(rw, rs, dz, stat, gamma, beta, slope,
                                                                                          rows, C, (int)rpc, sums);
    AVH_CUDA_OK(cudaGetLastError());
    bn_act_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, C, tot, dgamma, dbeta, dslope);
    AVH_CUDA_OK(cudaGetLastError());
    bn_act_bwd_apply_vec_kernel<<<grid_for(rows * (C / 8)), 256, 0, stream>>>(rw, rs, dz, stat, gamma, beta, slope, tot, rows, C,
                                                                            d_raw, d_raw_dt == DT_BF16, d_res, res_accumulate);
    AVH_CUDA_OK(cudaGetLastError());
    count_launch(3);
    return 0;
  }
  const int phases = 256 / C > 0 ? 256 / C : 1;
  long long rpc = (rows + 4 * 148 - 1) / (4 * 148);
  rpc = (rpc + phases - 1) / phases * phases;
  if (rpc < 4 * phases) rpc = 4 * phases;
  if (rpc > BWD_ROWS_PER_CTA_MAX) rpc = BWD_ROWS_PER_CTA_MAX;
  const size_t smem = (size_t)phases * C * 3 * sizeof(double);
  AVH_CHECK(smem <= 48 * 1024, "bn backward: reduction buffer too large");
  bn_act_bwd_reduce_kernel<<<(unsigned)((rows + rpc - 1) / rpc), 256, smem, stream>>>(raw, dt, res, dz, stat, gamma, beta, slope,
                                                                                    rows, C, (int)rpc, sums);
  AVH_CUDA_OK(cudaGetLastError());
  bn_act_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, C, tot, dgamma, dbeta, dslope);
  AVH_CUDA_OK(cudaGetLastError());
  AVH_CHECK(d_raw_dt == DT_F32, "bf16 gradient maps need bf16 activations");
  bn_act_bwd_apply_kernel<<<grid_for(rows * C), 256, 0, stream>>>(raw, dt, res, dz, stat, gamma, beta, slope, tot, rows, C,
                                                                  reinterpret_cast<float*>(d_raw), d_res, res_accumulate);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(3);
  return 0;
}
int launch_add_f32(float* a, const float* b, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  add_f32_kernel<<<grid_for(n), 256, 0, stream>>>(a, b, n);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

int launch_im2colT(const void* x, long long n, int H, int C, int ks, int stride, int pad, int Ho, void* colT, long long kp,
                   cudaStream_t stream) {
  AVH_CHECK(C % 64 == 0 && kp % 64 == 0 && kp >= n * Ho * Ho, "transposed patches: channels / K padding must be multiples of 64");
  AVH_CHECK(kp / 64 <= 65535 && n * Ho * Ho < (1ll << 31), "transposed patches: too many rows for one launch");
  dim3 grid((unsigned)(ks * ks * (C / 64)), (unsigned)(kp / 64));
  im2colT_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), (unsigned)(n * Ho * Ho), H, C, ks, stride, pad, Ho,
                                           reinterpret_cast<__nv_bfloat16*>(colT), kp);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_transposeT(const void* in, int dt, long long ld, long long rows, int C, void* out, long long kp, long long out_ld,
                      cudaStream_t stream, float scale) {
  AVH_CHECK((dt == DT_BF16 || dt == DT_F32) && C % 8 == 0 && ld % 8 == 0 && kp % 64 == 0 && kp >= rows && out_ld % 8 == 0,
            "tile transpose: bf16 / fp32 input, channel count and strides multiples of 8");
  dim3 grid((unsigned)(kp / 64), (unsigned)((C + 63) / 64));
  if (dt == DT_F32)
    transposeT_kernel<true><<<grid, 256, 0, stream>>>(in, ld, rows, C, reinterpret_cast<__nv_bfloat16*>(out), out_ld, scale);
  else
    transposeT_kernel<false><<<grid, 256, 0, stream>>>(in, ld, rows, C, reinterpret_cast<__nv_bfloat16*>(out), out_ld, scale);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace avh
