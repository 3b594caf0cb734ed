// Training step of the lip ResNet (`feature_grad_mult > 0`): element-wise / gather kernels of a
// correctness-first path in which every convolution is "explicit patches -> one GEMM" on DENSE NHWC maps:
//   forward   col = im2col(x);  raw = col W^T (tcgen05 GEMM);  BatchNorm with batch statistics + PReLU (+ residual)
//   backward  dW = d_raw^T col (GEMM over K = pixels);  d_col = d_raw W (GEMM);  dx = col2im(d_col) (gather, no atomics)
// avhubert/resnet.py:35-74 (BasicBlock), :131-169 (ResEncoder: Conv3d stem, BatchNorm3d, PReLU, MaxPool3d, trunk, avgpool).
// The inference path keeps its fused kernels; this file only serves avh_train_forward / avh_encoder_backward.
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

// ---- stem patches: video [B,1,T,88,88] -> col [B*T*1936, planes*320]; column dt*64 + kh*7 + kw (49..63 of a block zero)
// = video[b, t+dt-2, 2*oy+kh-3, 2*ox+kw-3] (Conv3d(1,64,(5,7,7),stride (1,2,2),pad (2,3,3)), resnet.py:137)
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const void* __restrict__ video, int dt_in, int B, int T, __nv_bfloat16* __restrict__ col, int planes) {
  const long long rows = (long long)B * T * 1936;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;      // (row, dt, j)
  if (idx >= rows * 320) return;
  const long long row = idx / 320;
  const int c = (int)(idx % 320), d = c / 64, j = c % 64;
  float v = 0.f;
  if (j < 49) {
    const int kh = j / 7, kw = j % 7;
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

inline unsigned grid_for(long long n) {
  const long long b = (n + 255) / 256;
  return (unsigned)(b < 148 * 32 ? (b > 0 ? b : 1) : 148 * 32);
}

}  // namespace

int launch_im2col_stem(const void* video, int dt, int B, int T, void* col, int planes, cudaStream_t stream) {
  const long long total = (long long)B * T * 1936 * 320;
  AVH_CHECK((total + 255) / 256 < (1ll << 31), "stem patch matrix too large");
  im2col_stem_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(video, dt, B, T, reinterpret_cast<__nv_bfloat16*>(col),
                                                                         planes);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_im2col2d(const void* x, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho, void* col, int planes,
                    cudaStream_t stream) {
  im2col2d_kernel<<<grid_for(n * Ho * Ho * ks * ks * C), 256, 0, stream>>>(x, dt, n, H, C, ks, stride, pad, Ho,
                                                                          reinterpret_cast<__nv_bfloat16*>(col), planes);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_col2im2d(const void* dcol, int dt, long long n, int H, int C, int ks, int stride, int pad, int Ho, float* dx,
                    int accumulate, cudaStream_t stream) {
  col2im2d_kernel<<<grid_for(n * H * H * C), 256, 0, stream>>>(dcol, dt, n, H, C, ks, stride, pad, Ho, dx, accumulate);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_maxpool_dense(const void* x, void* y, int dt, long long n, int H, int C, int Ho, cudaStream_t stream) {
  maxpool_fwd_kernel<<<grid_for(n * Ho * Ho * C), 256, 0, stream>>>(x, y, dt, n, H, C, Ho);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}
int launch_maxpool_bwd(const void* x, int dt, const float* dy, float* dx, long long n, int H, int C, int Ho, cudaStream_t stream) {
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
                      const float* beta, const float* slope, long long rows, int C, double* sums, float* tot, float* d_raw,
                      float* d_res, int res_accumulate, float* dgamma, float* dbeta, float* dslope, cudaStream_t stream) {
  AVH_CHECK(C >= 32 && C <= 512 && (C <= 256 ? 256 % C == 0 : C % 256 == 0), "bn backward: channel count must divide or double 256");
  if (rows <= 0) return 0;
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
  bn_act_bwd_apply_kernel<<<grid_for(rows * C), 256, 0, stream>>>(raw, dt, res, dz, stat, gamma, beta, slope, tot, rows, C, d_raw,
                                                                  d_res, res_accumulate);
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

}  // namespace avh
