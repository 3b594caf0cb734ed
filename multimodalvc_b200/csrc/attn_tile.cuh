// Shared-memory tile loader of the fp32 CUDA-core attention kernels (qformer.cu, backward_ops.cu): R rows x 64 channels
// of a [rows, ld] matrix (bf16 / fp16 / fp32) -> fp32 rows of stride DS in shared memory, rows >= limit zero-filled.
// 16-byte global loads when the layout allows (8 bf16 or 4 fp32 per load), scalar loads otherwise.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace avh {

template <int R, int DS, int NT>
__device__ __forceinline__ void attn_load_tile(const void* __restrict__ M, int dt, long long ld, long long row_base, int first,
                                               int limit, int col0, float* __restrict__ dst) {
  const bool vec_ok = (ld % 8 == 0) && (col0 % 8 == 0) && ((reinterpret_cast<uintptr_t>(M) & 15) == 0);
  if (dt == DT_BF16 && vec_ok) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(M);
    for (int e = threadIdx.x; e < R * 8; e += NT) {
      const int r = e >> 3, v = e & 7;
      float f[8];
      if (first + r < limit) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + (row_base + first + r) * ld + col0 + v * 8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h2[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[r * DS + v * 8 + k] = f[k];
    }
    return;
  }
  if (dt == DT_F32 && vec_ok) {
    const float* base = reinterpret_cast<const float*>(M);
    for (int e = threadIdx.x; e < R * 16; e += NT) {
      const int r = e >> 4, v = e & 15;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (first + r < limit) u = *reinterpret_cast<const float4*>(base + (row_base + first + r) * ld + col0 + v * 4);
      dst[r * DS + v * 4] = u.x; dst[r * DS + v * 4 + 1] = u.y; dst[r * DS + v * 4 + 2] = u.z; dst[r * DS + v * 4 + 3] = u.w;
    }
    return;
  }
  for (int e = threadIdx.x; e < R * 64; e += NT) {
    const int r = e >> 6, d = e & 63;
    float x = 0.f;
    if (first + r < limit) {
      const long long i = (row_base + first + r) * ld + col0 + d;
      if (dt == DT_BF16) x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(M)[i]);
      else if (dt == DT_F16) x = __half2float(reinterpret_cast<const __half*>(M)[i]);
      else x = reinterpret_cast<const float*>(M)[i];
    }
    dst[r * DS + d] = x;
  }
}

}  // namespace avh
