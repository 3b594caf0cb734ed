// Experiment: can a K-major SWIZZLE_128B UMMA operand descriptor start at an arbitrary ROW of a larger smem
// window (start address = base + s*128 B, not 1024-aligned)?  With or without the descriptor's base_offset field?
// A window: 160 rows x 64 bf16, value A[r][k] = (r==... pattern); B = 64x64 identity.  D[i][j] should equal
// A[s+i][j] for a descriptor starting at row s.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../multimodalvc_b200/csrc/common.cuh"
namespace avh { void set_last_error(const std::string&) {} int device_sm_count() { return 148; } void count_launch(int) {} bool pdl_enabled() { return false; } }
using namespace avh;

__device__ __forceinline__ uint64_t desc_sw128_off(uint32_t addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: base_offset = 0; mode 1: base_offset = (addr >> 7) & 7
__global__ void __launch_bounds__(128, 1) test(int shift, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // 160 rows x 128 B, swizzled
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 32768);         // 64 rows x 128 B, swizzled
  // element (r, k) lives at byte r*128 + ((k/8) ^ (r%8))*16 + (k%8)*2  (TMA SWIZZLE_128B, absolute-address based:
  // the window base is 1024-aligned so r%8 == address bits [7:9])
  for (int i = threadIdx.x; i < 160 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const float v = (float)(r * 0.5f) + (float)k * 0.001953125f * 8.0f;     // exactly representable-ish in bf16? keep small
    A[(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) / 2] = __float2bfloat16_rn((float)((r * 64 + k) % 251));
    (void)v;
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    B[(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) / 2] = __float2bfloat16_rn(r == k ? 1.f : 0.f);
  }
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(&slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 64);
      const uint32_t a = smem_u32(smem) + shift * 128, b = smem_u32(smem + 32768);
      const uint32_t bo = mode == 1 ? ((a >> 7) & 7) : 0;
      const uint64_t ad = desc_sw128_off(a, bo), bd = desc_sw128_off(b, 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * 4);
  cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> h(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 2, 3, 7, 8, 9, 23, 24}) {
      cudaMemset(d_out, 0, 128 * 64 * 4);
      test<<<1, 128, 64 * 1024>>>(shift, mode, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first_bad = -1;
      for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 64; ++j) {
          const float want = (float)(((shift + i) * 64 + j) % 251);
          if (h[i * 64 + j] != want) { if (!bad) first_bad = i * 64 + j; ++bad; }
        }
      printf("mode %d (base_offset %s) shift %2d: %s (%d mismatches%s)\n", mode, mode ? "set" : "0", shift,
             bad ? "WRONG" : "ok", bad, bad ? "" : "");
      if (bad && shift <= 2) printf("   first mismatch at row %d col %d: got %.0f want %.0f\n", first_bad / 64, first_bad % 64,
                                    h[first_bad], (float)(((shift + first_bad / 64) * 64 + first_bad % 64) % 251));
    }
  return 0;
}
