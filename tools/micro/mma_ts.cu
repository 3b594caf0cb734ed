// Micro-test: (1) tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st): layout check against a host
// reference and issue rate; (2) M = 64 vs M = 128 instruction rate at wide N from smem operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_ts mma_ts.cu && ./mma_ts
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../multimodalvc_b200/csrc/common.cuh"
namespace avh { void set_last_error(const std::string&) {} int device_sm_count() { return 148; } void count_launch(int) {} }
using namespace avh;

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__host__ __device__ inline int aval(int m, int k) { return (m * 3 + k * 5) % 7 - 3; }
__host__ __device__ inline int bval(int n, int k) { return (n + 2 * k) % 5 - 2; }

// ---- correctness: D[128 x N] = A[128 x 16] (TMEM) * B[N x 16]^T (smem, SW128 K-major, only the first 16 of 64 K used)
__global__ void __launch_bounds__(128, 1) ts_check(int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B tile: row n, 128 B per row, 16-byte chunk c at (c ^ (n & 7))
  for (int i = threadIdx.x; i < 256 * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)(c * 8 + e < 16 ? bval(n, c * 8 + e) : 0));
    *reinterpret_cast<uint4*>(smem + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // A: thread m (lane m of TMEM) writes 8 columns at column 256: column j = (k=2j low half, k=2j+1 high half)
  {
    const int m = threadIdx.x;
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn((float)aval(m, 2 * j), (float)aval(m, 2 * j + 1));
      r[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st_32x8(tmem + ((uint32_t)(warp * 32) << 16) + 256, r);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      umma_bf16_ts(tmem, tmem + 256, umma_desc_sw128(smem_u32(smem)), umma_idesc_bf16(128, N), 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)threadIdx.x * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- rate: mode 0 = SS M=128, 1 = SS M=64, 2 = TS M=128
__global__ void __launch_bounds__(128, 1) rate(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(mode == 1 ? 64 : 128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
    const uint64_t ad = umma_desc_sw128(a), bd = umma_desc_sw128(b);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      // a k-block as the real kernels issue it: 4 MMAs back to back inside one elected section; alternate between
      // two accumulators (mode >= 3) or keep one
      const uint32_t d = tmem + ((mode >= 3 && (i & 4)) ? 128 : 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 2 || mode == 3) umma_bf16_ts(d, tmem + 256 + 8 * k, bd + 2 * k, idesc, 1);
          else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 256 * sizeof(float));
  cudaFuncSetAttribute(ts_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int N : {64, 224}) {
    cudaMemset(d_out, 0, 128 * 256 * sizeof(float));
    ts_check<<<1, 128, 64 * 1024>>>(N, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts_check N=%d failed: %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(128 * 256);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int k = 0; k < 16; ++k) ref += aval(m, k) * bval(n, k);
        if (h[m * 256 + n] != (float)ref) {
          if (bad < 5) printf("  mismatch m=%d n=%d got %g want %d\n", m, n, h[m * 256 + n], ref);
          ++bad;
        }
      }
    printf("TS layout check N=%d: %s (%d mismatches)\n", N, bad ? "FAIL" : "OK", bad);
  }
  long long* d_t;
  cudaMalloc(&d_t, 8 * sizeof(long long));
  const int iters = 4000;
  for (int mode : {0, 1, 2, 3, 4})
    for (int N : {64, 112, 128, 224, 256}) {
      if (mode >= 3 && N > 128) continue;
      rate<<<1, 128, 80 * 1024>>>(N, mode, iters, d_t);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate mode %d N=%d failed: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      long long t;
      cudaMemcpy(&t, d_t, sizeof(t), cudaMemcpyDeviceToHost);
      const char* names[5] = {"SS M=128", "SS M=64", "TS M=128", "TS M=128 2acc", "SS M=128 2acc"};
      printf("%s N=%3d: %.1f clk/MMA (M=128 pipe ideal %d)\n", names[mode], N, (double)t / iters, N / 2);
    }
  return 0;
}
