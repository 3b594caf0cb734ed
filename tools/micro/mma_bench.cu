// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128 per CTA) from resident smem operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../multimodalvc_b200/csrc/common.cuh"
namespace avh { void set_last_error(const std::string&) {} int device_sm_count() { return 148; } void count_launch(int) {} }
using namespace avh;

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode 0: same accumulator, k advances 0..3 (as the GEMM does); mode 1: alternate 2 accumulators;
// mode 2: k fixed 0; mode 3: SWIZZLE_32B layout (K16 sub-tiles contiguous, 4 KB apart)
template <int PAIR>
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = PAIR == 1 || cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 1) { if (PAIR == 2) tmem_alloc_pair(&slot, 512); else tmem_alloc(&slot, 512); }
  tc_fence_before();
  if (PAIR == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (mode == 4 || mode == 5) {
    // two issuing threads (warps 0 and 2): mode 4 -> different accumulators, mode 5 -> the SAME accumulator
    __shared__ uint64_t bar2[2];
    if (threadIdx.x == 0) { mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); mbar_fence_init(); }
    __syncthreads();
    if ((warp == 0 || warp == 2) && lane == 0) {
      const int w = warp >> 1;
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = i & 3;
        const uint64_t ad = umma_desc_sw128(a) + 2 * k, bd = umma_desc_sw128(b) + 2 * k;
        umma_bf16(tmem + ((mode == 4 && w) ? 256 : 0), ad, bd, idesc, 1);
      }
      umma_commit(&bar2[w]);
      mbar_wait(&bar2[w], 0);
      const long long t1 = clock64();
      if (w == 0) out[blockIdx.x] = (t1 - t0) / 2;     // two issuers: report clk per MMA of the pair
    }
  } else if (mode == 6 || mode == 7) {
    // canonical form: the WHOLE warp runs the loop with warp-uniform operands, one elected lane issues
    if (warp == 0 && leader) {
      const uint32_t idesc = umma_idesc_bf16(128 * PAIR, N);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
        const uint64_t ad = umma_desc_sw128(a), bd = umma_desc_sw128(b);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1);
          if (mode == 7) umma_commit(&bar);       // a commit per k-block, like the GEMM (barrier never waited here)
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      if (mode == 6) mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
  } else if (warp == 0 && lane == 0 && leader) {
    const uint32_t idesc = umma_idesc_bf16(128 * PAIR, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint64_t ad, bd;
      const int k = i & 3;
      if (mode == 3) { ad = desc_generic(a + k * 4096, 256, 6); bd = desc_generic(b + k * 8192, 256, 6); }
      else if (mode == 2) { ad = umma_desc_sw128(a); bd = umma_desc_sw128(b); }
      else { ad = umma_desc_sw128(a) + 2 * k; bd = umma_desc_sw128(b) + 2 * k; }
      const uint32_t d = tmem + ((mode == 1 && (i & 4)) ? 256 : 0);
      if (PAIR == 2) umma_bf16_pair(d, ad, bd, idesc, 1); else umma_bf16(d, ad, bd, idesc, 1);
    }
    if (PAIR == 2) umma_commit_pair(&bar); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (PAIR == 2 && warp == 0 && lane == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  if (PAIR == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { if (PAIR == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int PAIR>
void run(int N, int mode, int iters, long long* d_out) {
  cudaFuncSetAttribute(bench<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_out, 0, 148 * 8);
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench<PAIR>, N, mode, iters, d_out);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; int n = 0;
  for (int i = 0; i < 148; ++i) if (h[i] > 0) { s += h[i]; ++n; }
  const double clk = s / n / iters;
  printf("pair=%d N=%3d mode=%d: %.1f clk/MMA  (ideal %d) -> %.0f%% of tensor peak\n", PAIR, N, mode, clk, 128 * N / 256,
         100.0 * (128.0 * N / 256) / clk);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  const int iters = 4096;
  for (int mode : {0, 6})
    for (int N : {32, 64, 128, 160, 192, 256}) run<1>(N, mode, iters, d_out);
  return 0;
}
