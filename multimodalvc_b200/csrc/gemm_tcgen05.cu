// tcgen05 / TMEM / TMA GEMM for sm_100a (B200): C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 in, fp32 accumulate.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one thread issues
// tcgen05.mma, accumulators live in TMEM, two accumulator stages so the epilogue of tile i overlaps the
// main loop of tile i+1), warp 2 = TMEM allocator, warps 4..7 = epilogue (tcgen05.ld -> fused
// scale/bias/activation/residual/mask -> global).  A and B tiles are K-major, 128-byte swizzled, moved
// by TMA with out-of-bounds zero fill; the K loop walks a table of (A column, A row shift, B column,
// B row shift) so that one kernel serves plain Linear layers, shifted-row implicit-GEMM convolutions
// and split-precision products.
//
// Replaces on the reference path (all cuBLAS/cuDNN library calls there): nn.Linear in
// avhubert/hubert.py:321,360-364, fairseq/fairseq/models/wav2vec/wav2vec2.py:955-956,
// multihead_attention.py:64-77; nn.Conv1d pos_conv wav2vec2.py:822-835; nn.Conv2d avhubert/resnet.py:15-24.
#include "common.cuh"
#include "gemm.h"

namespace avh {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_THREADS = 256;

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int KTABLE_BYTES = GEMM_MAX_KSTEPS * 16;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + KTABLE_BYTES + BAR_BYTES;
};

struct KernelParams {
  long long M;
  int N;
  int num_kb;
  int num_m_blk, num_n_blk;
  int a_col_per_nblk;
  const int* a_col_nblk;
  const int4* ktable;
  Epilogue ep;
};

__device__ __forceinline__ void load8(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const KernelParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
  int4* ktab = reinterpret_cast<int4*>(smem_b + C::STAGES * C::B_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ktab) + C::KTABLE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blk * p.num_n_blk;

  if (p.ktable != nullptr) {
    for (int i = threadIdx.x; i < p.num_kb; i += NUM_THREADS) ktab[i] = __ldg(p.ktable + i);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blk;
        const int n_blk = tile / p.num_m_blk;
        const int m0 = m_blk * BM;
        const int n0 = n_blk * BN;
        const int a_col_base = p.a_col_nblk != nullptr ? __ldg(p.a_col_nblk + n_blk) : n_blk * p.a_col_per_nblk;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          int4 e;
          if (p.ktable != nullptr) e = ktab[kb];
          else e = make_int4(kb * BK, 0, kb * BK, 0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + C::B_STAGE_BYTES);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tma_a, &full_bar[stage], e.x + a_col_base, m0 + e.y);
          tma_load_2d(smem_b + stage * C::B_STAGE_BYTES, &tma_b, &full_bar[stage], e.z, n0 + e.w);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * C::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 (32 B) along K inside the 128-byte swizzle atom: +2 in 16-byte units
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);        // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const Epilogue& ep = p.ep;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % p.num_m_blk;
      const int n_blk = tile / p.num_m_blk;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long r = (long long)m_blk * BM + q * 32 + lane;
      bool store = r < p.M;
      bool zero = false;
      long long orow = r;
      if (ep.map_mode == MAP_2LEVEL) {
        const long long n = r / ep.S2;
        const int rem = (int)(r - n * ep.S2);
        const int h = rem / ep.S1;
        const int w = rem - h * ep.S1;
        orow = n * ep.O2 + (long long)h * ep.O1 + w + ep.O0;
        if (!(h < ep.H && w < ep.W)) {
          if (ep.invalid_zero) zero = true; else store = false;
        }
      }
      if (store && ep.row_zero != nullptr && ep.row_zero[orow]) zero = true;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        __syncwarp();
        uint32_t raw[32];
        tmem_ld_32x32(taddr + ch * 32, raw);
        tmem_ld_wait();
        if (ch == BN / 32 - 1) {
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);     // accumulator stage drained
        }
        const int col0 = n_blk * BN + ch * 32;
        if (store && col0 < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (zero) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
          float t[32];
          if (ep.col_scale != nullptr) {
            load8(ep.col_scale + col0, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= t[j];
          }
          if (ep.col_bias != nullptr) {
            load8(ep.col_bias + col0, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += t[j];
          }
          if (ep.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (ep.act == ACT_PRELU) {
            load8(ep.slope1 + col0, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * t[j];
          }
          if (ep.R != nullptr) {
            if (ep.r_fp32) {
              load8(reinterpret_cast<const float*>(ep.R) + orow * ep.ldr + col0, t);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += t[j];
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>(
                  reinterpret_cast<const __nv_bfloat16*>(ep.R) + orow * ep.ldr + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u = __ldg(rp + i);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 f = __bfloat1622float2(h2[j]);
                  v[8 * i + 2 * j] += f.x;
                  v[8 * i + 2 * j + 1] += f.y;
                }
              }
            }
          }
          if (ep.slope2 != nullptr) {
            load8(ep.slope2 + col0, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * t[j];
          }
        }
        if (ep.c_fp32) {
          float4* cp = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.C) + orow * ep.ldc + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) cp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          uint4* cp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.C) + orow * ep.ldc + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
            cp[i] = u;
          }
        }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] (row stride ld elements), box = 64 cols x box_rows, 128B swizzle.
int encode_2d(CUtensorMap* map, const void* base, long long rows, int cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  AVH_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  AVH_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  AVH_CHECK((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return 0;
}

template <int BN>
int launch_t(const GemmPlan& plan, const KernelParams& kp, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AVH_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)Cfg<BN>::SMEM));
    configured = true;
  }
  gemm_kernel<BN><<<plan.grid, NUM_THREADS, Cfg<BN>::SMEM, stream>>>(plan.tma_a, plan.tma_b, kp);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace

int gemm_pick_block_n(long long M, int N) {
  // Prefer the widest tile that still yields at least ~one wave of CTAs (148 SMs).
  const long long mt = (M + BM - 1) / BM;
  const int sms = device_sm_count();
  for (int bn : {256, 128}) {
    if (N % bn == 0 || N > 4 * bn) {
      long long tiles = mt * ((N + bn - 1) / bn);
      if (tiles >= sms) return bn;
    }
  }
  return (N % 128 == 0 && mt * (N / 128) >= sms / 2) ? 128 : 64;
}

int gemm_plan(const GemmProblem& pr, GemmPlan* plan) {
  AVH_CHECK(pr.block_n == 64 || pr.block_n == 128 || pr.block_n == 256, "block_n must be 64/128/256");
  AVH_CHECK(pr.N % 32 == 0, "N must be a multiple of 32");
  AVH_CHECK(pr.num_kb >= 1, "num_kb must be >= 1");
  AVH_CHECK(pr.ktable == nullptr || pr.num_kb <= GEMM_MAX_KSTEPS, "too many K steps for the smem table");
  AVH_CHECK(pr.M >= 1 && pr.M < (1ll << 31) - BM, "M out of range");
  AVH_CHECK(pr.ep.C != nullptr, "output pointer is null");
  plan->prob = pr;
  if (encode_2d(&plan->tma_a, pr.A, pr.a_rows, pr.a_cols, pr.lda, BM)) return 1;
  if (encode_2d(&plan->tma_b, pr.B, pr.b_rows, pr.b_cols, pr.ldb, pr.block_n)) return 1;
  const long long mt = (pr.M + BM - 1) / BM;
  const long long nt = (pr.N + pr.block_n - 1) / pr.block_n;
  const long long tiles = mt * nt;
  const int sms = device_sm_count();
  plan->grid = (int)(tiles < sms ? tiles : sms);
  plan->smem = pr.block_n == 256 ? Cfg<256>::SMEM : pr.block_n == 128 ? Cfg<128>::SMEM : Cfg<64>::SMEM;
  return 0;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  const GemmProblem& pr = plan.prob;
  KernelParams kp;
  kp.M = pr.M;
  kp.N = pr.N;
  kp.num_kb = pr.num_kb;
  kp.num_m_blk = (int)((pr.M + BM - 1) / BM);
  kp.num_n_blk = (pr.N + pr.block_n - 1) / pr.block_n;
  kp.a_col_per_nblk = pr.a_col_per_nblk;
  kp.a_col_nblk = pr.a_col_nblk;
  kp.ktable = reinterpret_cast<const int4*>(pr.ktable);
  kp.ep = pr.ep;
  switch (pr.block_n) {
    case 256: return launch_t<256>(plan, kp, stream);
    case 128: return launch_t<128>(plan, kp, stream);
    default: return launch_t<64>(plan, kp, stream);
  }
}

}  // namespace avh
