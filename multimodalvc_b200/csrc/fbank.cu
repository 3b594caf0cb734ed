// Audio frontend of the AV-HuBERT hot path on sm_100a: int16 waveform -> 26-bin log mel filterbank
// (25 ms / 10 ms frames, pre-emphasis 0.97, rectangular window, 512-point FFT, natural log) -> 4-frame
// stacking to 104-dim rows -> audio/video length alignment -> per-row LayerNorm -> zero-padded collation
// + key-padding mask, in ONE kernel.  Also the RMS-matched additive-noise mixer that precedes it in the
// noisy-eval configuration.
//
// Replaces (CPU numpy, per sample in DataLoader workers in the reference):
//   python_speech_features.logfbank(wav, 16000)   called at avhubert/hubert_dataset.py:286
//   stacker(feats, 4)                              avhubert/hubert_dataset.py:259-274,287
//   audio/video length alignment                   avhubert/hubert_dataset.py:290-295
//   F.layer_norm(audio, audio.shape[1:])           avhubert/hubert_dataset.py:351-353
//   collater_audio zero-pad + padding_mask         avhubert/hubert_dataset.py:430-456
//   add_noise                                      avhubert/hubert_dataset.py:317-346
//
// The library computes the spectrum in float64 (numpy); so does this kernel: B200 has full-rate FP64
// units and the whole frontend is ~0.01 % of the encoder's time, so exactness is bought for free.
// One warp transforms one frame: the 400 real samples (zero-padded to 512) are packed into a 256-point
// complex sequence, transformed by a Stockham radix-8 x radix-8 x radix-4 FFT (one radix-8 butterfly per
// lane per pass, operands exchanged through shared memory), unpacked to the 257 real-FFT bins, squared,
// and reduced against the sparse triangular mel filters.  Four warps = the four frames of one stacked
// output row; warp 0 then normalises and stores the 104 floats with coalesced 16-byte stores.
#include "common.cuh"
#include "kernels.h"

#include <cmath>
#include <mutex>
#include <vector>

namespace avh {
namespace {

constexpr int NFFT = 512;
constexpr int NH = 256;          // complex FFT length
constexpr int NBINS = 257;
constexpr int NFILT = 26;
constexpr int FRAME_LEN = 400;
constexpr int FRAME_STEP = 160;
constexpr int STACK = 4;
constexpr int FEAT = NFILT * STACK;   // 104
constexpr double PREEMPH = 0.97;
constexpr double DBL_EPS = 2.220446049250313e-16;   // np.finfo(float).eps

struct FbankTables {
  double2 w256[NH];        // exp(-2 pi i m / 256)
  double2 w512[NBINS];     // exp(-2 pi i k / 512), k = 0..256
  double fb[NFILT][NBINS]; // triangular filters (python_speech_features.get_filterbanks defaults)
  int lo[NFILT], hi[NFILT];// non-zero support [lo, hi) of each filter
};
__device__ FbankTables g_tab;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward transform rotation by -90 degrees)
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }

// in-place forward 4-point DFT, natural order
__device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3) {
  const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  const double2 s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cadd(d02, d13);
  a3 = csub(d02, d13);
}
// in-place forward 8-point DFT, natural order
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
  constexpr double R = 0.70710678118654752440;
  // even / odd 4-point transforms
  double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  double2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  // twiddles W8^k: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
  o1 = make_double2(R * (o1.x + o1.y), R * (o1.y - o1.x));
  o2 = mul_mi(o2);
  o3 = make_double2(R * (o3.y - o3.x), -R * (o3.x + o3.y));
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

constexpr int WARPS = 4;
struct __align__(16) Smem {
  double2 w256[NH];
  double2 w512[NBINS + 1];
  double2 buf[WARPS][2][NH];     // per-warp ping-pong
  float row[FEAT];
  int any_frame;
};

__device__ __forceinline__ double sample_preemph(const int16_t* __restrict__ w, long long s, long long len) {
  // sigproc.preemphasis then zero padding (framesig): y[0]=x[0], y[s]=x[s]-0.97*x[s-1], y[s>=len]=0
  if (s >= len) return 0.0;
  const double x = (double)w[s];
  if (s == 0) return x;
  return __dsub_rn(x, __dmul_rn(PREEMPH, (double)w[s - 1]));
}

// One warp: log mel energies of frame `f` of a clip -> out26[0..25] (float32-rounded like .astype(float32)).
__device__ void frame_logfbank(const int16_t* __restrict__ wav, long long len, long long f, Smem& sm, int warp,
                               int lane, float* out26) {
  double2* b0 = sm.buf[warp][0];
  double2* b1 = sm.buf[warp][1];
  const long long base = f * FRAME_STEP;
  double2 v[8];
  // ---- pass 1: radix-8, Ns = 1 (no twiddles).  butterfly j = lane reads z[lane + 32 t]
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int n = lane + 32 * t;            // complex index; real samples 2n, 2n+1 of the frame
    double re = 0.0, im = 0.0;
    if (2 * n < FRAME_LEN) {
      re = sample_preemph(wav, base + 2 * n, len);
      im = sample_preemph(wav, base + 2 * n + 1, len);
    }
    v[t] = make_double2(re, im);
  }
  dft8(v);
#pragma unroll
  for (int t = 0; t < 8; ++t) b0[lane * 8 + t] = v[t];      // out[(j/1)*8 + 0 + t*1]
  __syncwarp();
  // ---- pass 2: radix-8, Ns = 8.  k = j % 8, twiddle W64^(k t) = W256^(4 k t)
  {
    const int j = lane, k = j & 7;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      double2 x = b0[j + 32 * t];
      if (t > 0) x = cmul(x, sm.w256[(4 * k * t) & 255]);
      v[t] = x;
    }
    dft8(v);
    const int ob = (j >> 3) * 64 + k;
#pragma unroll
    for (int t = 0; t < 8; ++t) b1[ob + t * 8] = v[t];
  }
  __syncwarp();
  // ---- pass 3: radix-4, Ns = 64.  64 butterflies, two per lane; twiddle W256^(k t)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;            // k = j % 64 = j
    double2 a0 = b1[j], a1 = b1[j + 64], a2 = b1[j + 128], a3 = b1[j + 192];
    a1 = cmul(a1, sm.w256[j]);
    a2 = cmul(a2, sm.w256[2 * j]);
    a3 = cmul(a3, sm.w256[(3 * j) & 255]);
    dft4(a0, a1, a2, a3);
    b0[j] = a0; b0[j + 64] = a1; b0[j + 128] = a2; b0[j + 192] = a3;
  }
  __syncwarp();
  // ---- real-FFT unpack + power spectrum: X[k] = E[k] + W512^k O[k], k = 0..256
  double* ps = reinterpret_cast<double*>(b1);
  for (int k = lane; k < NBINS; k += 32) {
    const double2 zk = b0[k & 255];
    const double2 zc = b0[(NH - k) & 255];
    const double2 e = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));
    const double2 o = make_double2(0.5 * (zk.y + zc.y), -0.5 * (zk.x - zc.x));   // (zk - conj(zc)) / (2i)
    const double2 x = cadd(e, cmul(sm.w512[k], o));
    ps[k] = (x.x * x.x + x.y * x.y) * (1.0 / NFFT);
  }
  __syncwarp();
  if (lane < NFILT) {
    double acc = 0.0;
    const int lo = g_tab.lo[lane], hi = g_tab.hi[lane];
    for (int i = lo; i < hi; ++i) acc += ps[i] * g_tab.fb[lane][i];
    if (acc == 0.0) acc = DBL_EPS;
    out26[lane] = (float)log(acc);
  }
  __syncwarp();
}

__host__ __device__ inline long long fbank_num_frames(long long n) {
  if (n <= FRAME_LEN) return 1;
  return 1 + (n - FRAME_LEN + FRAME_STEP - 1) / FRAME_STEP;
}

__global__ void __launch_bounds__(WARPS * 32)
fbank_kernel(FbankArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NH; i += blockDim.x) sm.w256[i] = g_tab.w256[i];
  for (int i = threadIdx.x; i < NBINS; i += blockDim.x) sm.w512[i] = g_tab.w512[i];
  __syncthreads();

  const long long rows = (long long)a.n_clips * a.T;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int clip = (int)(r / a.T);
    const int t = (int)(r - (long long)clip * a.T);
    const long long s0 = a.offsets[clip], len = a.offsets[clip + 1] - s0;
    const long long nfr = fbank_num_frames(len);
    const long long stacked = (nfr + STACK - 1) / STACK;
    const long long own = a.video_len != nullptr ? (long long)a.video_len[clip] : stacked;   // clip length in rows
    const bool padded = t >= own;              // collater zero padding -> mask True
    float* orow = a.out + r * FEAT;
    if (threadIdx.x == 0 && a.padding_mask != nullptr) a.padding_mask[r] = padded ? 1 : 0;
    if (padded || t >= stacked) {              // collation padding or alignment padding: a zero row
      if (threadIdx.x < FEAT / 4) reinterpret_cast<float4*>(orow)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;                                // uniform across the CTA
    }
    const long long f = (long long)t * STACK + warp;
    if (f < nfr) frame_logfbank(a.wav + s0, len, f, sm, warp, lane, sm.row + warp * NFILT);
    else if (lane < NFILT) sm.row[warp * NFILT + lane] = 0.f;     // stacker zero rows
    __syncthreads();
    if (warp == 0) {
      float v[4];
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < FEAT ? sm.row[c] : 0.f;
        s += (double)v[i];
      }
      if (a.normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double mean = s / FEAT;
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = lane + 32 * i;
          const double d = c < FEAT ? (double)v[i] - mean : 0.0;
          q += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const double rstd = 1.0 / sqrt(q / FEAT + 1e-5);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (float)(((double)v[i] - mean) * rstd);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < FEAT) orow[c] = v[i];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ noise mixer
// Every pass walks a clip in chunks of 8 consecutive samples per thread (one 16-byte load when the clip starts on an
// 8-sample boundary, as packed clips of whole video frames do) and tracks the position inside the tiled noise clip
// incrementally: ONE 64-bit modulo per chunk instead of one per sample (the per-sample form was ALU-bound at
// 0.5 TB/s: profiles/r2_ncu_hbm_kernels.txt).
template <typename F>
__device__ __forceinline__ void for_each_chunk8(const int16_t* __restrict__ clean, long long s0, long long len,
                                                const float* __restrict__ noise, long long noise_len, F&& body) {
  const bool aligned = ((s0 & 7) == 0) && ((reinterpret_cast<uintptr_t>(clean) & 15) == 0);
  const long long nchunk = (len + 7) >> 3;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (long long)gridDim.x * blockDim.x) {
    const long long i0 = c << 3;
    const int n = (int)(len - i0 < 8 ? len - i0 : 8);
    int16_t v[8];
    if (aligned && n == 8) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(clean + s0 + i0));
      *reinterpret_cast<uint4*>(v) = raw;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = k < n ? clean[s0 + i0 + k] : (int16_t)0;
    }
    float nz[8];
    long long j = i0 % noise_len;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      nz[k] = noise[j];
      if (++j == noise_len) j = 0;
    }
    body(i0, n, v, nz, aligned);
  }
}

// pass 1: per clip sum(clean^2), sum(noise^2) in float64.  scratch[4*clip + {0,1}]
__global__ void __launch_bounds__(256)
noise_stats_kernel(const int16_t* __restrict__ clean, const long long* __restrict__ offsets,
                   const float* __restrict__ noise, long long noise_len, double* scratch) {
  const int clip = blockIdx.y;
  const long long s0 = offsets[clip], len = offsets[clip + 1] - s0;
  double sc = 0.0, sn = 0.0;
  for_each_chunk8(clean, s0, len, noise, noise_len, [&](long long, int n, const int16_t* v, const float* nz, bool) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) {
        const float c = (float)v[k];
        sc += (double)(c * c);       // np.square on float32, then accumulated
        sn += (double)(nz[k] * nz[k]);
      }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&scratch[4 * clip + 0], sc);
    atomicAdd(&scratch[4 * clip + 1], sn);
  }
}
__device__ __forceinline__ float noise_scale(const double* scratch, int clip, long long len, float snr_db) {
  const float clean_rms = sqrtf((float)(scratch[4 * clip + 0] / (double)len));
  const float noise_rms = sqrtf((float)(scratch[4 * clip + 1] / (double)len));
  // adjusted_noise_rms = clean_rms / 10**(snr/20): python float power; float32 scalar / python float is a
  // float32 division under NumPy >= 2 promotion rules (weak python scalars); (adjusted / noise_rms) likewise
  const float adjusted = __fdiv_rn(clean_rms, (float)pow(10.0, (double)snr_db / 20.0));
  return __fdiv_rn(adjusted, noise_rms);
}
__device__ __forceinline__ unsigned long long f2ord(double v) {   // order-preserving map for atomicMax/Min
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2f(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}
// pass 2: per clip max / min of mixed = clean + noise * scale (float32).  scratch[4*clip + {2,3}] as ordered u64
__global__ void __launch_bounds__(256)
noise_minmax_kernel(const int16_t* __restrict__ clean, const long long* __restrict__ offsets,
                                    const float* __restrict__ noise, long long noise_len, float snr_db,
                                    double* scratch) {
  const int clip = blockIdx.y;
  const long long s0 = offsets[clip], len = offsets[clip + 1] - s0;
  const float scale = noise_scale(scratch, clip, len, snr_db);
  float mx = -INFINITY, mn = INFINITY;
  for_each_chunk8(clean, s0, len, noise, noise_len, [&](long long, int n, const int16_t* v, const float* nz, bool) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) {
        const float m = __fadd_rn((float)v[k], __fmul_rn(nz[k], scale));
        mx = fmaxf(mx, m);
        mn = fminf(mn, m);
      }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0 && mx >= mn) {
    unsigned long long* u = reinterpret_cast<unsigned long long*>(scratch);
    atomicMax(&u[4 * clip + 2], f2ord((double)mx));
    atomicMin(&u[4 * clip + 3], f2ord((double)mn));
  }
}
// pass 3: rescale if clipping, truncate toward zero to int16 (numpy astype)
__global__ void __launch_bounds__(256)
noise_mix_kernel(const int16_t* __restrict__ clean, const long long* __restrict__ offsets,
                                 const float* __restrict__ noise, long long noise_len, float snr_db,
                                 const double* __restrict__ scratch, int16_t* __restrict__ out) {
  const int clip = blockIdx.y;
  const long long s0 = offsets[clip], len = offsets[clip + 1] - s0;
  const float scale = noise_scale(scratch, clip, len, snr_db);
  const unsigned long long* u = reinterpret_cast<const unsigned long long*>(scratch);
  const float mx = (float)ord2f(u[4 * clip + 2]), mn = (float)ord2f(u[4 * clip + 3]);
  float rate = 1.f;
  bool rescale = false;
  if (mx > 32767.f || mn < -32768.f) {
    rescale = true;
    rate = (mx >= fabsf(mn)) ? (32767.f / mx) : (-32768.f / mn);
  }
  for_each_chunk8(clean, s0, len, noise, noise_len, [&](long long i0, int n, const int16_t* v, const float* nz, bool aligned) {
    int16_t o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float m = __fadd_rn((float)v[k], __fmul_rn(nz[k], scale));
      if (rescale) m = __fmul_rn(m, rate);
      m = fminf(fmaxf(m, -32768.f), 32767.f);      // guard; the rescale already bounds |m|
      o[k] = (int16_t)(int)m;                      // C cast truncates toward zero like numpy astype(int16)
    }
    if (aligned && n == 8 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      *reinterpret_cast<uint4*>(out + s0 + i0) = *reinterpret_cast<const uint4*>(o);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < n) out[s0 + i0 + k] = o[k];
    }
  });
}
__global__ void noise_init_kernel(double* scratch, int n_clips) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_clips) return;
  unsigned long long* u = reinterpret_cast<unsigned long long*>(scratch);
  scratch[4 * i + 0] = 0.0;
  scratch[4 * i + 1] = 0.0;
  u[4 * i + 2] = 0ull;                      // smallest ordered value
  u[4 * i + 3] = ~0ull;                     // largest ordered value
}

// ------------------------------------------------------------------------------------ host tables
double hz2mel(double hz) { return 2595.0 * std::log10(1.0 + hz / 700.0); }
double mel2hz(double mel) { return 700.0 * (std::pow(10.0, mel / 2595.0) - 1.0); }

int upload_tables() {
  static std::mutex mu;
  static std::vector<int> done_devices;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  AVH_CUDA_OK(cudaGetDevice(&dev));
  for (int d : done_devices)
    if (d == dev) return 0;
  std::vector<uint8_t> raw(sizeof(FbankTables), 0);
  FbankTables& t = *reinterpret_cast<FbankTables*>(raw.data());
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < NH; ++m) t.w256[m] = make_double2(std::cos(2.0 * PI * m / NH), -std::sin(2.0 * PI * m / NH));
  for (int k = 0; k < NBINS; ++k)
    t.w512[k] = make_double2(std::cos(2.0 * PI * k / NFFT), -std::sin(2.0 * PI * k / NFFT));
  // get_filterbanks(nfilt=26, nfft=512, samplerate=16000, lowfreq=0, highfreq=8000)
  double bin[NFILT + 2];
  const double lowmel = hz2mel(0.0), highmel = hz2mel(8000.0);
  for (int i = 0; i < NFILT + 2; ++i) {
    // numpy.linspace: start + i*step, last point set exactly to stop
    const double step = (highmel - lowmel) / (NFILT + 1);
    const double mel = (i == NFILT + 1) ? highmel : lowmel + i * step;
    bin[i] = std::floor((NFFT + 1) * mel2hz(mel) / 16000.0);
  }
  for (int j = 0; j < NFILT; ++j) {
    for (int i = (int)bin[j]; i < (int)bin[j + 1]; ++i) t.fb[j][i] = (i - bin[j]) / (bin[j + 1] - bin[j]);
    for (int i = (int)bin[j + 1]; i < (int)bin[j + 2]; ++i) t.fb[j][i] = (bin[j + 2] - i) / (bin[j + 2] - bin[j + 1]);
    t.lo[j] = (int)bin[j];
    t.hi[j] = (int)bin[j + 2];
  }
  AVH_CUDA_OK(cudaMemcpyToSymbol(g_tab, raw.data(), sizeof(FbankTables)));
  done_devices.push_back(dev);
  return 0;
}

}  // namespace

int launch_fbank(const FbankArgs& a, cudaStream_t stream) {
  AVH_CHECK(a.n_clips >= 0 && a.T >= 0, "negative sizes");
  const long long rows = (long long)a.n_clips * a.T;
  if (rows == 0) return 0;
  AVH_CHECK(a.wav != nullptr && a.offsets != nullptr && a.out != nullptr, "null pointer");
  if (upload_tables()) return 1;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(fbank_kernel), (int)sizeof(Smem))) return 1;
  const int sms = device_sm_count();
  const long long want = rows < (long long)sms * 5 ? rows : (long long)sms * 5;
  fbank_kernel<<<(int)want, WARPS * 32, sizeof(Smem), stream>>>(a);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return 0;
}

int launch_add_noise(const int16_t* clean, const long long* offsets, int n_clips, const float* noise,
                     long long noise_len, float snr_db, int16_t* out, double* scratch, cudaStream_t stream) {
  if (n_clips <= 0) return 0;
  AVH_CHECK(clean && offsets && noise && out && scratch, "null pointer");
  AVH_CHECK(noise_len > 0, "empty noise clip");
  noise_init_kernel<<<(n_clips + 127) / 128, 128, 0, stream>>>(scratch, n_clips);
  dim3 grid(64, n_clips);
  noise_stats_kernel<<<grid, 256, 0, stream>>>(clean, offsets, noise, noise_len, scratch);
  noise_minmax_kernel<<<grid, 256, 0, stream>>>(clean, offsets, noise, noise_len, snr_db, scratch);
  noise_mix_kernel<<<grid, 256, 0, stream>>>(clean, offsets, noise, noise_len, snr_db, scratch, out);
  AVH_CUDA_OK(cudaGetLastError());
  count_launch(4);
  return 0;
}

}  // namespace avh
