// fft.cu — batched multi-sample-rate STFT / iSTFT for sm_100a.
//
// One CTA owns a run of consecutive frames of one utterance.  Frames are windowed while they are staged into
// shared memory, transformed there by a Stockham mixed-radix FFT (radices 4/2/3/5/7, generic odd primes for the
// FlowSE 22.05/44.1 kHz sizes 705 = 3*5*47 and 1411 = 17*83), and written back with coalesced float2 stores in the
// (B,T,F,2) layout BandSplit reads.  The inverse kernel fuses the complex mask (s = m*x + r), the optional inverse
// spectral compression, the synthesis window, the overlap-add and the window-envelope division, so the mask and
// the residual are never re-read and the masked spectrum is written exactly once.
//
// Replaces: espnet2 Stft.forward / Stft.inverse (torch.stft / torch.istft), called from
// baseline_code/models/bsrnn.py:37,40 and baseline_code/flow_model.py:136,145.
#include "common.cuh"

#include "fft_core.cuh"

namespace bsrnn {

__global__ void twiddle_kernel(float2* tw, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double s, c;
    sincospi(-2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
}

constexpr int kStftThreads = 256;

// grid (ceil(T/FPB), B).  ADJ = false: the STFT encoder.  ADJ = true: the ADJOINT of istft_kernel (training backward of
// the iSTFT decoder, replaces autograd through torch.istft at d_model.py:71-74): `wav` is dL/d(enhanced waveform) of L
// samples; frame t is w[n] * g[t*hop + n - N/2] / env (zero outside [0, L): trimmed samples carry no gradient; env = the
// window-square envelope the forward divides by), and bin k of its DFT is scaled by c_k / N (c_0 = c_{N/2} = 1, else 2;
// the imaginary parts of DC / Nyquist are ignored by the forward, so their gradient is 0).
// Two real frames share ONE complex FFT (z = x_a + j x_b;  X_a[k] = (Z[k] + conj Z[N-k]) / 2,  X_b[k] = (Z[k] - conj
// Z[N-k]) / (2j)): half the transforms of the one-frame-per-FFT version for any N, odd sizes included.
// band_stats != nullptr (encoder only): per (utterance, band) sum and sum of squares of the OUTPUT spectrum (after the
// optional compression), accumulated in double -- the reduction half of BandSplit's GroupNorm(1, 2 s_k) over (2 s_k x T)
// [reference bsrnn_flowse.py:72-73], which otherwise costs a second pass over the spectrum (bsrnn_band_stats).
template <bool ADJ>
__global__ void __launch_bounds__(kStftThreads)
stft_kernel(const float* __restrict__ wav, const int32_t* __restrict__ lens, float2* __restrict__ spec,
            const float2* __restrict__ twiddle, FftPlan plan, int L, int T, int hop, int fpb, int transform,
            float exponent, float factor, double* __restrict__ band_stats, const int* __restrict__ band_bin0, int n_bands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = plan.n;
  const int F = N / 2 + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* buf0 = tw + N;
  const int ppb = (fpb + 1) >> 1;                     // frame PAIRS per block
  float2* buf1 = buf0 + (size_t)ppb * N + fpb;       // + fpb: a buffer also holds fpb one-sided spectra (fpb * (N/2 + 1))
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * fpb;
  const int nfr = min(fpb, T - t0);
  const int len_b = lens ? lens[b] : L;
  const int olen = ADJ ? T : (len_b + 2 * (N / 2) - N) / hop + 1;
  // live frames of this CTA form a prefix [t0, t0+nlive)
  const int nlive = max(0, min(nfr, olen - t0));
  const int npair = (nlive + 1) >> 1;

  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = twiddle[i];
  __syncthreads();
  const float* row = wav + (size_t)b * L;
  auto sample = [&](int f, int n) -> float {         // windowed sample n of live frame f
    int i = (t0 + f) * hop - N / 2 + n;
    const float w = 0.5f - 0.5f * tw[n].x;            // periodic Hann: cos(2*pi*n/N) = Re tw[n]
    if (ADJ) {
      if (i < 0 || i >= L) return 0.f;
      const long ip = (long)i + N / 2;                // position in the padded (un-trimmed) overlap-add buffer
      long a = ip - N + 1;
      a = a <= 0 ? 0 : (a + hop - 1) / hop;
      long z = ip / hop;
      if (z > T - 1) z = T - 1;
      float env = 0.f;
      for (long t = a; t <= z; ++t) {
        const float we = 0.5f - 0.5f * tw[ip - t * hop].x;
        env += we * we;
      }
      return env > 1e-11f ? row[i] * w / env : 0.f;
    }
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return row[i] * w;
  };
  if constexpr (!ADJ) {
    // sample n of every live frame at once: up to 8 independent global loads in flight per thread and no division (the
    // one-element-at-a-time loop spent 45 % of the kernel's samples waiting on its single load, profiles/r02 call49)
    const long base0 = (long)t0 * hop - N / 2;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float w = 0.5f - 0.5f * tw[n].x;            // periodic Hann: cos(2*pi*n/N) = Re tw[n]
      float v[8];
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        v[f] = 0.f;
        if (f < nlive) {
          long i = base0 + (long)f * hop + n;
          if (i < 0) i = -i;
          if (i >= L) i = 2 * ((long)L - 1) - i;
          v[f] = __ldg(row + i);
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < npair) buf0[(size_t)p * N + n] = make_float2(v[2 * p] * w, v[2 * p + 1] * w);
    }
  } else {
  for (int idx = threadIdx.x; idx < npair * N; idx += blockDim.x) {
    const int p = idx / N;
    const int n = idx - p * N;
    const float xa = sample(2 * p, n);
    const float xb = (2 * p + 1 < nlive) ? sample(2 * p + 1, n) : 0.f;
    buf0[idx] = make_float2(xa, xb);
  }
  }
  __syncthreads();
  float2* res = buf0;
  if (npair > 0) res = fft_frames<false>(buf0, buf1, tw, plan, npair);
  float2* outb = (res == buf0) ? buf1 : buf0;        // the finished spectra of the block's frames (for the statistics)
  for (int f = 0; f < nfr; ++f)
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (f < nlive) {
      const float2 zk = res[(size_t)(f >> 1) * N + k];
      const float2 zc = res[(size_t)(f >> 1) * N + (k == 0 ? 0 : N - k)];
      v = (f & 1) ? make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x))
                  : make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      if (ADJ) {
        const bool edge = k == 0 || (N % 2 == 0 && k == N / 2);
        const float c = (edge ? 1.0f : 2.0f) / (float)N;
        v.x *= c;
        v.y = edge ? 0.f : v.y * c;
      }
      if (transform == 1) {
        const float mag = sqrtf(v.x * v.x + v.y * v.y);
        const float sc = mag > 0.f ? __powf(mag, exponent - 1.0f) * factor : 0.f;
        v.x *= sc; v.y *= sc;
      }
    }
    spec[((size_t)b * T + t0 + f) * F + k] = v;
    if (!ADJ && band_stats) outb[(size_t)f * F + k] = v;       // fpb * F <= ppb * N + fpb float2 slots
  }
  if (!ADJ && band_stats) {
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int band = warp; band < n_bands; band += kStftThreads / 32) {
      const int k0 = band_bin0[band], k1 = min(band_bin0[band + 1], F);
      if (k0 >= F) break;
      const int wdt = k1 - k0;
      double s1 = 0.0, s2 = 0.0;
      for (int e = lane; e < nlive * wdt; e += 32) {
        const int f = e / wdt;
        const float2 v = outb[(size_t)f * F + k0 + (e - f * wdt)];
        s1 += (double)v.x + (double)v.y;
        s2 += (double)v.x * v.x + (double)v.y * v.y;
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0 && nlive > 0) {
        atomicAdd(band_stats + ((size_t)b * n_bands + band) * 2, s1);
        atomicAdd(band_stats + ((size_t)b * n_bands + band) * 2 + 1, s2);
      }
    }
  }
}

// grid (nblocks, B); block c covers padded output samples [c*G*hop, (c+1)*G*hop) and owns frames [c*G, (c+1)*G).
__global__ void __launch_bounds__(kStftThreads)
istft_kernel(const float2* __restrict__ spec, const float2* __restrict__ mask, const float2* __restrict__ resid,
             float2* __restrict__ spec_out, float* __restrict__ wav_out, const float2* __restrict__ twiddle,
             FftPlan plan, int T, int L_out, int hop, int fpb, int G, int transform, float inv_exponent,
             float inv_factor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = plan.n;
  const int F = N / 2 + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* buf0 = tw + N;
  const int ppb = (fpb + 1) >> 1;                     // frame PAIRS per batch: two Hermitian spectra -> one complex inverse
  float2* buf1 = buf0 + (size_t)ppb * N;              // FFT (Z = X_a + j X_b;  x_a = Re z, x_b = Im z)
  float* acc = reinterpret_cast<float*>(buf1 + (size_t)ppb * N);
  const int b = blockIdx.y;
  const int span = G * hop;
  const long p0 = (long)blockIdx.x * span;
  const long p1 = p0 + span;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = twiddle[i];
  for (int i = threadIdx.x; i < span; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  // frames touching [p0, p1): t*hop + N > p0  and  t*hop < p1
  long tlo = (p0 - N + 1 + hop - 1);
  tlo = tlo <= 0 ? 0 : tlo / hop;
  long thi = (p1 - 1) / hop;
  const int own_lo = blockIdx.x * G, own_hi = own_lo + G;     // frames whose masked spectrum this CTA writes
  if (thi < own_hi - 1) thi = own_hi - 1;
  if (thi > T - 1) thi = T - 1;
  const float inv_n = 1.0f / (float)N;
  const bool nyq = (N % 2 == 0);

  for (long tb = tlo; tb <= thi; tb += fpb) {
    const int nfr = (int)min((long)fpb, thi - tb + 1);
    // ---- stage the masked spectra of a frame pair as ONE Hermitian-combined complex spectrum
    const int npair = (nfr + 1) >> 1;
    auto masked = [&](int f, int k) -> float2 {
      const int t = (int)tb + f;
      const size_t g = ((size_t)b * T + t) * F + k;
      float2 v = spec[g];
      if (mask) {
        const float2 m = mask[g], r = resid[g];
        v = make_float2(m.x * v.x - m.y * v.y + r.x, m.x * v.y + m.y * v.x + r.y);
      }
      if (spec_out && t >= own_lo && t < own_hi) spec_out[g] = v;
      if (transform == 1) {
        v.x *= inv_factor; v.y *= inv_factor;
        const float mag = sqrtf(v.x * v.x + v.y * v.y);
        const float sc = mag > 0.f ? __powf(mag, inv_exponent - 1.0f) : 0.f;
        v.x *= sc; v.y *= sc;
      }
      if (k == 0 || (nyq && k == N / 2)) v.y = 0.f;   // irfft ignores Im of DC / Nyquist
      return v;
    };
    // bin k of every frame of the batch at once (fpb <= 8): the loads of all frames are independent and in flight together
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      float2 v[8];
#pragma unroll
      for (int f = 0; f < 8; ++f) v[f] = (f < nfr) ? masked(f, k) : make_float2(0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (p < npair) {
          const float2 va = v[2 * p], vb = v[2 * p + 1];
          float2* fr = buf0 + (size_t)p * N;
          fr[k] = make_float2(va.x - vb.y, va.y + vb.x);
          if (k != 0 && !(nyq && k == N / 2)) fr[N - k] = make_float2(va.x + vb.y, vb.x - va.y);
        }
      }
    }
    __syncthreads();
    float2* res = fft_frames<true>(buf0, buf1, tw, plan, npair);
    // ---- synthesis window + overlap-add; frames of a batch overlap, so they are added one after another
    for (int f = 0; f < nfr; ++f) {
      const long base = (tb + f) * hop;
      const float2* fr = res + (size_t)(f >> 1) * N;
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const long i = base + n;
        if (i >= p0 && i < p1) {
          const float w = 0.5f - 0.5f * tw[n].x;
          const float2 z = fr[n];
          acc[i - p0] += ((f & 1) ? z.y : z.x) * inv_n * w;
        }
      }
      __syncthreads();
    }
  }
  // ---- envelope division, trim, store
  const long half = N / 2;
  for (int o = threadIdx.x; o < span; o += blockDim.x) {
    const long i = p0 + o;
    const long i_out = i - half;
    if (i_out < 0 || i_out >= L_out) continue;
    long a = i - N + 1;
    a = a <= 0 ? 0 : (a + hop - 1) / hop;
    long z = i / hop;
    if (z > T - 1) z = T - 1;
    float env = 0.f;
    for (long t = a; t <= z; ++t) {
      const float w = 0.5f - 0.5f * tw[i - t * hop].x;
      env += w * w;
    }
    wav_out[(size_t)b * L_out + i_out] = env > 1e-11f ? acc[o] / env : 0.f;
  }
}

}  // namespace bsrnn

using namespace bsrnn;

extern "C" int bsrnn_fft_twiddle(float* twiddle, int n_fft, void* stream) {
  BSRNN_CHECK_ARG(twiddle && n_fft >= 2, "fft_twiddle: bad arguments");
  twiddle_kernel<<<cdiv(n_fft, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(twiddle), n_fft);
  BSRNN_LAUNCH_OK();
  return 0;
}

// one CTA = 8 frames = 4 frame pairs: shared memory = twiddles + two ping-pong buffers of 4 complex frames
static int stft_smem_fpb(int n_fft, int* fpb_out) {
  int fpb = 8;
  while (fpb > 2 && (size_t)n_fft * 8 * (1 + 2 * ((fpb + 1) / 2)) > 100 * 1024) fpb >>= 1;
  *fpb_out = fpb;
  return (int)((size_t)n_fft * 8 * (1 + 2 * ((fpb + 1) / 2)) + 2 * (size_t)fpb * 8);
}
static int launch_stft(const float* wav, const int32_t* lens, float* spec, const float* twiddle, const FftPlan& plan, int B,
                       int L, int T, int hop, int transform, float exponent, float factor, double* band_stats,
                       const int32_t* band_bin0, int n_bands, void* stream) {
  int fpb = 8;
  const int smem = stft_smem_fpb(plan.n, &fpb);
  BSRNN_CUDA_OK(cudaFuncSetAttribute(stft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(cdiv(T, fpb), B);
  stft_kernel<false><<<grid, kStftThreads, smem, (cudaStream_t)stream>>>(
      wav, lens, reinterpret_cast<float2*>(spec), reinterpret_cast<const float2*>(twiddle), plan, L, T, hop, fpb, transform,
      exponent, factor, band_stats, band_bin0, n_bands);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_stft_fwd(const float* wav, const int32_t* lens, float* spec, const float* twiddle, int B,
                              int L, int n_fft, int hop, int transform, float exponent, float factor,
                              void* stream) {
  BSRNN_CHECK_ARG(wav && spec && twiddle, "stft_fwd: null pointer");
  BSRNN_CHECK_ARG(B > 0 && n_fft >= 4 && hop > 0 && L > n_fft / 2, "stft_fwd: bad dims B=%d L=%d n_fft=%d hop=%d", B, L,
                  n_fft, hop);
  FftPlan plan;
  BSRNN_CHECK_ARG(make_plan(n_fft, &plan), "stft_fwd: cannot factorise n_fft=%d", n_fft);
  const int T = 1 + L / hop;
  return launch_stft(wav, lens, spec, twiddle, plan, B, L, T, hop, transform, exponent, factor, nullptr, nullptr, 0, stream);
}

extern "C" int bsrnn_stft_stats_fwd(const float* wav, const int32_t* lens, float* spec, const float* twiddle, int B, int L,
                                    int n_fft, int hop, int transform, float exponent, float factor, double* band_stats,
                                    const int32_t* band_bin0, int n_bands, void* stream) {
  BSRNN_CHECK_ARG(wav && spec && twiddle && band_stats && band_bin0 && n_bands > 0, "stft_stats_fwd: null pointer");
  BSRNN_CHECK_ARG(B > 0 && n_fft >= 4 && hop > 0 && L > n_fft / 2, "stft_stats_fwd: bad dims B=%d L=%d n_fft=%d hop=%d", B, L,
                  n_fft, hop);
  FftPlan plan;
  BSRNN_CHECK_ARG(make_plan(n_fft, &plan), "stft_stats_fwd: cannot factorise n_fft=%d", n_fft);
  BSRNN_CUDA_OK(cudaMemsetAsync(band_stats, 0, (size_t)B * n_bands * 2 * sizeof(double), (cudaStream_t)stream));
  return launch_stft(wav, lens, spec, twiddle, plan, B, L, 1 + L / hop, hop, transform, exponent, factor, band_stats, band_bin0,
                     n_bands, stream);
}

extern "C" int bsrnn_istft_fwd(const float* spec, const float* mask, const float* resid, float* spec_out,
                               float* wav_out, const float* twiddle, int B, int T, int L_out, int n_fft, int hop,
                               int transform, float exponent, float factor, void* stream) {
  BSRNN_CHECK_ARG(spec && wav_out && twiddle, "istft_fwd: null pointer");
  BSRNN_CHECK_ARG((mask == nullptr) == (resid == nullptr), "istft_fwd: mask and resid must come together");
  BSRNN_CHECK_ARG(B > 0 && T > 0 && L_out > 0 && n_fft >= 4 && hop > 0, "istft_fwd: bad dims");
  FftPlan plan;
  BSRNN_CHECK_ARG(make_plan(n_fft, &plan), "istft_fwd: cannot factorise n_fft=%d", n_fft);
  int G = 8;
  int fpb = 8;                                        // frames per batch = 2 x frame pairs
  while (fpb > 2 && (size_t)n_fft * 8 * (1 + 2 * ((fpb + 1) / 2)) + (size_t)G * hop * 4 > 100 * 1024) fpb >>= 1;
  // A block of G hops is touched by G + ceil(N / hop) - 1 frames.  When that fits ONE batch, take G = fpb - overlap: with
  // G = fpb = 8 at N = 2 hop every block ran a second batch for its 9th frame (one frame pair at the latency of four).
  const int overlap = (n_fft + hop - 1) / hop - 1;
  if (fpb - overlap >= 4) G = fpb - overlap;
  const size_t smem = (size_t)n_fft * 8 * (1 + 2 * ((fpb + 1) / 2)) + (size_t)G * hop * 4;
  BSRNN_CUDA_OK(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long span = (long)G * hop;
  int nblocks = cdiv(n_fft / 2 + (long)L_out, span);
  if (nblocks < cdiv(T, G)) nblocks = cdiv(T, G);
  dim3 grid(nblocks, B);
  istft_kernel<<<grid, kStftThreads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(spec), reinterpret_cast<const float2*>(mask),
      reinterpret_cast<const float2*>(resid), reinterpret_cast<float2*>(spec_out), wav_out,
      reinterpret_cast<const float2*>(twiddle), plan, T, L_out, hop, fpb, G, transform,
      transform == 1 ? 1.0f / exponent : 1.0f, transform == 1 ? 1.0f / factor : 1.0f);
  BSRNN_LAUNCH_OK();
  return 0;
}

// Backward of bsrnn_istft_fwd without mask / transform (training): d_wav (B, L_out) -> d_spec (B, T, F, 2).
extern "C" int bsrnn_istft_bwd(const float* d_wav, float* d_spec, const float* twiddle, int B, int T, int L_out, int n_fft,
                               int hop, void* stream) {
  BSRNN_CHECK_ARG(d_wav && d_spec && twiddle, "istft_bwd: null pointer");
  BSRNN_CHECK_ARG(B > 0 && T > 0 && L_out > 0 && n_fft >= 4 && hop > 0, "istft_bwd: bad dims");
  FftPlan plan;
  BSRNN_CHECK_ARG(make_plan(n_fft, &plan), "istft_bwd: cannot factorise n_fft=%d", n_fft);
  int fpb = 8;
  const int smem = stft_smem_fpb(n_fft, &fpb);
  BSRNN_CUDA_OK(cudaFuncSetAttribute(stft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(cdiv(T, fpb), B);
  stft_kernel<true><<<grid, kStftThreads, smem, (cudaStream_t)stream>>>(
      d_wav, nullptr, reinterpret_cast<float2*>(d_spec), reinterpret_cast<const float2*>(twiddle), plan, L_out, T, hop, fpb,
      0, 1.0f, 1.0f, nullptr, nullptr, 0);
  BSRNN_LAUNCH_OK();
  return 0;
}
