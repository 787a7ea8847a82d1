// fft_core.cuh — shared-memory Stockham mixed-radix FFT used by the STFT / iSTFT kernels (fft.cu) and the
// multi-resolution spectral loss kernel (loss.cu).  Radices 4/2/3/5/7 are unrolled; other odd primes (FlowSE 22.05 /
// 44.1 kHz sizes 705 = 3*5*47, 1411 = 17*83) take the generic butterfly.
#pragma once
#include "common.cuh"

namespace bsrnn {

constexpr int kMaxStages = 16;
constexpr int kMaxGenericRadix = 96;

struct FftPlan {
  int n;
  int nstages;
  int radix[kMaxStages];
};

static inline bool make_plan(int n, FftPlan* p) {
  p->n = n;
  p->nstages = 0;
  int m = n;
  const int pref[] = {4, 2, 3, 5, 7};
  for (int r : pref) {
    while (m % r == 0) {
      if (p->nstages >= kMaxStages) return false;
      p->radix[p->nstages++] = r;
      m /= r;
    }
  }
  for (int r = 11; m > 1; r += 2) {
    while (m % r == 0) {
      if (p->nstages >= kMaxStages || r > kMaxGenericRadix) return false;
      p->radix[p->nstages++] = r;
      m /= r;
    }
    if (r > kMaxGenericRadix && m > 1) return false;
  }
  return true;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// tw[i] = exp(-2*pi*i/N); inverse transform uses the conjugate.
template <bool INV>
__device__ __forceinline__ float2 twd(const float2* tw, int i) {
  float2 w = tw[i];
  if (INV) w.y = -w.y;
  return w;
}

// One Stockham butterfly of radix R: element index j in [0, N/R).
template <int R, bool INV>
__device__ __forceinline__ void butterfly(const float2* __restrict__ in, float2* __restrict__ out,
                                          const float2* __restrict__ tw, int N, int Ns, int j) {
  const int k = j % Ns;
  const int nr = N / R;
  const int tstep = N / (Ns * R);
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    v[q] = in[j + q * nr];
    if (q > 0 && k > 0) v[q] = cmul(v[q], twd<INV>(tw, q * k * tstep));
  }
  float2 o[R];
  if (R == 2) {
    o[0] = make_float2(v[0].x + v[1].x, v[0].y + v[1].y);
    o[1] = make_float2(v[0].x - v[1].x, v[0].y - v[1].y);
  } else if (R == 4) {
    float2 a = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
    float2 b = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
    float2 c = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
    float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
    // forward: multiply d by -i ; inverse: by +i
    float2 dj = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    o[0] = make_float2(a.x + c.x, a.y + c.y);
    o[1] = make_float2(b.x + dj.x, b.y + dj.y);
    o[2] = make_float2(a.x - c.x, a.y - c.y);
    o[3] = make_float2(b.x - dj.x, b.y - dj.y);
  } else {
#pragma unroll
    for (int p = 0; p < R; ++p) {
      float2 acc = v[0];
#pragma unroll
      for (int q = 1; q < R; ++q) {
        float2 w = twd<INV>(tw, ((p * q) % R) * nr);
        acc.x += v[q].x * w.x - v[q].y * w.y;
        acc.y += v[q].x * w.y + v[q].y * w.x;
      }
      o[p] = acc;
    }
  }
  const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) out[j0 + q * Ns] = o[q];
}

// Generic (runtime) radix for large odd primes; rare path (FlowSE at 22.05 / 44.1 kHz).
template <bool INV>
__device__ void butterfly_generic(const float2* __restrict__ in, float2* __restrict__ out,
                                  const float2* __restrict__ tw, int N, int Ns, int R, int j) {
  const int k = j % Ns;
  const int nr = N / R;
  const int tstep = N / (Ns * R);
  float2 v[kMaxGenericRadix];
  for (int q = 0; q < R; ++q) {
    v[q] = in[j + q * nr];
    if (q > 0 && k > 0) v[q] = cmul(v[q], twd<INV>(tw, q * k * tstep));
  }
  const int j0 = (j / Ns) * Ns * R + k;
  for (int p = 0; p < R; ++p) {
    float2 acc = v[0];
    int pq = 0;
    for (int q = 1; q < R; ++q) {
      pq += p;
      if (pq >= R) pq -= R;
      float2 w = twd<INV>(tw, pq * nr);
      acc.x += v[q].x * w.x - v[q].y * w.y;
      acc.y += v[q].x * w.y + v[q].y * w.x;
    }
    out[j0 + p * Ns] = acc;
  }
}

// Transforms `nframes` frames held contiguously (N float2 each) in buf0 using buf1 as the ping-pong partner.
// Returns the buffer holding the result.  All threads of the CTA must call it.
template <bool INV>
__device__ float2* fft_frames(float2* buf0, float2* buf1, const float2* tw, const FftPlan& plan, int nframes) {
  const int N = plan.n;
  int Ns = 1;
  float2* src = buf0;
  float2* dst = buf1;
  for (int s = 0; s < plan.nstages; ++s) {
    const int R = plan.radix[s];
    const int per = N / R;
    const int total = nframes * per;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int f = idx / per;
      const int j = idx - f * per;
      const float2* in = src + (size_t)f * N;
      float2* out = dst + (size_t)f * N;
      switch (R) {
        case 2: butterfly<2, INV>(in, out, tw, N, Ns, j); break;
        case 3: butterfly<3, INV>(in, out, tw, N, Ns, j); break;
        case 4: butterfly<4, INV>(in, out, tw, N, Ns, j); break;
        case 5: butterfly<5, INV>(in, out, tw, N, Ns, j); break;
        case 7: butterfly<7, INV>(in, out, tw, N, Ns, j); break;
        default: butterfly_generic<INV>(in, out, tw, N, Ns, R, j); break;
      }
    }
    __syncthreads();
    Ns *= R;
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

}  // namespace bsrnn
