// flow.cu — FlowSE-only pieces: Gaussian-Fourier time embedding, GradDecoder 5x5 conv + GLU, prior sampling,
// fused Euler update.  References: baseline_code/models/bsrnn_flowse.py:90-99,114-117,163-167,
// baseline_code/models/odes.py:84-91, baseline_code/sampling/odesolvers.py:76-81.
#include "common.cuh"

namespace bsrnn {

__global__ void time_embed_kernel(const float* __restrict__ t, const float* __restrict__ W, float* __restrict__ out,
                                  int B, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, e = i % E;
  // same association as `x[:, None] * W[None, :] * 2 * pi` evaluated left to right in f32
  const float p = ((t[b] * W[e]) * 2.0f) * 3.14159265358979323846f;
  out[(size_t)b * 2 * E + e] = sinf(p);
  out[(size_t)b * 2 * E + E + e] = cosf(p);
}

constexpr int CT = 8, CF = 32, CC = 16;   // output tile: 8 frames x 32 bins, 16 input channels

// g (B,T,Fp,16) -> out (B,T,F,2);  grid (ceil(Fp/32), ceil(T/8), B), block 256 = 8 x 32
__global__ void __launch_bounds__(256)
conv5x5_glu_kernel(const float* __restrict__ g, const float* __restrict__ weight, const float* __restrict__ bias,
                   float* __restrict__ out, int T, int Fp, int F) {
  __shared__ float tile[CT + 4][CF + 4][CC + 1];
  __shared__ float w[4][25][CC];                      // [co][df*5+dt][ci]
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * CT, f0 = blockIdx.x * CF;
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * CC * 25; i += 256) {
    const int co = i / (CC * 25), rem = i % (CC * 25), ci = rem / 25, kk = rem % 25;   // weight (4,16,5,5)
    w[co][kk][ci] = weight[i];
  }
  for (int i = tid; i < (CT + 4) * (CF + 4) * CC; i += 256) {
    const int ci = i % CC, ff = (i / CC) % (CF + 4), tt = i / (CC * (CF + 4));
    const int t = t0 + tt - 2, f = f0 + ff - 2;
    float v = 0.f;
    if (t >= 0 && t < T && f >= 0 && f < Fp) v = g[(((size_t)b * T + t) * Fp + f) * CC + ci];
    tile[tt][ff][ci] = v;
  }
  __syncthreads();
  const int lf = tid & 31, lt = tid >> 5;
  const int t = t0 + lt, f = f0 + lf;
  if (t >= T || f >= F) return;
  float2 o = make_float2(0.f, 0.f);
  if (f < Fp) {
    float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
    for (int df = 0; df < 5; ++df)
#pragma unroll
      for (int dt = 0; dt < 5; ++dt) {
        const float* in = tile[lt + dt][lf + df];
        const int kk = df * 5 + dt;                    // kernel dims are (freq, time): image is (F', T)
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
          const float v = in[ci];
          acc[0] = fmaf(v, w[0][kk][ci], acc[0]);
          acc[1] = fmaf(v, w[1][kk][ci], acc[1]);
          acc[2] = fmaf(v, w[2][kk][ci], acc[2]);
          acc[3] = fmaf(v, w[3][kk][ci], acc[3]);
        }
      }
    o.x = acc[0] / (1.f + expf(-acc[2]));              // GLU(dim=1): first half * sigmoid(second half)
    o.y = acc[1] / (1.f + expf(-acc[3]));
  }
  reinterpret_cast<float2*>(out)[((size_t)b * T + t) * F + f] = o;
}

__global__ void euler_step_kernel(float2* __restrict__ x, const float2* __restrict__ m, const float2* __restrict__ r,
                                  float step, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float2 xv = x[i], mv = m[i], rv = r[i];
    const float gx = mv.x * xv.x - mv.y * xv.y + rv.x;
    const float gy = mv.x * xv.y + mv.y * xv.x + rv.y;
    x[i] = make_float2(xv.x + step * gx, xv.y + step * gy);
  }
}

__global__ void complex_mask_kernel(float2* __restrict__ out, const float2* __restrict__ x, const float2* __restrict__ m,
                                    const float2* __restrict__ r, float sign, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float2 xv = x[i], mv = m[i], rv = r[i];
    out[i] = make_float2(sign * (mv.x * xv.x - mv.y * xv.y + rv.x), sign * (mv.x * xv.y + mv.y * xv.x + rv.y));
  }
}

__global__ void axpy_complex_kernel(float2* __restrict__ out, const float2* __restrict__ y, const float2* __restrict__ z,
                                    float sigma, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float2 a = y[i], b = z[i];
    out[i] = make_float2(a.x + sigma * b.x, a.y + sigma * b.y);
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_time_embed(const float* t, const float* W, float* out, int B, int E, void* stream) {
  BSRNN_CHECK_ARG(t && W && out && B > 0 && E > 0, "time_embed: bad arguments");
  time_embed_kernel<<<cdiv((long)B * E, 128), 128, 0, (cudaStream_t)stream>>>(t, W, out, B, E);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_conv5x5_glu(const float* g, const float* weight, const float* bias, float* out, int B, int T,
                                 int Fp, int F, void* stream) {
  BSRNN_CHECK_ARG(g && weight && bias && out && B > 0 && T > 0 && Fp > 0 && F > 0, "conv5x5_glu: bad arguments");
  dim3 grid(cdiv(F > Fp ? F : Fp, CF), cdiv(T, CT), B);
  conv5x5_glu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, weight, bias, out, T, Fp, F);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_euler_step(float* x, const float* mask, const float* resid, float step, long n_complex,
                                void* stream) {
  BSRNN_CHECK_ARG(x && mask && resid && n_complex > 0, "euler_step: bad arguments");
  euler_step_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float2*>(x), reinterpret_cast<const float2*>(mask), reinterpret_cast<const float2*>(resid),
      step, n_complex);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_axpy_complex(float* out, const float* y, const float* z, float sigma, long n_complex,
                                  void* stream) {
  BSRNN_CHECK_ARG(out && y && z && n_complex > 0, "axpy_complex: bad arguments");
  axpy_complex_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float2*>(out), reinterpret_cast<const float2*>(y), reinterpret_cast<const float2*>(z), sigma,
      n_complex);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_complex_mask(float* out, const float* x, const float* mask, const float* resid, float sign,
                                  long n_complex, void* stream) {
  BSRNN_CHECK_ARG(out && x && mask && resid && n_complex > 0, "complex_mask: bad arguments");
  complex_mask_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float2*>(out), reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(mask),
      reinterpret_cast<const float2*>(resid), sign, n_complex);
  BSRNN_LAUNCH_OK();
  return 0;
}
