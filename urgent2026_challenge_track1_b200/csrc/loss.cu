// loss.cu — MultiResL1SpecLoss core, value AND gradient in one pass (training only).
//
// Replaces the body of espnet2 MultiResL1SpecLoss.forward as the reference configures it at d_model.py:24
// (window_sz = [256, 512, 768, 1024], hop = w/2, rectangular window, center=True reflect padding, onesided,
// reduction "sum", time_domain_weight 0.5) and its autograd backward (d_model.py:74 -> loss.backward()):
//
//   S[b] = w_td * sum_n |e[b,n] - t[b,n]|  +  sum_w w_sp * sum_{frames, bins} | |STFT_w(e)| - |STFT_w(t)| |
//
// for the already scaled / variance-normalised estimate e and target t (the scale alpha and the std normalisation
// are a handful of (B,)-sized reductions kept in the host-side autograd graph, losses.py).  d S / d e is produced in
// the same pass, which is all the backward needs (the loss is the root of the graph): grad[b,n] accumulates
//   w_td * sgn(e - t)  +  sum_w w_sp * Re sum_k sgn(|E_k| - |T_k|) * E_k/|E_k| * exp(+j 2 pi k n / w)   (reflect-adjoint).
//
// One CTA = a run of frames of one sample.  The two real signals of a frame are packed into ONE complex FFT
// (z = e + j t;  E_k = (Z_k + conj Z_{w-k})/2,  T_k = (Z_k - conj Z_{w-k})/(2j)), transformed in shared memory by the
// Stockham kernels of fft_core.cuh; the one-sided gradient spectrum goes back through the inverse transform of the same
// buffers and is scatter-added with the adjoint of the reflect padding.
#include "fft_core.cuh"

namespace bsrnn {

constexpr int kLossThreads = 256;

__device__ __forceinline__ int reflect_index(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// grid (ceil(T/fpb), B)
__global__ void __launch_bounds__(kLossThreads)
mrl1_spec_kernel(const float* __restrict__ est, const float* __restrict__ tgt, double* __restrict__ loss,
                 float* __restrict__ grad, const float2* __restrict__ twiddle, FftPlan plan, int L, int T, int hop, int fpb,
                 float weight) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = plan.n;
  const int F = N / 2 + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* buf0 = tw + N;
  float2* buf1 = buf0 + (size_t)fpb * N;
  __shared__ double s_part[kLossThreads / 32];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * fpb;
  const int nfr = min(fpb, T - t0);
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = twiddle[i];
  const float* er = est + (size_t)b * L;
  const float* tr = tgt + (size_t)b * L;
  for (int idx = threadIdx.x; idx < nfr * N; idx += blockDim.x) {
    const int f = idx / N;
    const int n = idx - f * N;
    const int i = reflect_index((t0 + f) * hop - N / 2 + n, L);
    buf0[idx] = make_float2(er[i], tr[i]);                    // rectangular window
  }
  __syncthreads();
  float2* res = fft_frames<false>(buf0, buf1, tw, plan, nfr);
  float2* other = (res == buf0) ? buf1 : buf0;
  // one-sided magnitudes, loss, gradient spectrum G_k (k <= N/2; the other bins stay zero)
  double part = 0.0;
  for (int idx = threadIdx.x; idx < nfr * N; idx += blockDim.x) {
    const int f = idx / N;
    const int k = idx - f * N;
    float2 g = make_float2(0.f, 0.f);
    if (k < F) {
      const float2 zk = res[(size_t)f * N + k];
      const float2 zc = res[(size_t)f * N + ((N - k) % N)];
      const float2 E = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 Tt = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float me = sqrtf(E.x * E.x + E.y * E.y), mt = sqrtf(Tt.x * Tt.x + Tt.y * Tt.y);
      const float d = me - mt;
      part += (double)fabsf(d);
      const float s = d > 0.f ? weight : (d < 0.f ? -weight : 0.f);
      if (me > 0.f) g = make_float2(s * E.x / me, s * E.y / me);
    }
    other[idx] = g;
  }
  __syncthreads();
  float2* gres = fft_frames<true>(other, res, tw, plan, nfr);      // unnormalised inverse: sum_k G_k e^{+j 2 pi k n / N}
  float* gr = grad + (size_t)b * L;
  for (int idx = threadIdx.x; idx < nfr * N; idx += blockDim.x) {
    const int f = idx / N;
    const int n = idx - f * N;
    const int i = reflect_index((t0 + f) * hop - N / 2 + n, L);
    atomicAdd(gr + i, gres[idx].x);
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kLossThreads / 32; ++i) t += s_part[i];
    atomicAdd(loss + b, t * (double)weight);
  }
}

// time-domain term: loss[b] += weight * sum |e - t| ; grad += weight * sgn(e - t).   grid (blocks, B)
__global__ void __launch_bounds__(256) l1_time_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                                                      double* __restrict__ loss, float* __restrict__ grad, int L,
                                                      float weight) {
  const int b = blockIdx.y;
  const float* er = est + (size_t)b * L;
  const float* tr = tgt + (size_t)b * L;
  float* gr = grad + (size_t)b * L;
  double part = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    const float d = er[i] - tr[i];
    part += (double)fabsf(d);
    gr[i] += d > 0.f ? weight : (d < 0.f ? -weight : 0.f);     // this kernel runs before the spectral ones: no race
  }
  part = warp_sum(part);
  __shared__ double s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_part[i];
    atomicAdd(loss + b, t * (double)weight);
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_l1_time_fwd_bwd(const float* est, const float* tgt, double* loss, float* grad, int B, int L,
                                     float weight, void* stream) {
  BSRNN_CHECK_ARG(est && tgt && loss && grad && B > 0 && L > 0, "l1_time_fwd_bwd: bad arguments");
  dim3 grid(cdiv(L, 256 * 4) < 64 ? cdiv(L, 256 * 4) : 64, B);
  l1_time_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(est, tgt, loss, grad, L, weight);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_mrl1_spec_fwd_bwd(const float* est, const float* tgt, double* loss, float* grad, const float* twiddle,
                                       int B, int L, int window, float weight, void* stream) {
  BSRNN_CHECK_ARG(est && tgt && loss && grad && twiddle, "mrl1_spec_fwd_bwd: null pointer");
  BSRNN_CHECK_ARG(B > 0 && window >= 4 && window % 2 == 0 && L > window / 2, "mrl1_spec_fwd_bwd: bad dims B=%d L=%d window=%d",
                  B, L, window);
  FftPlan plan;
  BSRNN_CHECK_ARG(make_plan(window, &plan), "mrl1_spec_fwd_bwd: cannot factorise window=%d", window);
  const int hop = window / 2;
  const int T = 1 + L / hop;
  int fpb = 8;
  while (fpb > 1 && (size_t)window * 8 * (1 + 2 * fpb) > 100 * 1024) fpb >>= 1;
  const size_t smem = (size_t)window * 8 * (1 + 2 * fpb);
  BSRNN_CUDA_OK(cudaFuncSetAttribute(mrl1_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, fpb), B);
  mrl1_spec_kernel<<<grid, kLossThreads, smem, (cudaStream_t)stream>>>(est, tgt, loss, grad,
                                                                       reinterpret_cast<const float2*>(twiddle), plan, L, T,
                                                                       hop, fpb, weight);
  BSRNN_LAUNCH_OK();
  return 0;
}
