// optim.cu — the post-allreduce tail of a training step on ONE flat f32 parameter / gradient buffer:
// global gradient norm (+ non-finite detection), clip-by-norm, AdamW, optional EMA — two launches, no host sync.
// Replaces Lightning's clip_grad_norm_(0.5) [reference train_se.py:78], the NaN guard of SEModel.optimizer_step
// [d_model.py:36-59], torch.optim.AdamW [d_model.py:102-109, flow_model.py:238-245] and torch_ema's update
// [flow_model.py:84], which the reference runs as hundreds of small kernels with a host sync per parameter.
#include "common.cuh"
#include <math.h>

namespace bsrnn {

// stats[0] += sum g^2 (double) ; stats[1] = 1 if any non-finite gradient was seen
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ stats) {
  double acc = 0.0;
  bool bad = false;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (long j = i; j < n; ++j) {
        bad |= !isfinite(g[j]);
        acc += (double)g[j] * g[j];
      }
    }
  }
  acc = warp_sum(acc);
  __shared__ double part[8];
  __shared__ int sbad;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) sbad = 1;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(stats, t);
    if (sbad) stats[1] = 1.0;
  }
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* ema;
  long n;
  const double* stats;      // {sum g^2, non-finite flag} from grad_sumsq_kernel (after the allreduce)
  float grad_scale;         // gradients are multiplied by this first (1/world for a summed allreduce)
  float max_norm;           // <= 0: no clipping
  float lr, beta1, beta2, eps, weight_decay;
  float bias1, bias2;       // 1 - beta1^t, 1 - beta2^t
  float ema_decay;          // effective decay for this update; ema == nullptr: no EMA
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a) {
  if (a.stats[1] != 0.0) return;                               // non-finite gradient: skip the whole update
  const double norm = sqrt(a.stats[0]) * (double)a.grad_scale;
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) {
    const double c = (double)a.max_norm / (norm + 1e-6);        // torch.nn.utils.clip_grad_norm_
    if (c < 1.0) coef *= (float)c;
  }
  const float step_size = a.lr / a.bias1;
  const float inv_sqrt_bias2 = rsqrtf(a.bias2);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const float g = a.g[i] * coef;
    float p = a.p[i] * (1.f - a.lr * a.weight_decay);          // decoupled weight decay
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.m[i] = m;
    a.v[i] = v;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bias2 + a.eps);
    a.p[i] = p;
    if (a.ema) a.ema[i] -= (1.f - a.ema_decay) * (a.ema[i] - p);
  }
}

// ---- v2 tail: step counters on the device, parameters that received no gradient are left alone ------------------
// state (double[8]): [0] sum g^2, [1] non-finite flag (both from grad_sumsq_kernel), [2] optimizer steps taken,
// [3] EMA updates taken, [4] 1-beta1^t, [5] 1-beta2^t, [6] effective EMA decay of this update.
// A step whose gradients are non-finite takes no update AND does not advance the counters, which is what the
// reference's optimizer_step does (zero_grad() before AdamW.step(): every p.grad is None, so AdamW skips the
// parameter and its per-parameter step counter; d_model.py:36-59).  torch_ema.update still runs on such a step
// (flow_model.py:66-84: it follows optimizer_step unconditionally), so the EMA counter and average always advance.
__global__ void adam_tick_kernel(double* state, float beta1, float beta2, float ema_decay) {
  if (state[1] == 0.0) {
    const double t = state[2] + 1.0;
    state[2] = t;
    state[4] = 1.0 - pow((double)beta1, t);
    state[5] = 1.0 - pow((double)beta2, t);
  }
  if (ema_decay > 0.f) {                                     // torch_ema: d = min(decay, (1+n)/(10+n))
    const double n = state[3] + 1.0;
    state[3] = n;
    const double w = (1.0 + n) / (10.0 + n);
    state[6] = w < (double)ema_decay ? w : (double)ema_decay;
  }
}

constexpr int MAX_SKIP = 32;
struct Adam2Args {
  float* p; const float* g; float* m; float* v; float* ema;
  long n;
  const double* state;
  const long* skip;         // n_skip half-open element ranges [lo, hi) the update leaves untouched (grad = None)
  int n_skip;
  float grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay;
};

__global__ void __launch_bounds__(256) adamw2_kernel(const Adam2Args a) {
  const bool bad = a.state[1] != 0.0;
  if (bad && !a.ema) return;
  __shared__ long s_lo[MAX_SKIP], s_hi[MAX_SKIP];
  if ((int)threadIdx.x < a.n_skip) { s_lo[threadIdx.x] = a.skip[2 * threadIdx.x]; s_hi[threadIdx.x] = a.skip[2 * threadIdx.x + 1]; }
  __syncthreads();
  const double norm = sqrt(a.state[0]) * (double)a.grad_scale;
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) {
    const double c = (double)a.max_norm / (norm + 1e-6);
    if (c < 1.0) coef *= (float)c;
  }
  const float step_size = a.lr / (float)a.state[4];
  const float inv_sqrt_bias2 = rsqrtf((float)a.state[5]);
  const float ema_keep = (float)a.state[6];
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    bool skip = bad;
    for (int r = 0; r < a.n_skip; ++r) skip |= (i >= s_lo[r] && i < s_hi[r]);
    float p = a.p[i];
    if (!skip) {
      const float g = a.g[i] * coef;
      p *= (1.f - a.lr * a.weight_decay);
      const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
      const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
      a.m[i] = m;
      a.v[i] = v;
      p -= step_size * m / (sqrtf(v) * inv_sqrt_bias2 + a.eps);
      a.p[i] = p;
    }
    if (a.ema) a.ema[i] -= (1.f - ema_keep) * (a.ema[i] - p);   // torch_ema averages every parameter it tracks
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_adamw_step2(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, long n,
                                 double* state, const long* skip_ranges, int n_skip, float grad_scale, float max_norm,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                                 void* stream) {
  BSRNN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && state && n > 0, "adamw_step2: bad arguments");
  BSRNN_CHECK_ARG(n_skip >= 0 && n_skip <= MAX_SKIP && (n_skip == 0 || skip_ranges), "adamw_step2: at most 32 skip ranges");
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2, ema ? ema_decay : 0.f);
  BSRNN_LAUNCH_OK();
  Adam2Args a{param, grad, exp_avg, exp_avg_sq, ema, n, state, skip_ranges, n_skip, grad_scale, max_norm, lr, beta1, beta2,
              eps, weight_decay};
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  adamw2_kernel<<<blocks, 256, 0, st>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_grad_sumsq(const float* grad, long n, double* stats, void* stream) {
  BSRNN_CHECK_ARG(grad && stats && n > 0, "grad_sumsq: bad arguments");
  BSRNN_CHECK_ARG((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_sumsq: grad must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  BSRNN_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
  const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 + 1 : 148 * 8);
  grad_sumsq_kernel<<<blocks, 256, 0, st>>>(grad, n, stats);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, long n,
                                const double* stats, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int step, float ema_decay, void* stream) {
  BSRNN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && stats && n > 0 && step > 0, "adamw_step: bad arguments");
  AdamArgs a{param, grad, exp_avg, exp_avg_sq, ema, n, stats, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay,
             (float)(1.0 - pow((double)beta1, (double)step)), (float)(1.0 - pow((double)beta2, (double)step)), ema_decay};
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}
