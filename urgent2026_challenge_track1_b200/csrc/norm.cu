// norm.cu — GroupNorm(1, C) statistics and the (scale, shift) tables consumers fold into their A-operand load.
// Replaces the reduction half of nn.GroupNorm at bsrnn_flowse.py:73 (BandSplit), :291/:302 (norm_time/norm_freq)
// and :146-152 / espnet2 MaskDecoder (per-band decoder norms).  Sums are accumulated in double so the
// E[x^2]-E[x]^2 form is safe; zeros of padded frames / bins are part of the population.
#include "common.cuh"

namespace bsrnn {

constexpr int kStatThreads = 256;

__device__ __forceinline__ void block_accumulate(double s, double q, double* out) {
  __shared__ double sh[2][kStatThreads / 32];
  s = warp_sum(s);
  q = warp_sum(q);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = q; }
  __syncthreads();
  if (w == 0) {
    s = l < kStatThreads / 32 ? sh[0][l] : 0.0;
    q = l < kStatThreads / 32 ? sh[1][l] : 0.0;
    s = warp_sum(s);
    q = warp_sum(q);
    if (l == 0) { atomicAdd(out, s); atomicAdd(out + 1, q); }
  }
}

// grid (chunks, B): contiguous-per-row population
__global__ void __launch_bounds__(kStatThreads)
gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, long rows, int C, long row_stride,
                long rows_per_block) {
  const int b = blockIdx.y;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  const float* base = x + (size_t)b * rows * row_stride;
  float s = 0.f, q = 0.f;      // per-thread partials stay short (<= rows_per_block*C/256 terms), promoted below
  double S = 0.0, Q = 0.0;
  if (row_stride == C) {
    const long e0 = r0 * C, e1 = r1 * C;
    int cnt = 0;
    for (long e = e0 + threadIdx.x; e < e1; e += kStatThreads) {
      const float v = base[e];
      s += v; q += v * v;
      if (++cnt == 64) { S += s; Q += q; s = q = 0.f; cnt = 0; }
    }
  } else {
    for (long r = r0; r < r1; ++r)
      for (int c = threadIdx.x; c < C; c += kStatThreads) {
        const float v = base[r * row_stride + c];
        s += v; q += v * v;
      }
  }
  S += s; Q += q;
  block_accumulate(S, Q, stats + 2 * b);
}

// grid (n_bands, B, tchunks): x rows of row_len floats, one row per (b,t); band k = floats [off_k, off_k+width_k)
__global__ void __launch_bounds__(kStatThreads)
band_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int T, long row_len,
                  const int32_t* __restrict__ off, const int32_t* __restrict__ width, int n_bands, int t_per_block) {
  const int k = blockIdx.x, b = blockIdx.y;
  const int t0 = blockIdx.z * t_per_block, t1 = min(T, t0 + t_per_block);
  const int o = off[k], w = width[k];
  double S = 0.0, Q = 0.0;
  float s = 0.f, q = 0.f;
  const int total = (t1 - t0) * w;
  int cnt = 0;
  for (int e = threadIdx.x; e < total; e += kStatThreads) {
    const int t = t0 + e / w, c = e % w;
    const float v = x[((size_t)b * T + t) * row_len + o + c];
    s += v; q += v * v;
    if (++cnt == 64) { S += s; Q += q; s = q = 0.f; cnt = 0; }
  }
  S += s; Q += q;
  block_accumulate(S, Q, stats + 2 * ((size_t)b * n_bands + k));
}

// grid (G), block 128
__global__ void gn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ extra,
                                   float* __restrict__ scale, float* __restrict__ shift, int C,
                                   const double* __restrict__ counts, float eps, int G_inner) {
  const int g = blockIdx.x;
  const int gi = g % G_inner;
  const double cnt = counts[gi];
  const double mean = stats[2 * g] / cnt;
  double var = stats[2 * g + 1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float mu = (float)mean;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float ga = gamma[(size_t)gi * C + c], be = beta[(size_t)gi * C + c];
    const float sc = rstd * ga;
    float sh = be - mu * sc;
    if (extra) sh += extra[(size_t)g * C + c];
    scale[(size_t)g * C + c] = sc;
    shift[(size_t)g * C + c] = sh;
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_gn_stats(const float* x, double* stats, int B, long rows_per_sample, int C, long row_stride,
                              void* stream) {
  BSRNN_CHECK_ARG(x && stats && B > 0 && rows_per_sample > 0 && C > 0, "gn_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BSRNN_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, st));
  long chunks = (148L * 8 + B - 1) / B;
  long rpb = (rows_per_sample + chunks - 1) / chunks;
  if (rpb < 1) rpb = 1;
  dim3 grid(cdiv(rows_per_sample, rpb), B);
  gn_stats_kernel<<<grid, kStatThreads, 0, st>>>(x, stats, rows_per_sample, C, row_stride, rpb);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_band_stats(const float* x, double* stats, int B, int T, long row_len, const int32_t* band_off,
                                const int32_t* band_width, int n_bands, void* stream) {
  BSRNN_CHECK_ARG(x && stats && band_off && band_width && B > 0 && T > 0 && n_bands > 0, "band_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BSRNN_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * n_bands, st));
  int tchunks = (148 * 8) / (B * n_bands) + 1;
  if (tchunks > T) tchunks = T;
  const int tpb = cdiv(T, tchunks);
  dim3 grid(n_bands, B, cdiv(T, tpb));
  band_stats_kernel<<<grid, kStatThreads, 0, st>>>(x, stats, T, row_len, band_off, band_width, n_bands, tpb);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_gn_finalize(const double* stats, const float* gamma, const float* beta, const float* extra,
                                 float* scale, float* shift, int G, int C, const double* counts, float eps,
                                 int G_inner, void* stream) {
  BSRNN_CHECK_ARG(stats && gamma && beta && scale && shift && counts && G > 0 && C > 0 && G_inner > 0,
                  "gn_finalize: bad arguments");
  gn_finalize_kernel<<<G, 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, extra, scale, shift, C, counts, eps,
                                                          G_inner);
  BSRNN_LAUNCH_OK();
  return 0;
}
