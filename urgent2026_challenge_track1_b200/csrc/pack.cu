// pack.cu — GroupNorm-apply + cast + re-tile: f32 token-major activations -> fp16 UMMA operand tiles (KB8 layout).
// This is the "normalise" half of nn.GroupNorm(1,N) [reference bsrnn_flowse.py:291,302; decoder norms :146-152]
// fused with the layout change the tensor-core GEMMs need, so the normalised activation is written exactly once, in
// the form the next GEMM's bulk copies fetch.
#include "common.cuh"
#include <cuda_fp16.h>

namespace bsrnn {

struct PackArgs {
  const float* x;          // token-major rows of ldx floats
  const float* scale;      // (groups, C) or null
  const float* shift;
  __half* out;             // [m_tiles][kcores][128][8]
  long ldx;
  int col0, C, kcores;
  int tiles_per_step, R;
  long seq_inner, seq_outer, seq_inner_stride, step_stride;
  long tokens_per_sample;  // scale row = (token / tokens_per_sample) * g_inner + (g_inner > 1 ? token % g_inner : 0)
  int g_inner;
};

// grid (m_tiles), block 256, dynamic smem 128 * (kcores*8 + 8) halves
__global__ void __launch_bounds__(256) norm_cast_kb8_kernel(const PackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* tile = reinterpret_cast<__half*>(smem_raw);
  const int ld = a.kcores * 8 + 8;                 // +8 halves: rows land in different banks
  const int m = blockIdx.x;
  const int step = m / a.tiles_per_step, j = m - step * a.tiles_per_step;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kw = a.kcores * 8;
  for (int r = warp; r < 128; r += 8) {
    const long seq = (long)j * 128 + r;
    const bool ok = seq < a.R;
    long token = 0;
    const float *sc = nullptr, *sh = nullptr;
    if (ok) {
      token = (seq / a.seq_inner) * a.seq_outer + (seq % a.seq_inner) * a.seq_inner_stride + (long)step * a.step_stride;
      if (a.scale) {
        const long g = (token / a.tokens_per_sample) * a.g_inner + (a.g_inner > 1 ? token % a.g_inner : 0);
        sc = a.scale + g * a.C;
        sh = a.shift + g * a.C;
      }
    }
    const float* row = a.x + token * a.ldx + a.col0;
    for (int c = lane; c < kw; c += 32) {
      float v = 0.f;
      if (ok && c < a.C) {
        v = row[c];
        if (sc) v = fmaf(v, sc[c], sh[c]);
      }
      tile[r * ld + c] = __float2half_rn(v);
    }
  }
  __syncthreads();
  __half* dst = a.out + (size_t)m * a.kcores * 128 * 8;
  for (int i = threadIdx.x; i < a.kcores * 128; i += 256) {
    const int kc = i >> 7, r = i & 127;
    *reinterpret_cast<uint4*>(dst + (size_t)i * 8) = *reinterpret_cast<const uint4*>(tile + r * ld + kc * 8);
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_norm_cast_kb8(const float* x, const float* scale, const float* shift, void* out, long ldx,
                                   int col0, int C, int kcores, int m_tiles, int tiles_per_step, int R,
                                   long seq_inner, long seq_outer, long seq_inner_stride, long step_stride,
                                   long tokens_per_sample, int g_inner, void* stream) {
  BSRNN_CHECK_ARG(x && out && C > 0 && kcores * 8 >= C && m_tiles > 0 && tiles_per_step > 0 && seq_inner > 0 &&
                  tokens_per_sample > 0 && g_inner > 0, "norm_cast_kb8: bad arguments");
  BSRNN_CHECK_ARG((scale == nullptr) == (shift == nullptr), "norm_cast_kb8: scale and shift come together");
  PackArgs a{x, scale, shift, reinterpret_cast<__half*>(out), ldx, col0, C, kcores, tiles_per_step, R,
             seq_inner, seq_outer, seq_inner_stride, step_stride, tokens_per_sample, g_inner};
  const size_t smem = (size_t)128 * (kcores * 8 + 8) * 2;
  BSRNN_CUDA_OK(cudaFuncSetAttribute(norm_cast_kb8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  norm_cast_kb8_kernel<<<m_tiles, 256, smem, (cudaStream_t)stream>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}
