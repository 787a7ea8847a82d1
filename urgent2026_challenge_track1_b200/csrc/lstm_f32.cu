// lstm_f32.cu — f32 BLSTM recurrence (CUDA cores), one launch per step covering both directions.
// Replaces nn.LSTM(N, 2N, batch_first, bidirectional) forward with zero initial state
// [bsrnn_flowse.py:226-238 construct; :296-297 time axis, :303-304 band axis]; gate order i,f,g,o;
// c' = sig(f)*c + sig(i)*tanh(g), h = sig(o)*tanh(c').  The bf16 mode runs the persistent tcgen05 kernel in
// lstm_tc.cu instead.  y doubles as the carrier of h between steps, so no separate hidden-state buffer exists.
#include "common.cuh"

namespace bsrnn {

constexpr int LM = 64, LU = 16, LK = 16, LT = 256;   // 64 sequences x 16 hidden units (x4 gates) per CTA

struct SeqAddr {
  long seq_inner, seq_outer, seq_inner_stride, step_stride;
  __device__ __forceinline__ long token(long r, long s) const {
    return (r / seq_inner) * seq_outer + (r % seq_inner) * seq_inner_stride + s * step_stride;
  }
};

// grid (ceil(H/16), ceil(R/64), 2)
__global__ void __launch_bounds__(LT)
lstm_step_f32_kernel(const float* __restrict__ gates_x, const float* __restrict__ w_hh, float* __restrict__ y,
                     float* __restrict__ c_state, int R, int H, int step, int steps, SeqAddr addr) {
  const int d = blockIdx.z;
  const int u0 = blockIdx.x * LU;
  const int r0 = blockIdx.y * LM;
  const int s_cur = d == 0 ? step : steps - 1 - step;
  const int s_prev = d == 0 ? s_cur - 1 : s_cur + 1;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;

  __shared__ float As[LK][LM + 4];
  __shared__ float Bs[LK][LM + 4];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (step > 0) {
    const int arow = r0 + lr;
    const float* hp = arow < R ? y + addr.token(arow, s_prev) * (2L * H) + (long)d * H : nullptr;
    const int gate = lr >> 4, ul = lr & 15;
    const bool w_ok = (u0 + ul) < H;
    const float* wp = w_hh + ((long)d * 4 * H + (long)gate * H + u0 + ul) * H;
    for (int k0 = 0; k0 < H; k0 += LK) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + lk + u;
        As[lk + u][lr] = (hp && k < H) ? hp[k] : 0.f;
        Bs[lk + u][lr] = (w_ok && k < H) ? wp[k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < LK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const int u = u0 + tx;
  if (u >= H) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= R) continue;
    const long tok = addr.token(r, s_cur);
    const float* gx = gates_x + tok * (8L * H) + (long)d * 4 * H;
    const float pi = acc[i][0] + gx[u];
    const float pf = acc[i][1] + gx[H + u];
    const float pg = acc[i][2] + gx[2 * H + u];
    const float po = acc[i][3] + gx[3 * H + u];
    const float ig = 1.f / (1.f + expf(-pi));
    const float fg = 1.f / (1.f + expf(-pf));
    const float gg = tanhf(pg);
    const float og = 1.f / (1.f + expf(-po));
    float* cp = c_state + ((long)d * R + r) * H + u;
    const float c_prev = step > 0 ? *cp : 0.f;
    const float c_new = fg * c_prev + ig * gg;
    *cp = c_new;
    y[tok * (2L * H) + (long)d * H + u] = og * tanhf(c_new);
  }
}

}  // namespace bsrnn
using namespace bsrnn;

extern "C" int bsrnn_blstm_recurrence_f32(const float* gates_x, const float* w_hh, float* y, float* c_state, int R,
                                          int steps, int H, long seq_inner, long seq_outer, long seq_inner_stride,
                                          long step_stride, void* stream) {
  BSRNN_CHECK_ARG(gates_x && w_hh && y && c_state, "blstm_recurrence_f32: null pointer");
  BSRNN_CHECK_ARG(R > 0 && steps > 0 && H > 0 && seq_inner > 0, "blstm_recurrence_f32: bad dims");
  SeqAddr addr{seq_inner, seq_outer, seq_inner_stride, step_stride};
  dim3 grid(cdiv(H, LU), cdiv(R, LM), 2);
  for (int s = 0; s < steps; ++s) {
    lstm_step_f32_kernel<<<grid, LT, 0, (cudaStream_t)stream>>>(gates_x, w_hh, y, c_state, R, H, s, steps, addr);
  }
  BSRNN_LAUNCH_OK();
  count_launches(steps - 1);
  return 0;
}
