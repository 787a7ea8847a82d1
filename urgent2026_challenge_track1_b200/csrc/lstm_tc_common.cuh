// lstm_tc_common.cuh — pieces shared by the BLSTM recurrence kernels (lstm_tc.cu: gates_x read from HBM;
// lstm_fused.cu: input projection fused into the recurrent MMA): tile geometry of the H = 392 decomposition, the
// global-memory flag primitives of the flag-group schedule, the LSTM cell update and the h-core store patterns.
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace bsrnn {
using namespace umma;

constexpr int LCL = 8;             // cluster size
constexpr int LUN = 49;            // hidden units per CTA
constexpr int LBN = 208;           // gate columns per CTA (4*49 = 196, padded to a multiple of 16)
constexpr int LGC = LBN / 8;       // 26 gates_x cores per CTA
constexpr int LKC = 50;            // k-cores of the recurrent operand (K = 400)
constexpr int LKS = 10;            // k-cores per A stage
constexpr int LNST = LKC / LKS;    // 5 stages per (step, slot)
constexpr int LSTAGES = 3;
constexpr int LNS = 3;             // slots (interleaved units per cluster)
constexpr int LACC = 256;          // TMEM columns per accumulator buffer
constexpr int LTHREADS = 512;      // 4 control warps + 3 slots x 4 epilogue warps

__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// pi, pf, po arrive pre-halved (see header): sigmoid = 0.5*tanh(.)+0.5
__device__ __forceinline__ void gate_update(float pi, float pf, float pg, float po, float& c, float& h) {
  const float ig = fmaf(tanh_fast(pi), 0.5f, 0.5f), fg = fmaf(tanh_fast(pf), 0.5f, 0.5f);
  const float gg = tanh_fast(pg), og = fmaf(tanh_fast(po), 0.5f, 0.5f);
  c = fmaf(fg, c, ig * gg);
  h = og * tanh_fast(c);
}

// training forward: the same update, and the ACTIVATED gates (i, f, g, o as 4 halves) + c_t written out for BPTT
// [what bsrnn_blstm_step_train_tc saves: gemm_tc.cu EPI_LSTM_STEP with save_gates]
// (the fused kernel stores them UNIT-major -- [unit][128 rows]: a warp's 32 rows are contiguous -- and a small transpose kernel
// turns them into the row-major buffers BPTT reads: per-row stores from the epilogue cost 32 sectors per instruction and
// tripled the step time, profiles/r02 call59)
__device__ __forceinline__ void gate_update_save(float pi, float pf, float pg, float po, float& c, float& h, __half* g4, float* cs) {
  const float ig = fmaf(tanh_fast(pi), 0.5f, 0.5f), fg = fmaf(tanh_fast(pf), 0.5f, 0.5f);
  const float gg = tanh_fast(pg), og = fmaf(tanh_fast(po), 0.5f, 0.5f);
  c = fmaf(fg, c, ig * gg);
  h = og * tanh_fast(c);
  const __half2 a = __floats2half2_rn(ig, fg), b = __floats2half2_rn(gg, og);
  *reinterpret_cast<uint2*>(g4) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  *cs = c;
}

// (f32 accumulator pair) + (fp16 pair packed in one register): sm_100a mixed-precision FHADD takes the fp16 operand
// (either half of the register) directly -- one instruction per gate instead of a convert and an add.
__device__ __forceinline__ float2 add_h2_f32(uint32_t g, uint32_t a0, uint32_t a1) {
  float2 d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %3;\n\tadd.rn.f32.f16 %1, hi, %4;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "r"(g), "f"(__uint_as_float(a0)), "f"(__uint_as_float(a1)));
  return d;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 p = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&p);
}
// slots [S0, S1) of one 16-byte core <- h[J0 ...]; 4-byte stores for aligned pairs, 2-byte for the edges
template <int S0, int S1, int J0>
__device__ __forceinline__ void store_partial(__half* core, const float (&h)[16]) {
#pragma unroll
  for (int sl = S0; sl < S1; ++sl) {
    const int j = J0 + (sl - S0);
    if ((sl & 1) == 0 && sl + 1 < S1) {
      *reinterpret_cast<uint32_t*>(core + sl) = pack_h2(h[j], h[j + 1]);
    } else if ((sl & 1) == 1 && sl > S0) {
      // upper half of a pair already written
    } else {
      core[sl] = __float2half_rn(h[j]);
    }
  }
}
template <int J0>
__device__ __forceinline__ void store_full(__half* core, const float (&h)[16]) {
  *reinterpret_cast<uint4*>(core) = make_uint4(pack_h2(h[J0], h[J0 + 1]), pack_h2(h[J0 + 2], h[J0 + 3]),
                                               pack_h2(h[J0 + 4], h[J0 + 5]), pack_h2(h[J0 + 6], h[J0 + 7]));
}

__device__ __forceinline__ void tmem_ld_pin16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

}  // namespace bsrnn
