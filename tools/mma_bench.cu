// mma_bench.cu — tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16, SS) issue-to-retire throughput on sm_100a for
// the no-swizzle K-major "KB8" operand layout used by gemm_tc.cu / lstm_tc.cu, as a function of N.  Operands sit in
// shared memory (garbage values), 1 CTA per SM, one thread issues REP batches of `nmma` MMAs + one commit and waits.
// Answers: is an M128 x N208 x K16 MMA really 104 cycles (the 128*N/256 floor) with this layout, or shared-memory bound?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I urgent2026_challenge_track1_b200/csrc -o tools/_bin/mma_bench tools/mma_bench.cu
#include "umma.cuh"
#include <stdio.h>
using namespace umma;

__global__ void k(int N, int nmma, int reps, int same_a, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < (200 << 10) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    const uint64_t da0 = smem_desc_kb8(smem_u32(smem), 2048, 128);                 // A: [kcore][128][8], 50 k-cores = 100 KB
    const uint64_t db0 = smem_desc_kb8(smem_u32(smem) + 100 * 1024, N * 16, 128);  // B: [kcore][N][8]
    const uint32_t b_step = (2 * N * 16) >> 4;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int j = 0; j < nmma; ++j) {
        const int ja = same_a ? 0 : (j % 25), jb = j % 3;                          // B region: 3 K-steps (stays under 100 KB for N = 256)
        mma_f16_ss(slot + (r & 1) * 256, da0 + (uint64_t)(ja * 256), db0 + (uint64_t)(jb * b_step), idesc, j != 0);
      }
      mma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  for (int same_a = 0; same_a < 2; ++same_a)
    for (int N : {64, 104, 128, 208, 256})
      for (int nmma : {13, 25, 100}) {
        const int reps = 200;
        k<<<148, 128, 200 << 10>>>(N, nmma, reps, same_a, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d, %3d MMAs per commit%s: %.1f cycles per MMA (floor 128*N/256 = %.0f)  [%s]\n", N, nmma,
               same_a ? " (same A tile)" : "", (double)c / ((double)reps * nmma), 128.0 * N / 256.0, cudaGetErrorString(e));
      }
  return 0;
}
