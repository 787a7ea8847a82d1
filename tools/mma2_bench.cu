// mma2_bench.cu — tcgen05.mma.cta_group::2 (kind::f16, M = 256, K = 16, SS) issue-to-retire throughput on sm_100a for the
// no-swizzle K-major "KB8" operand layout, as a function of N (each CTA holds N/2 rows of B), against cta_group::1 at
// M = 128 (tools/mma_bench.cu).  2-CTA clusters, 1 CTA per SM, the leader issues REP batches of `nmma` MMAs + a commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I urgent2026_challenge_track1_b200/csrc -o tools/_bin/mma2_bench tools/mma2_bench.cu
#include "umma.cuh"
#include <stdio.h>
using namespace umma;

__global__ void k2(int N, int nmma, int reps, int same_a, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc2(&slot, 512);
  for (int i = threadIdx.x; i < (200 << 10) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (rank == 0 && warp == 1 && elect_one()) {
    const uint32_t idesc = idesc_f16_f32(256, N);
    const uint32_t BH = N / 2;
    const uint64_t da0 = smem_desc_kb8(smem_u32(smem), 2048, 128);                  // A: [kcore][128][8], 50 k-cores = 100 KB
    const uint64_t db0 = smem_desc_kb8(smem_u32(smem) + 100 * 1024, BH * 16, 128);  // B half: [kcore][N/2][8]
    const uint32_t b_step = (2 * BH * 16) >> 4;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int j = 0; j < nmma; ++j) {
        const int ja = same_a ? 0 : (j % 25), jb = j % 6;
        mma_f16_ss_2cta(slot + (r & 1) * 256, da0 + (uint64_t)(ja * 256), db0 + (uint64_t)(jb * b_step), idesc, j != 0);
      }
      mma_commit2_multicast(&bar, (uint16_t)1);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(slot, 512); }
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  for (int same_a = 0; same_a < 2; ++same_a)
    for (int N : {64, 128, 192, 208, 224, 256})
      for (int nmma : {13, 38, 100}) {
        const int reps = 200;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(144); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 << 10;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e0 = cudaLaunchKernelEx(&cfg, k2, N, nmma, reps, same_a, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("2CTA M=256 N=%3d, %3d MMAs per commit%s: %.1f cycles per MMA (per-SM floor 128*N/256 = %.0f)  [%s / %s]\n", N, nmma,
               same_a ? " (same A tile)" : "", (double)c / ((double)reps * nmma), 128.0 * N / 256.0, cudaGetErrorString(e0),
               cudaGetErrorString(e));
      }
  return 0;
}
