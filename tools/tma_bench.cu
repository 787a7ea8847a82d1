// tma_bench.cu — cp.async.bulk (global/L2 -> shared) latency and per-SM throughput on sm_100a as a function of copy size
// and copies in flight; the source is an L2-resident region (private per CTA, or one region shared by all CTAs).
// Behind the "h ring" budget of the BLSTM recurrence kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma_bench tools/tma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

__global__ void k(const uint8_t* src, size_t region, int shared_src, uint32_t chunk, int stages, int iters, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[16];
  const uint8_t* base = src + (shared_src ? 0 : (size_t)blockIdx.x * region);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    size_t off = 0;
    const long long t0 = clock64();
    for (int i = 0; i < stages; ++i) {                       // fill the pipe
      expect_tx(full + i, chunk);
      bulk(smem + (size_t)i * chunk, base + off, chunk, full + i);
      off += chunk; if (off + chunk > region) off = 0;
    }
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      wait(full + s, (it / stages) & 1);
      expect_tx(full + s, chunk);
      bulk(smem + (size_t)s * chunk, base + off, chunk, full + s);
      off += chunk; if (off + chunk > region) off = 0;
    }
    for (int i = 0; i < stages; ++i) { const int it = iters + i; wait(full + it % stages, (it / stages) & 1); }
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
}

int main() {
  const size_t region = 128 << 10;
  uint8_t* src; long long* cyc;
  cudaMalloc(&src, 148 * region); cudaMemset(src, 1, 148 * region); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  const int iters = 4000;
  for (int shared_src = 0; shared_src < 2; ++shared_src)
    for (uint32_t chunk : {4096u, 8192u, 20480u, 32768u})
      for (int stages : {1, 2, 3, 6}) {
        if ((size_t)chunk * stages > (192u << 10)) continue;
        k<<<148, 32, (size_t)chunk * stages>>>(src, region, shared_src, chunk, stages, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)chunk * (iters + stages);
        printf("%s src, %5u B x %d in flight: %6.1f B/clk/SM, %6.0f cycles per copy round trip  [%s]\n",
               shared_src ? "shared " : "private", chunk, stages, bytes / c, (double)c * stages / (iters + stages), cudaGetErrorString(e));
      }
  return 0;
}
