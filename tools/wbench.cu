// wbench.cu — HBM write-bandwidth micro-benchmark behind the input-projection GEMM's "write roofline" (DESIGN.md §4).
// Is the ~3.9 TB/s seen with torch fill_ a property of HBM writes on B200 or of plain st.global?  Compares, on one
// 8 GiB buffer: (1) st.global.v4 grid-stride, (2) st.global.cs.v4 (streaming / evict-first), (3) cp.async.bulk
// shared->global (TMA store) of 16 KB chunks, (4) cudaMemsetAsync, (5) read-only v4 loads, (6) copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/wbench tools/wbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void st_plain(uint4* p, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void st_cs(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
// each CTA owns contiguous 16 KB chunks: chunk c = blockIdx + k*grid
__global__ void st_bulk(uint8_t* p, size_t nchunks, int chunk_bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < chunk_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    int inflight = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + c * (size_t)chunk_bytes), "r"(s), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
__global__ void ld_only(const uint4* p, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(p + i);
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) *sink = acc;
}
__global__ void cp_kernel(const uint4* a, uint4* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = __ldg(a + i);
}

template <class F> static double time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t bytes = (size_t)8 << 30;
  uint8_t *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  const size_t n16 = bytes / 16;
  for (int ctas_per_sm : {2, 4, 8}) {
    const int grid = 148 * ctas_per_sm;
    printf("grid %d x 512:\n", grid);
    printf("  st.global.v4      %.0f GB/s\n", bytes / time_ms([&] { st_plain<<<grid, 512>>>((uint4*)a, n16); }) / 1e6);
    printf("  st.global.cs.v4   %.0f GB/s\n", bytes / time_ms([&] { st_cs<<<grid, 512>>>((uint4*)a, n16); }) / 1e6);
    printf("  ld.global.nc.v4   %.0f GB/s\n", bytes / time_ms([&] { ld_only<<<grid, 512>>>((const uint4*)a, n16, (uint4*)b); }) / 1e6);
    printf("  copy (r+w bytes)  %.0f GB/s\n", 2.0 * bytes / time_ms([&] { cp_kernel<<<grid, 512>>>((const uint4*)a, (uint4*)b, n16); }) / 1e6);
  }
  for (int chunk : {16384, 65536}) {
    cudaFuncSetAttribute(st_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
    for (int ctas_per_sm : {1, 2}) {
      if (chunk * ctas_per_sm > 200 * 1024) continue;
      const int grid = 148 * ctas_per_sm;
      printf("  bulk s2g %d B chunks, grid %d: %.0f GB/s\n", chunk, grid,
             bytes / time_ms([&] { st_bulk<<<grid, 128, chunk>>>(a, bytes / chunk, chunk); }) / 1e6);
    }
  }
  printf("  cudaMemsetAsync   %.0f GB/s\n", bytes / time_ms([&] { cudaMemsetAsync(a, 1, bytes); }) / 1e6);
  printf("  cudaMemcpyAsync d2d (r+w) %.0f GB/s\n", 2.0 * bytes / time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }) / 1e6);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
