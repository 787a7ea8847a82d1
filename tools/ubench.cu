// ubench.cu — B200 micro-benchmarks behind the recurrence-epilogue design (standalone: nvcc -arch=sm_100a).
//   1. MUFU.TANH (tanh.approx.f32 / .f16x2) and ex2 issue rate per SM at 8 and 16 warps
//   2. the LSTM gate update (5 MUFU per unit) for 24 units per thread, registers only
//   3. tcgen05.ld 32x32b.x32 rate per SM with 4 / 8 warps
// Prints SM cycles per warp-instruction / per item.  Build+run: see tools/gpu_call9.sh
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float tanh_fast(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_h2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.01f * (threadIdx.x + i);
  uint32_t hv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hv[i] = 0x3c003800u + threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = tanh_fast(v[i]);
      if (MODE == 1) v[i] = ex2_fast(v[i]);
      if (MODE == 2) hv[i] = tanh_h2(hv[i]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(hv[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ float sigm_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

// 24 units per thread, c in registers, "accumulator" and "gx" synthesized from registers: pure ALU/MUFU cost of one item
__global__ void gates_kernel(float* out, long long* cyc, int iters) {
  float c[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) c[i] = 0.f;
  float a0 = 0.001f * threadIdx.x, hsum = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      const float pi = a0 + 0.1f * u, pf = a0 - 0.05f * u, pg = a0 * 0.5f + u, po = a0 + 0.2f;
      const float ig = sigm_fast(pi), fg = sigm_fast(pf), gg = tanh_fast(pg), og = sigm_fast(po);
      c[u] = fmaf(fg, c[u], ig * gg);
      hsum += og * tanh_fast(c[u]);
    }
    a0 += 1e-3f * hsum;
  }
  const long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = hsum;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ldtm_kernel(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float s = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(base + ch * 32) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]) * 1e-30f;
    }
  }
  const long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

static long long run_max(long long* d, int n) {
  long long h[256], m = 0;
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m;
  return m;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 256 * sizeof(long long));
  const int iters = 2000;
  const char* names[3] = {"tanh.approx.f32", "ex2.approx.f32", "tanh.approx.f16x2"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) mufu_kernel<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) mufu_kernel<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) mufu_kernel<2><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      const double c = (double)run_max(cyc, 148);
      printf("%-18s warps/SM=%2d: %.2f cycles per warp-instr per SMSP  (=> %.1f lanes/clk/SM)\n", names[mode], warps,
             c / (iters * 8.0 * warps / 4.0), 32.0 * iters * 8 * warps / c);
    }
  }
  for (int warps : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { gates_kernel<<<148, warps * 32>>>(out, cyc, 200); cudaDeviceSynchronize(); }
    const double c = (double)run_max(cyc, 148);
    printf("gate update 24 units/thread, warps/SM=%2d: %.0f cycles per item-equivalent (all warps, 24 units each)\n", warps, c / 200);
  }
  for (int warps : {4, 8}) {
    for (int rep = 0; rep < 2; ++rep) { ldtm_kernel<<<148, warps * 32>>>(out, cyc, 500); cudaDeviceSynchronize(); }
    const double c = (double)run_max(cyc, 148);
    printf("tcgen05.ld 32x32b.x32 (+wait), warps/SM=%d: %.1f cycles per ld per warp; %.1f B/clk/SM\n", warps, c / (500 * 6),
           4096.0 * 500 * 6 * warps / c);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
