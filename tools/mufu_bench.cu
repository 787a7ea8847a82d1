// mufu_bench.cu — MUFU throughput on sm_100a: warp-instruction issue interval per SM sub-partition for
// tanh.approx / ex2.approx / rcp.approx, with 1..8 warps per sub-partition (behind the LSTM gate-math budget in DESIGN.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_bench tools/mufu_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else y = fmaf(x, 1.0001f, 0.5f);
  return y;
}

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 4096;
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<OP><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)(warps / 4) * iters * 8;          // warp instructions per sub-partition
    printf("%-6s %2d warps/SM: %.2f cycles per warp-instruction per sub-partition (%.1f lanes/clk/SM)\n", name, warps,
           c / per_smsp, 32.0 * 4 * per_smsp / c);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0>("tanh", out, cyc); run<1>("ex2", out, cyc); run<2>("rcp", out, cyc); run<3>("ffma", out, cyc);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
