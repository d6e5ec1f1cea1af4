// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 x = make_float2(a, a * 0.999f), y = make_float2(b, b * 1.001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, x.x, y.x); acc[i].y = fmaf(acc[i].y, x.y, y.y); }
      else acc[i] = __ffma2_rn(acc[i], x, y);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 8;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) kern<0><<<grid, 256>>>(out, iters, 0.9999f, 1e-4f);
      else kern<1><<<grid, 256>>>(out, iters, 0.9999f, 1e-4f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)grid * 256 * iters * 16;
      if (rep == 2) printf("mode %d (%s): %.3f ms  %.1f TFLOP/s (FMA=2 flops)\n", mode, mode ? "FFMA2" : "FFMA", ms,
                           2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
