// Micro-benchmark: legacy mma.sync m16n8k8 TF32 throughput on sm_100a (registers only).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) kern(float* out, int iters) {
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3f800000u, 0x3e800000u};
  unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int occ = 1; occ <= 8; occ *= 2) {
    const int grid = 148 * occ;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      kern<<<grid, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double mac = (double)grid * 8 /*warps*/ * iters * 8 /*mma*/ * 16 * 8 * 8;
    printf("CTAs/SM %d: %.3f ms  %.1f TFLOP/s tf32 (dense)  = %.0f MAC/clk/SM @1.965GHz\n", occ, ms, 2 * mac / ms * 1e-9,
           mac / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
