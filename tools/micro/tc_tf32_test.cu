// Micro-test: tcgen05.mma kind::tf32 with the 3xTF32 split (A_hi B_hi + A_lo B_hi + A_hi B_lo), operands in
// shared memory in the no-swizzle K-major canonical layout, accumulator in TMEM, read back with tcgen05.ld.
// Pins the shared-memory descriptor fields (LBO / SBO meaning) and the instruction descriptor on real
// hardware before the head kernel depends on them, reports the error against an fp64 product (and the
// error of a plain fp32 FMA loop beside it), and times the issue -> commit -> wait -> tcgen05.ld chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_tf32_test tc_tf32_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#include "../../dgppo_b200/csrc/tc_common.cuh"

using namespace dgppo::tc;

// A: [kc][128][4] hi then lo (fp32 canonical, K = 64 -> 16 chunks); B: [kc][N][4] hi then lo.
// LBO = K-direction chunk stride, SBO = 8-row group stride (128 B) (the swapped reading faults: measured).
// rounds: how many times the 24-MMA sequence is issued per commit (timing only; rounds > 1 re-accumulates).
__global__ void __launch_bounds__(128, 1)
tc_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ D, int N, int rounds, int ts,
          int reps, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);                 // 2 x 32 KB
  float* sB = sA + 2 * 128 * 64;                                  // 2 x N x 64 floats
  __shared__ __align__(8) unsigned long long bar_tma, bar_mma;
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned a_bytes = 2u * 128 * 64 * 4, b_bytes = 2u * N * 64 * 4;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, a_bytes + b_bytes);
    bulk_g2s(sA, Ag, a_bytes, &bar_tma);
    bulk_g2s(sB, Bg, b_bytes, &bar_tma);
  }
  if (!mbar_wait(&bar_tma, 0)) { if (tid == 0) atomicExch(err, 1); }
  __syncthreads();
  if (ts) {      // A operand into TMEM: thread = row, hi at columns 256 + k, lo at columns 320 + k
    for (int part = 0; part < 2; ++part)
      for (int c0 = 0; c0 < 64; c0 += 16) {
        unsigned v[16];
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(sA[part * 128 * 64 + ((c0 + j) / 4 * 128 + tid) * 4 + ((c0 + j) & 3)]);
        tmem_st16(tmem + ((unsigned)(warp * 32) << 16) + 256 + part * 64 + c0, v);
      }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
  }
  const unsigned lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
  const unsigned idesc = make_idesc_tf32(128, N);
  long long t0 = clock64();
  unsigned phase = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      tc_fence_after();
      for (int rd = 0; rd < rounds; ++rd)
      for (int part = 0; part < 3; ++part) {
        const float* a = sA + (part == 1 ? 128 * 64 : 0);         // part 1: A_lo B_hi
        const float* b = sB + (part == 2 ? N * 64 : 0);           // part 2: A_hi B_lo
        for (int ks = 0; ks < 8; ++ks) {
          const unsigned long long da = make_sdesc(a + ks * 2 * 128 * 4, lbo_a, sbo);
          const unsigned long long db = make_sdesc(b + ks * 2 * N * 4, lbo_b, sbo);
          if (ts) mma_tf32_ts(tmem, tmem + 256 + (part == 1 ? 64 : 0) + ks * 8, db, idesc, (part | ks | rd) != 0);
          else mma_tf32_ss(tmem, da, db, idesc, (part | ks | rd) != 0);
        }
      }
      mma_commit(&bar_mma);
    }
    if (!mbar_wait(&bar_mma, phase)) { if (tid == 0) atomicExch(err, 2); break; }
    phase ^= 1;
    tc_fence_after();
    if (rep + 1 < reps) {      // the dependent chain of the real kernel: read a little, then go again
      unsigned v[16];
      tmem_ld16(tmem + ((unsigned)(warp * 32) << 16), v);
      tmem_ld_wait();
      if (v[0] == 0x7fc12345u) atomicExch(err, 3);
      tc_fence_before();
      __syncthreads();
    }
  }
  long long t1 = clock64();
  if (tid == 0) *cycles = t1 - t0;
  // epilogue: thread = row
  for (int c0 = 0; c0 < N; c0 += 16) {
    unsigned v[16];
    tmem_ld16(tmem + ((unsigned)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float host_tf32_rna(float x) {
  unsigned u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xffffe000u;
  float r; memcpy(&r, &u, 4); return r;
}

int main() {
  const int M = 128, K = 64;
  int ok_all = 1;
  for (int cfg = 0; cfg < 9; ++cfg) {
    const int ts = cfg >= 6;
    const int N = (cfg == 2) ? 192 : (cfg == 3 ? 256 : (cfg == 4 ? 128 : (cfg == 5 ? 16 : (cfg == 7 ? 192 : (cfg == 8 ? 256 : 64)))));
    const int rounds = (cfg == 1) ? 4 : 1;
    std::vector<float> X(M * K), W(K * N);
    srand(1234 + cfg);
    for (auto& v : X) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.25f;
    std::vector<float> A(2 * M * K), B(2 * N * K);
    for (int kc = 0; kc < K / 4; ++kc)
      for (int r = 0; r < M; ++r)
        for (int j = 0; j < 4; ++j) {
          const float x = X[r * K + 4 * kc + j], hi = host_tf32_rna(x);
          A[(kc * M + r) * 4 + j] = hi; A[M * K + (kc * M + r) * 4 + j] = host_tf32_rna(x - hi);
        }
    for (int kc = 0; kc < K / 4; ++kc)
      for (int n = 0; n < N; ++n)
        for (int j = 0; j < 4; ++j) {
          const float w = W[(4 * kc + j) * N + n], hi = host_tf32_rna(w);
          B[(kc * N + n) * 4 + j] = hi; B[N * K + (kc * N + n) * 4 + j] = host_tf32_rna(w - hi);
        }
    float *dA, *dB, *dD; long long* dcy; int* derr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, M * N * 4);
    cudaMalloc(&dcy, 8); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4); cudaMemset(derr, 0, 4);
    const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4 + 1024;
    cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int reps : {1, 101}) {
      tc_kernel<<<1, 128, smem>>>(dA, dB, dD, N, (reps == 1) ? 1 : rounds, ts, reps, dcy, derr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cfg %d: CUDA error %s\n", cfg, cudaGetErrorString(e)); return 1; }
      long long cy; int err;
      cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost);
      std::vector<float> Dh(M * N);
      cudaMemcpy(Dh.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
      double emax = 0, e32max = 0, ref_max = 0;
      for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
          double ref = 0; float f32 = 0.f;
          for (int k = 0; k < K; ++k) { ref += (double)X[r * K + k] * W[k * N + n]; f32 = fmaf(X[r * K + k], W[k * N + n], f32); }
          emax = fmax(emax, fabs(Dh[r * N + n] - ref)); e32max = fmax(e32max, fabs(f32 - ref)); ref_max = fmax(ref_max, fabs(ref));
        }
      printf("cfg %d ts=%d N=%d rounds=%d reps=%d err_flag=%d: max|D-ref|=%.3e (fp32 fma loop %.3e, max|ref|=%.2f) cycles/rep=%.0f\n",
             cfg, ts, N, (reps == 1) ? 1 : rounds, reps, err, emax, e32max, ref_max, (double)cy / reps);
      if (reps == 1 && (emax > 4e-6 || err)) ok_all = 0;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dcy); cudaFree(derr);
  }
  printf(ok_all ? "TC_TF32_TEST PASS\n" : "TC_TF32_TEST FAIL\n");
  return ok_all ? 0 : 2;
}
