// K5: Dec-OCP GAE and the CBF-residual advantage merge for sm_100a.
#include "common.cuh"

namespace dgppo {

// ---------------------------------------------------------------------- GAE
// compute_dec_ocp_gae (dgppo/algo/utils.py:11-79), literal O(T^2) DP.
// One thread per (trajectory, agent, column) with column in [0, nh] (nh cost
// columns + the Vl column); the thread's row of T+1 partial returns lives in
// shared memory laid out [k][column] so a warp's accesses are conflict-free.
// gae coefficients at iteration ii (dgppo/algo/utils.py:57-60):
//   c[0] = lambda^ii, c[k] = lambda^(ii-k) (1-lambda) for 1 <= k <= ii.
__global__ void gae_kernel(const float* __restrict__ hs, const float* __restrict__ l,
                           const float* __restrict__ Vh, const float* __restrict__ Vl,
                           float gamma, float lam, float* __restrict__ Qh, float* __restrict__ Ql,
                           int b, int T, int n, int nh, int traj_per_cta) {
  extern __shared__ float smem[];
  const int cols = n * (nh + 1);
  float* lam_pow = smem;                         // [T+1]
  float* rows = smem + (T + 1);                  // [traj_per_cta][T+1][cols]
  for (int i = threadIdx.x; i <= T; i += blockDim.x) lam_pow[i] = powf(lam, (float)i);
  __syncthreads();
  const int tl = threadIdx.x / cols, col = threadIdx.x - tl * cols;
  const int traj = blockIdx.x * traj_per_cta + tl;
  if (tl >= traj_per_cta || traj >= b) return;
  const int a = col / (nh + 1), h = col - a * (nh + 1);
  const bool is_l = (h == nh);
  if (is_l && a != 0) return;                    // Ql is read from agent 0 (algo/utils.py:78)
  float* row = rows + (size_t)tl * (T + 1) * cols + col;   // row[k*cols]
  const float* hs_t = hs + (size_t)traj * T * n * nh;
  const float* Vh_t = Vh + (size_t)traj * (T + 1) * n * nh;
  const float* l_t = l + (size_t)traj * T;
  const float* Vl_t = Vl + (size_t)traj * (T + 1);
  const float one_m_g = 1.f - gamma, one_m_lam = 1.f - lam;

  row[0] = is_l ? Vl_t[T] : Vh_t[((size_t)T * n + a) * nh + h];
  for (int ii = 0; ii < T; ++ii) {
    const int t = T - 1 - ii;
    float hval = 0.f, disc = 0.f, lval = 0.f;
    if (is_l) {
      lval = l_t[t];
    } else {
      const float* hrow = hs_t + ((size_t)t * n + a) * nh;
      hval = hrow[h];
      float hmax = hrow[0];
      for (int q = 1; q < nh; ++q) hmax = nanmax(hmax, hrow[q]);
      disc = one_m_g * hmax;
    }
    float acc = 0.f;
    for (int k = 0; k <= ii; ++k) {
      const float prev = row[k * cols];
      const float v = is_l ? (lval + gamma * prev) : nanmax(hval, disc + gamma * prev);
      row[k * cols] = v;
      const float c = (k == 0) ? lam_pow[ii] : lam_pow[ii - k] * one_m_lam;
      acc = fmaf(v, c, acc);
    }
    if (is_l) { if (a == 0) Ql[(size_t)traj * T + t] = acc; }
    else Qh[(((size_t)traj * T + t) * n + a) * nh + h] = acc;
    row[(ii + 1) * cols] = is_l ? Vl_t[t] : Vh_t[((size_t)t * n + a) * nh + h];
  }
}

// ---------------------------------------------------- CBF advantage merge
// dgppo/algo/dgppo.py:239-259.  One warp per trajectory: lanes stride over T
// for the Al standardisation (mean / population std over T), then over
// (t, agent) for the residual and merge.
__global__ void cbf_advantage_kernel(const float* __restrict__ Ql, const float* __restrict__ Vl,
                                     const float* __restrict__ Vh, float inv_dt_unused, float dt,
                                     float alpha, float cbf_eps, float cbf_weight,
                                     float* __restrict__ A, float* __restrict__ deriv_out,
                                     float* __restrict__ acbf_out, uint8_t* __restrict__ safe_out,
                                     int b, int T, int n, int nh) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= b) return;
  const float* ql = Ql + (size_t)warp * T;
  const float* vl = Vl + (size_t)warp * (T + 1);
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += ql[t] - vl[t];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)T;
  float v = 0.f;
  for (int t = lane; t < T; t += 32) { const float d = (ql[t] - vl[t]) - mean; v = fmaf(d, d, v); }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float stdv = sqrtf(v / (float)T);
  const float* vh = Vh + (size_t)warp * (T + 1) * n * nh;
  for (int idx = lane; idx < T * n; idx += 32) {
    const int t = idx / n, a = idx - t * n;
    const float al = ((ql[t] - vl[t]) - mean) / (stdv + 1e-8f);
    bool safe = true;
    float amax = -INFINITY;
    for (int h = 0; h < nh; ++h) {
      const float v0 = vh[((size_t)t * n + a) * nh + h], v1 = vh[((size_t)(t + 1) * n + a) * nh + h];
      const float d = (v1 - v0) / dt + alpha * v0;
      const float ac = nanmax(d + cbf_eps, 0.f);            // jnp.maximum: a NaN residual stays NaN
      safe = safe && (d <= 0.f);
      amax = nanmax(amax, ac);
      const size_t o = (((size_t)warp * T + t) * n + a) * nh + h;
      if (deriv_out) deriv_out[o] = d;
      if (acbf_out) acbf_out[o] = ac;
    }
    const size_t o = ((size_t)warp * T + t) * n + a;
    A[o] = -((safe ? al : 0.f) + amax * cbf_weight);
    if (safe_out) safe_out[o] = safe ? 1 : 0;
  }
}

}  // namespace dgppo

using namespace dgppo;

extern "C" int dgppo_gae(void* stream, const float* hs, const float* l, const float* Vh,
                         const float* Vl, float gamma, float gae_lambda, float* Qh, float* Ql,
                         int32_t b, int32_t T, int32_t n, int32_t nh) {
  if (b == 0) return 0;
  if (b < 0 || T < 1 || n < 1 || nh < 1 || !hs || !l || !Vh || !Vl || !Qh || !Ql) return DGPPO_EINVAL;
  const int cols = n * (nh + 1);
  if (cols > 1024) return DGPPO_ENOTSUP;
  const size_t per_traj = (size_t)(T + 1) * cols * sizeof(float);
  const size_t budget = 200 * 1024 - (size_t)(T + 1) * sizeof(float);
  if (per_traj > budget) return DGPPO_ENOTSUP;
  int tpc = (int)(budget / per_traj);
  const int by_threads = 256 / cols > 0 ? 256 / cols : 1;
  if (tpc > by_threads) tpc = by_threads;
  if (tpc > 4) tpc = 4;
  if (tpc < 1) tpc = 1;
  int threads = ((tpc * cols + 31) / 32) * 32;
  const size_t smem = (size_t)(T + 1) * sizeof(float) + tpc * per_traj;
  cudaError_t err = cudaFuncSetAttribute(gae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  const int grid = (b + tpc - 1) / tpc;
  gae_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(hs, l, Vh, Vl, gamma, gae_lambda, Qh, Ql,
                                                             b, T, n, nh, tpc);
  return (int)cudaGetLastError();
}

extern "C" int dgppo_cbf_advantage(void* stream, const float* Ql, const float* Vl, const float* Vh,
                                   float dt, float alpha, float cbf_eps, float cbf_weight, float* A,
                                   float* cbf_deriv, float* Acbf, uint8_t* is_safe, int32_t b,
                                   int32_t T, int32_t n, int32_t nh) {
  if (b == 0) return 0;
  if (b < 0 || T < 1 || n < 1 || nh < 1 || !Ql || !Vl || !Vh || !A) return DGPPO_EINVAL;
  const int threads = 128, warps = threads / 32;
  const int grid = (b + warps - 1) / warps;
  cbf_advantage_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(Ql, Vl, Vh, 0.f, dt, alpha, cbf_eps,
                                                                   cbf_weight, A, cbf_deriv, Acbf, is_safe,
                                                                   b, T, n, nh);
  return (int)cudaGetLastError();
}
