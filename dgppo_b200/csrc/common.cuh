// Shared device helpers for libdgppo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dgppo_abi.h"

namespace dgppo {

// --- individually rounded fp32 arithmetic --------------------------------
// Every value that feeds a mask comparison, an index, or an env state is
// computed with these: nvcc may not contract them into FFMA, and div/sqrt are
// IEEE-rounded regardless of -prec-div / fast-math flags.  This is the same
// arithmetic NumPy performs in oracle/env_np.py, which makes the env kernels
// comparable bit for bit.
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// jnp.linalg.norm of a 2-vector: sqrt(dx*dx + dy*dy)
__device__ __forceinline__ float norm2(float dx, float dy) {
  return fsqrt(fadd(fmul(dx, dx), fmul(dy, dy)));
}
// jnp.clip(x, lo, hi) = min(max(x, lo), hi) for finite x
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);
}
// NaN-propagating min (XLA / NumPy reduce-min semantics)
__device__ __forceinline__ float nanmin(float a, float b) {
  return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
}
// NaN-propagating max (jnp.maximum / jnp.max semantics; fmaxf would drop the NaN)
__device__ __forceinline__ float nanmax(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));       // one FMNMX.NAN
  return r;
}

struct GraphDims {
  int sd, nd, n_on, N, E, n_ag, n_ao, n, g;
};

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may start while its
// predecessor on the stream is still running; everything before pdl_wait() (weight staging into shared
// memory) overlaps the predecessor's tail, pdl_wait() returns once the predecessor has completed and
// its writes are visible.  pdl_launch_dependents() lets the NEXT kernel's blocks be scheduled early.
// Both are no-ops for a kernel launched without the attribute / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__host__ __device__ inline bool is_lidar(int kind) {
  return kind == DGPPO_ENV_LIDAR_SPREAD || kind == DGPPO_ENV_LIDAR_TARGET ||
         kind == DGPPO_ENV_LIDAR_BICYCLE_TARGET || kind == DGPPO_ENV_LIDAR_LINE;
}
__host__ __device__ inline bool is_mpe(int kind) { return !is_lidar(kind); }
// paired goals (one goal per agent, one agent-goal edge each); every other kind connects each agent to every goal node
__host__ __device__ inline bool is_target(int kind) {
  return kind == DGPPO_ENV_LIDAR_TARGET || kind == DGPPO_ENV_LIDAR_BICYCLE_TARGET || kind == DGPPO_ENV_MPE_TARGET;
}
__host__ __device__ inline bool is_spread(int kind) { return !is_target(kind); }
__host__ __device__ inline bool is_line(int kind) { return kind == DGPPO_ENV_LIDAR_LINE || kind == DGPPO_ENV_MPE_LINE; }
// goal NODES (landmarks for the Line / Formation families: lidar_line.py:36, mpe_line.py:36, mpe_formation.py:36)
__host__ __device__ inline int n_goals_of(int kind, int n) {
  return is_line(kind) ? 2 : (kind == DGPPO_ENV_MPE_FORMATION ? 1 : n);
}
__host__ __device__ inline int n_cost_of(int kind) { return kind == DGPPO_ENV_MPE_CONNECT_SPREAD ? 3 : 2; }
// agent-obstacle edges always on, y range doubled (mpe_corridor.py:64-67,93; mpe_connect_spread.py:140-143,168)
__host__ __device__ inline bool is_tall_mpe(int kind) {
  return kind == DGPPO_ENV_MPE_CORRIDOR || kind == DGPPO_ENV_MPE_CONNECT_SPREAD;
}
__host__ __device__ inline bool is_bicycle(int kind) { return kind == DGPPO_ENV_LIDAR_BICYCLE_TARGET; }

__host__ __device__ inline GraphDims graph_dims(const DgppoEnvCfg& c) {
  GraphDims d;
  d.n = c.n_agents;
  d.g = n_goals_of(c.kind, c.n_agents);
  d.sd = is_bicycle(c.kind) ? 5 : 4;
  d.nd = d.sd + 3;
  const bool lid = is_lidar(c.kind);
  d.n_on = (c.n_obs > 0) ? (lid ? c.top_k * c.n_agents : c.n_obs) : 0;
  d.N = d.n + d.g + d.n_on + 1;
  d.n_ag = is_spread(c.kind) ? d.g : 1;
  d.n_ao = (c.n_obs > 0) ? (lid ? c.top_k : c.n_obs) : 0;
  d.E = d.n * d.n + d.n * d.n_ag + d.n * d.n_ao;
  return d;
}

// K2 with predict = 1 casts the rays from the position the agents will have AFTER the coming step
// (computed from the current state, which is all it depends on); dgppo_lidar is predict = 0.
// a_pitch / h_pitch: agent / hits are slots of (b, pitch, ...) records (1: plain batches).
int launch_lidar(void* stream, const DgppoEnvCfg* cfg, const float* agent, const float* obstacles,
                 const float* ray_dirs, float* hits, int32_t b, int predict, int32_t a_pitch = 1, int32_t h_pitch = 1);
// K1 with agent / next_agent / Lidar hits addressed as slots of (b, st_pitch, ...) records.
int launch_env_step(void* stream, const DgppoEnvCfg* cfg, const float* agent, const float* goal,
                    const float* obs_nodes, const float* action, float* next_agent, float* reward,
                    float* cost, int32_t io_pitch, int32_t b, int32_t st_pitch);

inline int check_env_cfg(const DgppoEnvCfg* c) {
  if (!c) return DGPPO_EINVAL;
  if (c->kind < 0 || c->kind > 9) return DGPPO_ENOTSUP;
  if (c->kind == DGPPO_ENV_MPE_CORRIDOR && c->n_obs != 2) return DGPPO_EINVAL;       // mpe_corridor.py:33-35
  if (c->kind == DGPPO_ENV_MPE_CONNECT_SPREAD && c->n_obs != 1) return DGPPO_EINVAL; // mpe_connect_spread.py:38-40
  if (is_line(c->kind) && c->n_agents < 2) return DGPPO_EINVAL;                      // n - 1 intervals between the landmarks
  if (c->kind == DGPPO_ENV_MPE_FORMATION && !c->goal_table) return DGPPO_EINVAL;
  if (c->n_agents < 1 || c->n_obs < 0) return DGPPO_EINVAL;
  if (is_lidar(c->kind) && c->n_obs > 0) {
    if (c->n_rays < 1 || c->n_rays > 1024) return DGPPO_ENOTSUP;
    if (c->top_k < 1 || c->top_k > c->n_rays) return DGPPO_EINVAL;
  }
  return 0;
}

}  // namespace dgppo
