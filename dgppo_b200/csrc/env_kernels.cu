// K1 (dynamics + reward + cost), K2 (LiDAR), K3 (radius graph) for sm_100a.
//
// All three are HBM-/latency-bound integer-and-fp32 kernels: one warp (K1,
// K2) or one CTA (K3) per environment, per-env tiles staged in shared memory,
// coalesced float4 / int stores for the graph record.  Arithmetic follows the
// reference op by op with individually rounded fp32 operations (common.cuh),
// so masks, indices and states match oracle/env_np.py bit for bit.
#include <stdlib.h>

#include "common.cuh"

namespace dgppo {

struct EnvConsts {
  int kind, n, n_obs, n_rays, top_k;
  float dt, R, R_diag, R_obs, R_mpe_obs, car2, car, car_obs, d2g, connect_r;
  float lo[5], hi[5];
  const float* goal_table;     // MPEFormation: (n, 2) goal offsets from the landmark
};

static EnvConsts make_consts(const DgppoEnvCfg& c) {
  EnvConsts k;
  k.kind = c.kind; k.n = c.n_agents; k.n_obs = c.n_obs; k.n_rays = c.n_rays; k.top_k = c.top_k;
  k.dt = (float)c.dt;
  k.R = (float)c.comm_radius;
  k.R_diag = (float)(c.comm_radius + 1.0);          // lidar_spread.py:64
  k.R_obs = (float)(c.comm_radius - 1e-1);          // lidar_spread.py:87
  // MPE agent-obstacle edges: within comm_radius (mpe_spread.py:73-75); always on in the corridor (x100: mpe_corridor.py:93)
  // (and in the connect-spread env: mpe_connect_spread.py:168)
  k.R_mpe_obs = (float)(is_tall_mpe(c.kind) ? c.comm_radius * 100 : c.comm_radius);
  k.connect_r = (float)c.connect_radius;            // mpe_connect_spread.py:117
  k.goal_table = c.goal_table;
  k.car2 = (float)(c.car_radius * 2.0);             // lidar_env/base.py:188
  k.car = (float)c.car_radius;                      // lidar_env/base.py:197
  k.car_obs = (float)(c.car_radius + c.obs_radius); // mpe/base.py:181
  k.d2g = (float)c.dist2goal;
  const float A = (float)c.area_size;
  if (is_bicycle(c.kind)) {                         // lidar_bicycle_target.py:120-123
    const float lo[5] = {0.f, 0.f, -1.f, -1.f, -0.5f}, hi[5] = {A, A, 1.f, 1.f, 0.5f};
    for (int i = 0; i < 5; ++i) { k.lo[i] = lo[i]; k.hi[i] = hi[i]; }
  } else {
    const float v = is_mpe(c.kind) ? 1.0f : 0.5f;   // mpe/base.py:243-246 | lidar_env/base.py:273-276
    const float Ay = is_tall_mpe(c.kind) ? (float)(c.area_size * 2) : A;   // mpe_corridor.py:64-67, mpe_connect_spread.py:140-143
    const float lo[5] = {0.f, 0.f, -v, -v, 0.f}, hi[5] = {A, Ay, v, v, 0.f};
    for (int i = 0; i < 5; ++i) { k.lo[i] = lo[i]; k.hi[i] = hi[i]; }
  }
  return k;
}

// Next position of one agent (the first two components of agent_step_euler + clip_state).  It depends on
// the CURRENT state only - the action moves the velocity / heading, not the position - which is what lets
// the LiDAR of step t + 1 run ahead of the policy of step t (rollout.cu).  K1 and K2 both call this, so the
// bits agree.
__device__ __forceinline__ void next_position(const EnvConsts& k, const float* s, int sd, float& x, float& y) {
  float ox, oy;
  if (sd == 5) {                                             // lidar_bicycle_target.py:96-103
    const float theta = atan2f(s[3], s[2]);
    ox = fadd(s[0], fmul(fmul(s[4], cosf(theta)), k.dt));
    oy = fadd(s[1], fmul(fmul(s[4], sinf(theta)), k.dt));
  } else {                                                   // lidar_env/base.py:146-147
    ox = fadd(fmul(s[2], k.dt), s[0]);
    oy = fadd(fmul(s[3], k.dt), s[1]);
  }
  x = clampf(ox, k.lo[0], k.hi[0]);
  y = clampf(oy, k.lo[1], k.hi[1]);
}

// ------------------------------------------------------------------- K1
// One warp per environment; lanes stride over agents / goals.
constexpr int K1_WARPS = 4;

__global__ void __launch_bounds__(K1_WARPS * 32)
env_step_kernel(EnvConsts k, const float* __restrict__ agent, const float* __restrict__ goal,
                const float* __restrict__ obs_nodes, const float* __restrict__ action,
                float* __restrict__ next_agent, float* __restrict__ reward,
                float* __restrict__ cost, int io_pitch, int b, int st_pitch) {
  extern __shared__ float smem[];
  pdl_launch_dependents();        // the next policy forward (or LiDAR) may be scheduled and stage its weights
  pdl_wait();                     // launched with programmatic serialization: the actions of the head are visible from here
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * K1_WARPS + warp;
  if (env >= b) return;
  const int n = k.n, sd = is_bicycle(k.kind) ? 5 : 4;
  float* px = smem + warp * (5 * n);
  float* py = px + n;
  float* an2 = py + n;     // |clipped action|^2
  float* d2g = an2 + n;    // per-goal distance
  float* far = d2g + n;

  // st_pitch: agent / next_agent / Lidar hits are slots of a (b, st_pitch, ...) record (1: plain batches)
  const float* ag = agent + (size_t)env * st_pitch * n * sd;
  const float* ac = action + (size_t)env * io_pitch * n * 2;
  float* nx = next_agent + (size_t)env * st_pitch * n * sd;

  for (int i = lane; i < n; i += 32) {
    float s[5];
    for (int c = 0; c < sd; ++c) s[c] = ag[i * sd + c];
    const float a0 = clampf(ac[i * 2 + 0], -1.f, 1.f);      // clip_action, env/base.py:84-86
    const float a1 = clampf(ac[i * 2 + 1], -1.f, 1.f);
    px[i] = s[0]; py[i] = s[1];
    const float nrm = norm2(a0, a1);
    an2[i] = fmul(nrm, nrm);
    float o[5];
    if (sd == 5) {                                           // lidar_bicycle_target.py:96-106
      const float theta = atan2f(s[3], s[2]);
      const float theta_next = fadd(theta, fmul(fmul(fmul(s[4], a0), k.dt), 10.f));
      o[2] = cosf(theta_next);
      o[3] = sinf(theta_next);
      o[4] = fadd(s[4], fmul(fmul(a1, k.dt), 10.f));
    } else {                                                 // lidar_env/base.py:146-147
      o[2] = fadd(fmul(fmul(a0, 10.f), k.dt), s[2]);
      o[3] = fadd(fmul(fmul(a1, 10.f), k.dt), s[3]);
    }
    next_position(k, s, sd, nx[i * sd + 0], nx[i * sd + 1]);
    for (int c = 2; c < sd; ++c) nx[i * sd + c] = clampf(o[c], k.lo[c], k.hi[c]);
  }
  __syncwarp();

  // cost on the pre-step state (lidar_env/base.py:180-207, mpe/base.py:164-191, mpe_connect_spread.py:105-138)
  const bool lid = is_lidar(k.kind);
  const int nh = n_cost_of(k.kind);
  const bool clip_hi = lid || nh == 3;           // a_max = 1 as well (lidar_env/base.py:205, mpe_connect_spread.py:135)
  float connect = -INFINITY;                     // max_i (min_dist_i - connect_radius), the third cost of every agent
  if (nh == 3) {
    for (int i = lane; i < n; i += 32) {
      float mind = INFINITY;
      for (int j = 0; j < n; ++j) {
        float d = norm2(fsub(px[i], px[j]), fsub(py[i], py[j]));
        d = fadd(d, (i == j) ? 1e6f : 0.f);
        mind = fminf(mind, d);
      }
      connect = fmaxf(connect, fsub(mind, k.connect_r));
    }
    for (int o = 16; o; o >>= 1) connect = fmaxf(connect, __shfl_xor_sync(0xffffffffu, connect, o));
    connect = (connect <= 0.f) ? fsub(connect, 0.5f) : fadd(connect, 0.5f);
    connect = clampf(connect, -1.f, 1.f);
  }
  for (int i = lane; i < n; i += 32) {
    const float xi = px[i], yi = py[i];
    float mind = INFINITY;
    for (int j = 0; j < n; ++j) {
      float d = norm2(fsub(xi, px[j]), fsub(yi, py[j]));
      d = fadd(d, (i == j) ? 1e6f : 0.f);
      mind = fminf(mind, d);
    }
    float c0 = fsub(k.car2, mind);
    float c1 = 0.f;
    if (k.n_obs > 0) {
      float mo = INFINITY;
      if (lid) {
        const float* h = obs_nodes + ((size_t)env * st_pitch * n + i) * k.top_k * 2;
        for (int q = 0; q < k.top_k; ++q)
          mo = nanmin(mo, norm2(fsub(h[2 * q], xi), fsub(h[2 * q + 1], yi)));
        c1 = fsub(k.car, mo);
      } else {
        const float* ob = obs_nodes + (size_t)env * k.n_obs * 4;
        for (int q = 0; q < k.n_obs; ++q)
          mo = fminf(mo, norm2(fsub(xi, ob[4 * q]), fsub(yi, ob[4 * q + 1])));
        c1 = fsub(k.car_obs, mo);
      }
    }
    c0 = (c0 <= 0.f) ? fsub(c0, 0.5f) : fadd(c0, 0.5f);
    c1 = (c1 <= 0.f) ? fsub(c1, 0.5f) : fadd(c1, 0.5f);
    if (clip_hi) { c0 = clampf(c0, -1.f, 1.f); c1 = (c1 != c1) ? c1 : clampf(c1, -1.f, 1.f); }
    else         { c0 = fmaxf(c0, -1.f);       c1 = fmaxf(c1, -1.f); }
    float* co = cost + ((size_t)env * io_pitch * n + i) * nh;
    co[0] = c0; co[1] = c1;
    if (nh == 3) co[2] = connect;
  }

  // reward (lidar_spread.py:35-52, lidar_target.py:35-52; Line / Formation: the n goals derived from the
  // landmark nodes by landmark2goal, lidar_line.py:128-150, mpe_line.py:119-152, mpe_formation.py:93-116)
  const int g_nodes = n_goals_of(k.kind, n);
  const float* gl = goal + (size_t)env * g_nodes * sd;
  const bool spread = is_spread(k.kind);
  const bool line = is_line(k.kind), formation = k.kind == DGPPO_ENV_MPE_FORMATION;
  const bool short_line = k.kind == DGPPO_ENV_MPE_LINE && n <= 3;             // mpe_line.py:121-124
  for (int q = lane; q < n; q += 32) {
    float gx, gy;
    if (line) {        // l0 + k * (l1 - l0) / n_interval, op by op
      const float dx = fsub(gl[sd], gl[0]), dy = fsub(gl[sd + 1], gl[1]);
      const float kq = (float)(short_line ? q + 1 : q), ni = (float)(short_line ? n + 1 : n - 1);
      gx = fadd(gl[0], fdiv(fmul(kq, dx), ni));
      gy = fadd(gl[1], fdiv(fmul(kq, dy), ni));
    } else if (formation) {
      gx = fadd(gl[0], __ldg(k.goal_table + 2 * q));
      gy = fadd(gl[1], __ldg(k.goal_table + 2 * q + 1));
    } else {
      gx = gl[q * sd]; gy = gl[q * sd + 1];
    }
    float d;
    if (spread) {
      d = INFINITY;
      for (int j = 0; j < n; ++j) d = fminf(d, norm2(fsub(gx, px[j]), fsub(gy, py[j])));
    } else {
      d = norm2(fsub(gx, px[q]), fsub(gy, py[q]));
    }
    d2g[q] = d;
    far[q] = (d > k.d2g) ? 1.f : 0.f;
  }
  __syncwarp();
  if (lane == 0) {                       // left-to-right means (oracle seq_mean)
    float s0 = d2g[0], s1 = far[0], s2 = an2[0];
    for (int j = 1; j < n; ++j) { s0 = fadd(s0, d2g[j]); s1 = fadd(s1, far[j]); s2 = fadd(s2, an2[j]); }
    const float fn = (float)n;
    float r = 0.f;
    r = fsub(r, fmul(fdiv(s0, fn), 0.01f));
    r = fsub(r, fmul(fdiv(s1, fn), 0.001f));
    r = fsub(r, fmul(fdiv(s2, fn), 0.0001f));
    reward[(size_t)env * io_pitch] = r;
  }
}

// ------------------------------------------------------------------- K2
// One warp per (environment, agent).
//   phase 1  lane = obstacle edge: the ray-independent terms of the 2x2 solve
//            (edge vector, agent - corner, numerator of alpha) go to the warp's smem,
//            edges within the sensing range packed to the front, the others to the back
//            (conservative distance test);
//   phase 2  lane = ray: a far edge can only matter when the ray is (numerically) parallel
//            to it, so it costs one determinant; a near edge is classified without dividing
//            (lidar_slot_near): clearly invalid -> nothing, clearly valid -> exact fraction
//            minimum (one division per ray at the end), borderline -> the literal arithmetic
//            of obstacle.py:82-104;
//   phase 3  stable top-k by counting rank over 64-bit (alpha bits, ray) keys.
// Every value that reaches the output is produced by the same individually
// rounded operations as the reference expression (bit-exact vs the oracle).
constexpr int K2_WARPS = 4;

__device__ __forceinline__ float lidar_slot_literal(float dx12, float dy12, float dx43, float dy43,
                                                    float dx13, float dy13, float na, float det_raw) {
  float det = det_raw;
  const float sg = (det > 0.f) ? 1.f : ((det < 0.f) ? -1.f : det);   // sign(0)=0, sign(NaN)=NaN
  det = fmul(sg, fminf(fmaxf(fabsf(det), 1e-7f), 1e7f));
  const float nb = fadd(fmul(-dy12, dx13), fmul(dx12, dy13));
  // Division-free early out.  With 1e-7 <= |det| <= 1e7 the rounded quotient q = n/det is
  // certainly > 1 when |n| > 1.0001 |det| (same sign) and certainly < 0 when the signs differ
  // and |n| > 1e-30 (|q| >= 1e-37, it cannot round to -0): the slot is then invalid and
  // contributes exactly 0*q + 1*1e6 = 1e6, as the literal expression would.  det == 0 or NaN
  // (exactly parallel ray: q = +-inf / NaN) always takes the literal path below.
  const float ad = fabsf(det);
  const bool regular = ad >= 1e-7f;                     // false for 0 and NaN
  const bool out_a = ((na > 0.f) == (det > 0.f)) ? (fabsf(na) > 1.0001f * ad) : (fabsf(na) > 1e-30f);
  const bool out_b = ((nb > 0.f) == (det > 0.f)) ? (fabsf(nb) > 1.0001f * ad) : (fabsf(nb) > 1e-30f);
  if (regular && (out_a || out_b)) return 1e6f;
  const float alpha = fdiv(na, det);
  const float beta = fdiv(nb, det);
  const float vf = (alpha <= 1.f && alpha >= 0.f && beta <= 1.f && beta >= 0.f) ? 1.f : 0.f;
  return fadd(fmul(vf, alpha), fmul(fsub(1.f, vf), 1e6f));
}

// A near edge for one ray.  Three outcomes, decided without dividing:
//   clearly invalid  (alpha or beta certainly outside [0, 1], see lidar_slot_literal): the slot
//       contributes 1e6, which never changes the running minimum: nothing to do;
//   clearly valid    (1e-30 < n < 0.9999 |det| for both numerators, taken over the positive
//       denominator |det|, 1e-7 <= |det| <= 1e7): both rounded quotients lie strictly inside
//       (0, 1), the slot contributes exactly alpha = RN(na / det).  Division is monotonic, so the
//       minimum over such slots is RN of the smallest FRACTION: fractions are compared exactly
//       (products of two floats as head + FMA tail) and only the winner is divided, once per ray;
//   anything else    (borderline, det == 0 / NaN / clipped): the literal arithmetic.
__device__ __forceinline__ void lidar_slot_near(float dx12, float dy12, float dx43, float dy43, float dx13,
                                                float dy13, float na, float det_raw,
                                                float& amin, float& best_num, float& best_den) {
  const float ad = fabsf(det_raw);
  const float nb = fadd(fmul(-dy12, dx13), fmul(dx12, dy13));
  const bool neg = det_raw < 0.f;
  const float an = neg ? -na : na, bn = neg ? -nb : nb;      // numerators over the denominator ad > 0
  const bool regular = (ad >= 1e-7f) && (ad <= 1e7f);        // false for 0 and NaN; sign * clip is the identity here
  const float hi = 1.0001f * ad, lo = 0.9999f * ad;
  const bool out = (an > hi) || (an < -1e-30f) || (bn > hi) || (bn < -1e-30f);
  const bool in = (an > 1e-30f) && (an < lo) && (bn > 1e-30f) && (bn < lo);
  if (regular && in) {
    // an / ad < best_num / best_den  <=>  an * best_den < best_num * ad   (all positive; exact)
    const float p1 = __fmul_rn(an, best_den), e1 = __fmaf_rn(an, best_den, -p1);
    const float p2 = __fmul_rn(best_num, ad), e2 = __fmaf_rn(best_num, ad, -p2);
    if (p1 < p2 || (p1 == p2 && e1 < e2)) { best_num = an; best_den = ad; }
  } else if (!(regular && out)) {
    amin = nanmin(amin, lidar_slot_literal(dx12, dy12, dx43, dy43, dx13, dy13, na, det_raw));
  }
}

__global__ void __launch_bounds__(K2_WARPS * 32)
lidar_kernel(EnvConsts k, const float* __restrict__ agent, const float* __restrict__ obstacles,
             const float* __restrict__ ray_dirs, float* __restrict__ hits, int b, int sd, int predict,
             int a_pitch, int h_pitch) {
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long item = (long)blockIdx.x * K2_WARPS + warp;
  const int n = k.n, R = k.n_rays, ne = k.n_obs * 4;
  if (item >= (long)b * n) return;
  // (b * n fits 32 bits whenever the grid does: the launcher's grid is an int)
  const int env = (int)((unsigned)item / (unsigned)n);
  const int per_warp = ne * 5 + ((ne + 3) & ~3) + 4 * R;        // floats (see the host launcher)
  float* base = smem + (size_t)warp * per_warp;
  float4* ed = reinterpret_cast<float4*>(base);                 // [ne] (dx43, dy43, dx13, dy13)
  float* nav = base + 4 * ne;                                   // [ne] numerator of alpha
  float* farf = nav + ne;                                       // [ne] spare (keeps the launcher's layout)
  unsigned long long* key = reinterpret_cast<unsigned long long*>(farf + ((ne + 3) & ~3));   // [R]
  float* hx = reinterpret_cast<float*>(key + R);                // [R]
  float* hy = hx + R;

  // slot-pointer addressing: env e of a (b, pitch, n, ...) record sits pitch * n rows after env e - 1
  const long a_item = (long)env * a_pitch * n + (item - (long)env * n);
  const long h_item = (long)env * h_pitch * n + (item - (long)env * n);
  float x1 = agent[a_item * sd + 0], y1 = agent[a_item * sd + 1];
  if (predict) {                      // agent holds the state BEFORE the step: cast from where it will be after it
    float s[5];
    for (int c = 0; c < sd; ++c) s[c] = agent[a_item * sd + c];
    next_position(k, s, sd, x1, y1);
  }
  const float* ob = obstacles + (size_t)env * k.n_obs * DGPPO_OBS_STRIDE;

  // inside_obstacles(start, r=0): obstacle.py:62-72 with r = 0 reduces to
  // (rel_xx < 0 && rel_yy < 0); the corner/circle clause needs sqrt(..) < 0.
  bool in_any = false;
  for (int o = lane; o < k.n_obs; o += 32) {
    const float4 a = *reinterpret_cast<const float4*>(ob + o * DGPPO_OBS_STRIDE);
    const float4 c = *reinterpret_cast<const float4*>(ob + o * DGPPO_OBS_STRIDE + 4);
    const float rel_x = fsub(x1, a.x), rel_y = fsub(y1, a.y);
    const float hw = fdiv(a.z, 2.f), hh = fdiv(a.w, 2.f);
    const float cs = c.y, sn = c.z;
    const float rxx = fsub(fabsf(fadd(fmul(rel_x, cs), fmul(rel_y, sn))), hw);
    const float ryy = fsub(fabsf(fsub(fmul(rel_x, sn), fmul(rel_y, cs))), hh);
    in_any |= (rxx < 0.f) && (ryy < 0.f);
  }
  in_any = __any_sync(0xffffffffu, in_any);
  const float keep = fsub(1.f, in_any ? 1.f : 0.f);

  // ---- phase 1: per-edge terms; near edges are packed to the front of the warp's list, far edges
  //      (beyond the sensing range) to the back, so phase 2 runs two branch-free loops
  const float reach = k.R * 1.002f + 1e-4f;            // a valid hit lies within comm_radius of the agent
  int n_near = 0, n_far = 0;
  for (int e0 = 0; e0 < ne; e0 += 32) {
    const int e = e0 + lane;
    const bool act = e < ne;
    float dx43 = 0.f, dy43 = 0.f, dx13 = 0.f, dy13 = 0.f;
    bool far = false;
    if (act) {
      const int o = e >> 2, q = e & 3;
      const float* pts = ob + o * DGPPO_OBS_STRIDE + 8;
      const float x3 = pts[2 * q], y3 = pts[2 * q + 1];
      const float x4 = pts[2 * ((q + 3) & 3)], y4 = pts[2 * ((q + 3) & 3) + 1];
      dx43 = fsub(x4, x3); dy43 = fsub(y4, y3);
      dx13 = fsub(x1, x3); dy13 = fsub(y1, y3);
      // distance agent -> segment (approximate arithmetic; only used with a safety margin)
      const float l2 = dx43 * dx43 + dy43 * dy43;
      float t = (l2 > 0.f) ? (dx13 * dx43 + dy13 * dy43) / l2 : 0.f;
      t = fminf(fmaxf(t, 0.f), 1.f);
      const float cx = dx13 - t * dx43, cy = dy13 - t * dy43;
      far = cx * cx + cy * cy > reach * reach;
    }
    const unsigned m_near = __ballot_sync(0xffffffffu, act && !far);
    const unsigned m_far = __ballot_sync(0xffffffffu, act && far);
    if (act) {
      const unsigned lt = (1u << lane) - 1u;
      const int p = far ? ne - 1 - (n_far + __popc(m_far & lt)) : n_near + __popc(m_near & lt);
      ed[p] = make_float4(dx43, dy43, dx13, dy13);
      nav[p] = fsub(fmul(dy43, dx13), fmul(dx43, dy13));
    }
    n_near += __popc(m_near); n_far += __popc(m_far);
  }
  __syncwarp();

  // ---- phase 2: per-ray minimum over the edges
  for (int r = lane; r < R; r += 32) {
    const float x2 = fadd(x1, ray_dirs[2 * r]), y2 = fadd(y1, ray_dirs[2 * r + 1]);
    const float dx12 = fsub(x1, x2), dy12 = fsub(y1, y2);
    float amin = 1e6f;                                  // every invalid slot contributes exactly 1e6
    float best_num = 1.f, best_den = 0.f;               // smallest clearly-valid alpha as a fraction (+inf)
    for (int e = 0; e < n_near; ++e) {
      const float4 g = ed[e];
      const float det = fsub(fmul(dx12, g.y), fmul(dy12, g.x));
      lidar_slot_near(dx12, dy12, g.x, g.y, g.z, g.w, nav[e], det, amin, best_num, best_den);
    }
    for (int e = n_near; e < ne; ++e) {
      // beyond reach: a regular solve cannot be valid; only a (near-)parallel ray, whose
      // determinant is clipped / zero, still has to go through the literal arithmetic
      const float2 g2 = *reinterpret_cast<const float2*>(&ed[e]);
      const float det = fsub(fmul(dx12, g2.y), fmul(dy12, g2.x));
      if (!(fabsf(det) >= 1e-7f)) {
        const float4 g = ed[e];
        amin = nanmin(amin, lidar_slot_literal(dx12, dy12, g.x, g.y, g.z, g.w, nav[e], det));
      }
    }
    if (best_den != 0.f) amin = nanmin(amin, fdiv(best_num, best_den));
    const float a = fmul(amin, keep);                      // env/utils.py:129
    // alpha is +0, positive or NaN: its bit pattern orders like the value; NaN sorts last;
    // the ray index in the low bits makes the order total and stable (jnp.argsort, env/utils.py:132)
    const unsigned bits = (a != a) ? 0xffffffffu : __float_as_uint(a);
    key[r] = ((unsigned long long)bits << 32) | (unsigned)r;
    hx[r] = fadd(x1, fmul(fsub(x2, x1), a));               // env/utils.py:134
    hy[r] = fadd(y1, fmul(fsub(y2, y1), a));
  }
  __syncwarp();
  // ---- phase 3: stable top-k
  float* out = hits + h_item * k.top_k * 2;
  for (int r = lane; r < R; r += 32) {
    const unsigned long long kr = key[r];
    int rank = 0;
    for (int j = 0; j < R; ++j) rank += (key[j] < kr) ? 1 : 0;
    if (rank < k.top_k) { out[2 * rank] = hx[r]; out[2 * rank + 1] = hy[r]; }
  }
}

// ------------------------------------------------------------------- K3
// One CTA per environment.  Agent / goal / obstacle-node tiles are staged in
// shared memory; every thread then produces whole output elements (a float4
// edge row + its receiver/sender, or one node scalar) so all stores are
// coalesced.
constexpr int K3_THREADS = 128;

template <int SD>          // state_dim is a template parameter: the index arithmetic divides by it
__global__ void __launch_bounds__(K3_THREADS)
build_graph_kernel(EnvConsts k, GraphDims d, const float* __restrict__ agent,
                   const float* __restrict__ goal, const float* __restrict__ obs_nodes,
                   float* __restrict__ nodes, float* __restrict__ edges, float* __restrict__ states,
                   int* __restrict__ receivers, int* __restrict__ senders, int* __restrict__ node_type,
                   int* __restrict__ n_node, int* __restrict__ n_edge, int pitch, int b) {
  extern __shared__ float smem[];
  pdl_launch_dependents();                       // the next policy forward may start staging its weights
  const int env = blockIdx.x;
  if (env >= b) return;
  constexpr int sd = SD, nd = SD + 3;
  const int n = d.n, g = d.g, N = d.N, E = d.E;
  const bool lid = is_lidar(k.kind), bic = (SD == 5);
  const int ow = lid ? 2 : 4;                    // floats per obstacle node
  float* sa = smem;                              // agent states   n*sd
  float* sg = sa + n * sd;                       // goal states    g*sd
  float* fa = sg + g * sd;                       // agent feats    n*4
  float* fg = fa + n * 4;                        // goal feats     g*4
  float* so = fg + g * 4;                        // obstacle nodes n_on*ow

  const float* ag = agent + (size_t)env * n * sd;
  const float* gl = goal + (size_t)env * g * sd;
  for (int i = threadIdx.x; i < n * sd; i += K3_THREADS) sa[i] = ag[i];
  for (int i = threadIdx.x; i < g * sd; i += K3_THREADS) sg[i] = gl[i];
  if (d.n_on > 0) {
    const float* on = obs_nodes + (size_t)env * d.n_on * ow;
    for (int i = threadIdx.x; i < d.n_on * ow; i += K3_THREADS) so[i] = on[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n + g; i += K3_THREADS) {   // state2feat
    const float* s = (i < n) ? sa + i * sd : sg + (i - n) * sd;
    float* f = (i < n) ? fa + i * 4 : fg + (i - n) * 4;
    f[0] = s[0]; f[1] = s[1];
    if (bic) { f[2] = fmul(s[4], s[2]); f[3] = fmul(s[4], s[3]); }   // lidar_bicycle_target.py:113-118
    else     { f[2] = s[2]; f[3] = s[3]; }
  }
  __syncthreads();

  const size_t slot = (size_t)env * pitch;
  // nodes (N, nd), states (N, sd), node_type (N): lidar_env/base.py:234-264.  One thread per node row:
  // the state row goes straight to global (a float4 when sd == 4), the feature row
  // [state | one-hot(obstacle, goal, agent)] is assembled in shared memory and copied out coalesced.
  float* nb = so + d.n_on * ow;                  // [N][nd]
  float* os_ = states + slot * N * sd;
  int* ot_ = node_type + slot * N;
  for (int row = threadIdx.x; row < N; row += K3_THREADS) {
    float st[SD];
    int type;
    if (row < n) {
      type = 0;
#pragma unroll
      for (int c = 0; c < SD; ++c) st[c] = sa[row * sd + c];
    } else if (row < n + g) {
      type = 1;
#pragma unroll
      for (int c = 0; c < SD; ++c) st[c] = sg[(row - n) * sd + c];
    } else if (row < N - 1) {
      type = 2;
      const int o = row - n - g;
#pragma unroll
      for (int c = 0; c < SD; ++c) st[c] = (c < ow) ? so[o * ow + c] : 0.f;
    } else {
      type = -1;                                  // pad node: state -1 (utils/graph.py:217), features 0
#pragma unroll
      for (int c = 0; c < SD; ++c) st[c] = -1.f;
    }
    if (SD == 4) {
      reinterpret_cast<float4*>(os_)[row] = make_float4(st[0], st[1], st[2], st[3]);
    } else {
#pragma unroll
      for (int c = 0; c < SD; ++c) os_[row * sd + c] = st[c];
    }
    ot_[row] = type;
    float* f = nb + row * nd;
#pragma unroll
    for (int c = 0; c < SD; ++c) f[c] = (type >= 0) ? st[c] : 0.f;
    f[sd] = (type == 2) ? 1.f : 0.f; f[sd + 1] = (type == 1) ? 1.f : 0.f; f[sd + 2] = (type == 0) ? 1.f : 0.f;
  }
  __syncthreads();
  float* on_ = nodes + slot * N * nd;
  for (int idx = threadIdx.x; idx < N * nd; idx += K3_THREADS) on_[idx] = nb[idx];
  if (threadIdx.x == 0) {
    if (n_node) n_node[slot] = N;
    if (n_edge) n_edge[slot] = E;
  }

  // edges: static slots, [a-a | a-goal | a-obs]; masked slot -> recv=send=pad
  float4* oe_ = reinterpret_cast<float4*>(edges + slot * E * 4);
  int* or_ = receivers + slot * E;
  int* osn_ = senders + slot * E;
  const int pad = N - 1, nn = n * n, nag = n * d.n_ag;
  for (int e = threadIdx.x; e < E; e += K3_THREADS) {
    float4 f; bool m; int rcv, snd;
    if (e < nn) {                                            // lidar_spread.py:59-67
      const int i = e / n, j = e - i * n;
      const float* a = fa + i * 4; const float* c = fa + j * 4;
      f = make_float4(fsub(a[0], c[0]), fsub(a[1], c[1]), fsub(a[2], c[2]), fsub(a[3], c[3]));
      float dist = norm2(fsub(sa[i * sd], sa[j * sd]), fsub(sa[i * sd + 1], sa[j * sd + 1]));
      dist = fadd(dist, (i == j) ? k.R_diag : 0.f);
      m = dist < k.R; rcv = i; snd = j;
    } else if (e < nn + nag) {                               // lidar_spread.py:69-76 | lidar_target.py:69-76
      const int r = e - nn;
      const bool paired = is_target(k.kind);                // one goal per agent (not: one goal NODE, as in Formation)
      const int i = paired ? r : r / g;
      const int q = paired ? r : r - i * g;
      const float* a = fa + i * 4; const float* c = fg + q * 4;
      f = make_float4(fsub(a[0], c[0]), fsub(a[1], c[1]), fsub(a[2], c[2]), fsub(a[3], c[3]));
      m = true; rcv = i; snd = n + q;
    } else {
      const int r = e - nn - nag;
      const int i = r / d.n_ao, q = r - i * d.n_ao;
      if (lid) {                                             // lidar_spread.py:78-94
        const float fx = fsub(sa[i * sd], so[r * 2]), fy = fsub(sa[i * sd + 1], so[r * 2 + 1]);
        f = make_float4(fx, fy, 0.f, 0.f);
        m = norm2(fx, fy) < k.R_obs; rcv = i; snd = n + g + r;
      } else {                                               // mpe_spread.py:70-79
        const float* a = sa + i * sd; const float* c = so + q * 4;
        f = make_float4(fsub(a[0], c[0]), fsub(a[1], c[1]), fsub(a[2], c[2]), fsub(a[3], c[3]));
        m = norm2(f.x, f.y) < k.R_mpe_obs; rcv = i; snd = n + g + q;
      }
    }
    oe_[e] = f;
    or_[e] = m ? rcv : pad;
    osn_[e] = m ? snd : pad;
  }
}

}  // namespace dgppo

using namespace dgppo;

extern "C" int dgppo_abi_version(void) { return DGPPO_ABI_VERSION; }

extern "C" int dgppo_n_goals(const DgppoEnvCfg* cfg) {
  if (!cfg || cfg->kind < 0 || cfg->kind > 9) return DGPPO_EINVAL;
  return n_goals_of(cfg->kind, cfg->n_agents);
}

extern "C" int dgppo_n_cost(const DgppoEnvCfg* cfg) {
  if (!cfg || cfg->kind < 0 || cfg->kind > 9) return DGPPO_EINVAL;
  return n_cost_of(cfg->kind);
}

extern "C" int dgppo_graph_dims(const DgppoEnvCfg* cfg, DgppoGraphDims* out) {
  if (int rc = check_env_cfg(cfg)) return rc;
  if (!out) return DGPPO_EINVAL;
  const GraphDims d = graph_dims(*cfg);
  out->state_dim = d.sd; out->node_dim = d.nd; out->edge_dim = 4;
  out->n_obs_nodes = d.n_on; out->n_nodes = d.N; out->n_edges = d.E;
  out->n_ag = d.n_ag; out->n_ao = d.n_ao;
  return 0;
}

extern "C" int dgppo_env_step(void* stream, const DgppoEnvCfg* cfg, const float* agent,
                              const float* goal, const float* obs_nodes, const float* action,
                              float* next_agent, float* reward, float* cost, int32_t io_pitch, int32_t b) {
  return dgppo::launch_env_step(stream, cfg, agent, goal, obs_nodes, action, next_agent, reward, cost, io_pitch, b, 1);
}

int dgppo::launch_env_step(void* stream, const DgppoEnvCfg* cfg, const float* agent, const float* goal,
                           const float* obs_nodes, const float* action, float* next_agent, float* reward,
                           float* cost, int32_t io_pitch, int32_t b, int32_t st_pitch) {
  if (int rc = check_env_cfg(cfg)) return rc;
  if (b == 0) return 0;
  if (b < 0 || io_pitch < 1 || !agent || !goal || !action || !next_agent || !reward || !cost) return DGPPO_EINVAL;
  if (cfg->n_obs > 0 && !obs_nodes) return DGPPO_EINVAL;
  const EnvConsts k = make_consts(*cfg);
  const size_t smem = (size_t)K1_WARPS * 5 * k.n * sizeof(float);
  if (smem > 48 * 1024) return DGPPO_ENOTSUP;
  const int grid = (b + K1_WARPS - 1) / K1_WARPS;
  // programmatic dependent launch after the head (saves the launch latency; measured at 512 envs: 5.99 -> 5.87 ms
  // per rollout).  K2 is launched normally: as a PDL node its look-ahead branch lost its overlap (6.8 ms).
  static const char* pe = getenv("DGPPO_PDL_STEP");
  launch_pdl(!(pe && pe[0] == '0'), env_step_kernel, grid, K1_WARPS * 32, smem, (cudaStream_t)stream,
             k, agent, goal, obs_nodes, action, next_agent, reward, cost, (int)io_pitch, (int)b, (int)st_pitch);
  return (int)cudaGetLastError();
}

extern "C" int dgppo_lidar(void* stream, const DgppoEnvCfg* cfg, const float* agent,
                           const float* obstacles, const float* ray_dirs, float* hits, int32_t b) {
  return dgppo::launch_lidar(stream, cfg, agent, obstacles, ray_dirs, hits, b, 0);
}

int dgppo::launch_lidar(void* stream, const DgppoEnvCfg* cfg, const float* agent, const float* obstacles,
                        const float* ray_dirs, float* hits, int32_t b, int predict, int32_t a_pitch, int32_t h_pitch) {
  if (int rc = check_env_cfg(cfg)) return rc;
  if (!is_lidar(cfg->kind) || cfg->n_obs == 0) return DGPPO_ENOTSUP;
  if (b == 0) return 0;
  if (b < 0 || !agent || !obstacles || !ray_dirs || !hits) return DGPPO_EINVAL;
  const EnvConsts k = make_consts(*cfg);
  const int ne = k.n_obs * 4;
  const size_t smem = (size_t)K2_WARPS * (ne * 5 + ((ne + 3) & ~3) + 4 * k.n_rays) * sizeof(float);
  if (smem > 48 * 1024) return DGPPO_ENOTSUP;
  const long items = (long)b * k.n;
  if (items > 0x7fffffffL) return DGPPO_ENOTSUP;
  const int grid = (int)((items + K2_WARPS - 1) / K2_WARPS);
  launch_pdl(false, lidar_kernel, grid, K2_WARPS * 32, smem, (cudaStream_t)stream,
             k, agent, obstacles, ray_dirs, hits, (int)b, is_bicycle(cfg->kind) ? 5 : 4, predict, (int)a_pitch, (int)h_pitch);
  return (int)cudaGetLastError();
}

extern "C" int dgppo_build_graph(void* stream, const DgppoEnvCfg* cfg, const float* agent,
                                 const float* goal, const float* obs_nodes, float* nodes,
                                 float* edges, float* states, int32_t* receivers, int32_t* senders,
                                 int32_t* node_type, int32_t* n_node, int32_t* n_edge,
                                 int32_t pitch, int32_t b) {
  if (int rc = check_env_cfg(cfg)) return rc;
  if (b == 0) return 0;
  if (b < 0 || pitch < 1 || !agent || !goal || !nodes || !edges || !states || !receivers ||
      !senders || !node_type) return DGPPO_EINVAL;
  const EnvConsts k = make_consts(*cfg);
  const GraphDims d = graph_dims(*cfg);
  if (d.n_on > 0 && !obs_nodes) return DGPPO_EINVAL;
  const int ow = is_lidar(cfg->kind) ? 2 : 4;
  const size_t smem = (size_t)(d.n * d.sd + d.g * d.sd + (d.n + d.g) * 4 + d.n_on * ow + d.N * d.nd) * sizeof(float);
  if (smem > 48 * 1024) return DGPPO_ENOTSUP;
  if (d.sd == 5)
    build_graph_kernel<5><<<b, K3_THREADS, smem, (cudaStream_t)stream>>>(
        k, d, agent, goal, obs_nodes, nodes, edges, states, receivers, senders, node_type, n_node, n_edge, pitch, b);
  else
    build_graph_kernel<4><<<b, K3_THREADS, smem, (cudaStream_t)stream>>>(
        k, d, agent, goal, obs_nodes, nodes, edges, states, receivers, senders, node_type, n_node, n_edge, pitch, b);
  return (int)cudaGetLastError();
}
