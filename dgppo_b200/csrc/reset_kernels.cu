// Device-side env.reset: obstacle sampling + the rejection sampler of
// get_node_goal_rng (dgppo/env/utils.py:139-244) for LidarEnv.reset
// (lidar_env/base.py:89-124), LidarBicycleTarget.reset
// (lidar_bicycle_target.py:60-90) and MPE.reset (mpe/base.py:81-127).
//
// One warp per environment: the sampler is sequential per env (agent by agent,
// candidate by candidate), the validity test of a candidate (distance to the
// nodes placed so far, inside-obstacle test) is lane-parallel.
//
// Random numbers: jax's threefry streams cannot be reproduced (and need not be:
// only the accept / reject rules are semantics).  Draw `c` of env key `k` is the
// counter-based hash splitmix64(k + c * golden) -> two 24-bit uniforms; the
// oracle (oracle/reset_np.py) uses the same function, so kernel and oracle place
// identical agents / goals / obstacles for identical keys.
#include "common.cuh"

namespace dgppo {

struct ResetConsts {
  int kind, n, n_obs;
  float area, area_y, goal_shift_y, min_dist, car, obs_r, len_lo, len_hi, th_lo, th_hi;
  float fixed_obs[4];          // MPECorridor: the two obstacle centres (mpe_corridor.py:53-54)
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}

struct Rng {
  unsigned long long key; unsigned ctr;
  __device__ __forceinline__ float2 next2() {          // two uniforms in [0, 1)
    const unsigned long long x = splitmix64(key + (unsigned long long)(ctr++) * 0x9E3779B97F4A7C15ULL);
    return make_float2((float)(x >> 40) * (1.f / 16777216.f), (float)((x >> 8) & 0xFFFFFFu) * (1.f / 16777216.f));
  }
};

// Rectangle.inside with margin r (obstacle.py:62-72), any obstacle; lane-parallel over obstacles
__device__ __forceinline__ bool inside_any(float px, float py, const float* rec, int n_obs, float r, int lane) {
  bool in = false;
  for (int o = lane; o < n_obs; o += 32) {
    const float* q = rec + o * DGPPO_OBS_STRIDE;
    const float rel_x = fsub(px, q[0]), rel_y = fsub(py, q[1]);
    const float cs = q[5], sn = q[6];
    const float xx = fsub(fabsf(fadd(fmul(rel_x, cs), fmul(rel_y, sn))), fdiv(q[2], 2.f));
    const float yy = fsub(fabsf(fsub(fmul(rel_x, sn), fmul(rel_y, cs))), fdiv(q[3], 2.f));
    const bool in_down = (xx < r) && (yy < 0.f), in_up = (xx < 0.f) && (yy < r);
    const bool corner = (xx > 0.f) && (yy > 0.f) && (fsqrt(fadd(fmul(xx, xx), fmul(yy, yy))) < r);
    in |= in_down || in_up || corner;
  }
  return __any_sync(0xffffffffu, in);
}

// min_j |p_j - cand| <= min_dist over ALL n slots (unplaced slots sit at the origin, as the
// reference's zero-initialised all_nodes do: env/utils.py:151-152,171)
__device__ __forceinline__ bool collides(float cx, float cy, const float* pts, int n, float min_dist, int lane) {
  bool hit = false;
  for (int j = lane; j < n; j += 32) hit |= norm2(fsub(pts[2 * j], cx), fsub(pts[2 * j + 1], cy)) <= min_dist;
  return __any_sync(0xffffffffu, hit);
}

constexpr int RESET_WARPS = 4;
constexpr int RESET_MAX_ITER = 1024;           // env/utils.py:150
// The reference restarts the whole placement for ever when an area is too crowded to hold the agents
// (env/utils.py:229-232: a hang).  A kernel must terminate: after this many restarts the last placement
// is kept and the environment is flagged through n_draws[env] < 0, which the host turns into an error.
constexpr int RESET_MAX_RESTARTS = 16;

__global__ void __launch_bounds__(RESET_WARPS * 32)
reset_kernel(ResetConsts k, const unsigned long long* __restrict__ keys, float* __restrict__ agent,
             float* __restrict__ goal, float* __restrict__ obst, int* __restrict__ n_draws, int b, int sd) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * RESET_WARPS + warp;
  if (env >= b) return;
  const int n = k.n;
  const bool lid = is_lidar(k.kind);
  float* st = smem + (size_t)warp * (4 * n + k.n_obs * DGPPO_OBS_STRIDE);   // agent xy [n][2]
  float* gl = st + 2 * n;                                                    // goal xy  [n][2]
  float* rec = gl + 2 * n;                                                   // obstacle records (Lidar)
  Rng rng{keys[env], 0u};

  // ---- obstacles (Lidar): centre, (width, height), theta -> Rectangle.create (obstacle.py:39-56)
  if (lid) {
    for (int o = 0; o < k.n_obs; ++o) {
      const float2 c = rng.next2(), l = rng.next2(), t = rng.next2();
      if (lane == 0) {
        float* q = rec + o * DGPPO_OBS_STRIDE;
        const float cx = fmul(c.x, k.area), cy = fmul(c.y, k.area);
        const float w = fadd(k.len_lo, fmul(l.x, fsub(k.len_hi, k.len_lo)));
        const float h = fadd(k.len_lo, fmul(l.y, fsub(k.len_hi, k.len_lo)));
        const float th = fadd(k.th_lo, fmul(t.x, fsub(k.th_hi, k.th_lo)));
        const float cs = cosf(th), sn = sinf(th), hw = fdiv(w, 2.f), hh = fdiv(h, 2.f);
        q[0] = cx; q[1] = cy; q[2] = w; q[3] = h; q[4] = th; q[5] = cs; q[6] = sn; q[7] = 0.f;
        const float bx[4] = {hw, -hw, -hw, hw}, by[4] = {hh, hh, -hh, -hh};
        for (int p = 0; p < 4; ++p) {
          q[8 + 2 * p] = fadd(fadd(fmul(cs, bx[p]), fmul(-sn, by[p])), cx);
          q[9 + 2 * p] = fadd(fadd(fmul(sn, bx[p]), fmul(cs, by[p])), cy);
        }
      }
    }
    __syncwarp();
  }
  const float* recp = lid ? rec : nullptr;
  const int n_rect = lid ? k.n_obs : 0;
  const float half = fdiv(k.min_dist, 2.f);

  // ---- get_node_goal_rng (env/utils.py:139-244)
  for (int j = lane; j < 2 * n; j += 32) { st[j] = 0.f; gl[j] = 0.f; }
  __syncwarp();
  int agent_id = 0, restarts = 0;
  while (agent_id < n) {
    // agent candidate: redraw while it collides / lies in an obstacle, at most max_iter times
    float2 u = rng.next2();
    float cx = fmul(u.x, k.area), cy = fmul(u.y, k.area_y);          // max_side = [side_length, side_length_y]
    int it_a = 0;
    while (it_a < RESET_MAX_ITER &&
           (collides(cx, cy, st, n, k.min_dist, lane) || (n_rect > 0 && inside_any(cx, cy, recp, n_rect, half, lane)))) {
      ++it_a; u = rng.next2(); cx = fmul(u.x, k.area); cy = fmul(u.y, k.area_y);
    }
    __syncwarp();
    if (lane == 0) { st[2 * agent_id] = cx; st[2 * agent_id + 1] = cy; }
    __syncwarp();
    u = rng.next2();
    float gx = fmul(u.x, k.area), gy = fmul(u.y, k.area_y);
    int it_g = 0;
    while (it_g < RESET_MAX_ITER &&
           (collides(gx, gy, gl, n, k.min_dist, lane) || (n_rect > 0 && inside_any(gx, gy, recp, n_rect, half, lane)) ||
            gx < 0.f || gy < 0.f || gx > k.area || gy > k.area)) {
      ++it_g; u = rng.next2(); gx = fmul(u.x, k.area); gy = fmul(u.y, k.area_y);
    }
    __syncwarp();
    if (lane == 0) { gl[2 * agent_id] = gx; gl[2 * agent_id + 1] = gy; }
    __syncwarp();
    ++agent_id;
    if (it_a >= RESET_MAX_ITER || it_g >= RESET_MAX_ITER) {      // no solution found: start over (utils.py:229-232)
      if (++restarts > RESET_MAX_RESTARTS) continue;            // infeasible: keep going with what we have
      agent_id = 0;
      for (int j = lane; j < 2 * n; j += 32) { st[j] = 0.f; gl[j] = 0.f; }
      __syncwarp();
    }
  }

  // ---- MPE obstacles (mpe/base.py:93-118): first candidate in [0, area]^2, redraws in [3r, area-3r]^2
  float* oo = obst + (size_t)env * k.n_obs * (lid ? DGPPO_OBS_STRIDE : 4);
  if (k.kind == DGPPO_ENV_MPE_CORRIDOR) {
    // goals move up by the corridor offset (mpe_corridor.py:50); the two obstacles are fixed (:53-54)
    for (int j = lane; j < n; j += 32) gl[2 * j + 1] = fadd(gl[2 * j + 1], k.goal_shift_y);
    if (lane < 2) {
      oo[4 * lane] = k.fixed_obs[2 * lane]; oo[4 * lane + 1] = k.fixed_obs[2 * lane + 1];
      oo[4 * lane + 2] = 0.f; oo[4 * lane + 3] = 0.f;
    }
    __syncwarp();
  } else if (!lid) {
    const float lo = fmul(k.car, 3.f), hi = fsub(k.area, fmul(k.car, 3.f));
    for (int o = 0; o < k.n_obs; ++o) {
      float2 u = rng.next2();
      float ox = fmul(u.x, k.area), oy = fmul(u.y, k.area);
      int guard = 0;
      while (guard < (1 << 20) &&
             (collides(ox, oy, st, n, fadd(k.car, k.obs_r), lane) ||
              collides(ox, oy, gl, n, fadd(fmul(k.car, 2.f), k.obs_r), lane) ||
              ox < lo || oy < lo || ox > hi || oy > hi)) {
        ++guard; u = rng.next2();
        ox = fadd(lo, fmul(u.x, fsub(hi, lo))); oy = fadd(lo, fmul(u.y, fsub(hi, lo)));
      }
      if (guard >= (1 << 20)) restarts = RESET_MAX_RESTARTS + 1;   // flag: no room for the obstacle
      if (lane == 0) { oo[4 * o] = ox; oo[4 * o + 1] = oy; oo[4 * o + 2] = 0.f; oo[4 * o + 3] = 0.f; }
    }
  } else {
    for (int j = lane; j < k.n_obs * DGPPO_OBS_STRIDE; j += 32) oo[j] = rec[j];
  }

  // ---- states: [x, y, 0, 0] or bicycle [x, y, cos th, sin th, 0] with th ~ U[0, 2 pi)
  float* ao = agent + (size_t)env * n * sd;
  float* go = goal + (size_t)env * n * sd;
  for (int i = 0; i < n; ++i) {
    float cs = 0.f, sn = 0.f;
    if (sd == 5) { const float2 u = rng.next2(); const float th = fmul(u.x, 6.283185307179586f); cs = cosf(th); sn = sinf(th); }
    if (lane == 0) {
      ao[i * sd] = st[2 * i]; ao[i * sd + 1] = st[2 * i + 1];
      go[i * sd] = gl[2 * i]; go[i * sd + 1] = gl[2 * i + 1];
      for (int c = 2; c < sd; ++c) { ao[i * sd + c] = 0.f; go[i * sd + c] = 0.f; }
      if (sd == 5) { ao[i * sd + 2] = cs; ao[i * sd + 3] = sn; }
    }
  }
  if (n_draws && lane == 0) n_draws[env] = (restarts > RESET_MAX_RESTARTS) ? -1 : (int)rng.ctr;
}


// ---------------------------------------------------------------------------------------------------
// The landmark families and the connected spread (env kinds 6-9): LidarLine.reset (lidar_line.py:38-126),
// MPELine.reset (mpe_line.py:38-117), MPEFormation.reset (mpe_formation.py:38-91), MPEConnectSpread.reset
// (mpe_connect_spread.py:52-103).  Same accept / reject rules, one warp per environment, the same counter
// stream as above.  Draw order (mirrored by oracle/reset_np.py:reset_landmark_states):
//   get_node_goal_rng(min_dist, no obstacles) [its goals are discarded except for the connected spread] ->
//   landmark 0 (+ the region draw) -> landmark 1 until far enough -> obstacles, each redrawn until valid.
// The quarter-turn of landmark 0 (rotation by region * pi / 2) is applied exactly (the reference multiplies by
// cos / sin of the fp32 angle, which differ from 0 / +-1 by < 1e-7).
struct LandmarkConsts {
  int kind, n, n_obs;
  float area, car, obs_r, connect_r, len_lo, len_hi, node_min_dist, lm_min_dist, side, side_y, goal_shift_y;
  float lm_lo, lm_hi;            // Formation: landmark range [R + 2 car, area - R - 2 car]
  const float* goal_table;       // Formation: (n, 2) offsets
};

// get_node_goal_rng without obstacles (env/utils.py:139-244) into st / gl; false when the placement was
// given up (more than RESET_MAX_RESTARTS restarts)
__device__ __forceinline__ bool sample_nodes(Rng& rng, float* st, float* gl, int n, float min_dist, float ax, float ay,
                                             float area, int lane) {
  for (int j = lane; j < 2 * n; j += 32) { st[j] = 0.f; gl[j] = 0.f; }
  __syncwarp();
  int agent_id = 0, restarts = 0;
  while (agent_id < n) {
    float2 u = rng.next2();
    float cx = fmul(u.x, ax), cy = fmul(u.y, ay);
    int it_a = 0;
    while (it_a < RESET_MAX_ITER && collides(cx, cy, st, n, min_dist, lane)) {
      ++it_a; u = rng.next2(); cx = fmul(u.x, ax); cy = fmul(u.y, ay);
    }
    __syncwarp();
    if (lane == 0) { st[2 * agent_id] = cx; st[2 * agent_id + 1] = cy; }
    __syncwarp();
    u = rng.next2();
    float gx = fmul(u.x, ax), gy = fmul(u.y, ay);
    int it_g = 0;
    while (it_g < RESET_MAX_ITER &&
           (collides(gx, gy, gl, n, min_dist, lane) || gx < 0.f || gy < 0.f || gx > area || gy > area)) {
      ++it_g; u = rng.next2(); gx = fmul(u.x, ax); gy = fmul(u.y, ay);
    }
    __syncwarp();
    if (lane == 0) { gl[2 * agent_id] = gx; gl[2 * agent_id + 1] = gy; }
    __syncwarp();
    ++agent_id;
    if (it_a >= RESET_MAX_ITER || it_g >= RESET_MAX_ITER) {
      if (++restarts > RESET_MAX_RESTARTS) continue;
      agent_id = 0;
      for (int j = lane; j < 2 * n; j += 32) { st[j] = 0.f; gl[j] = 0.f; }
      __syncwarp();
    }
  }
  return restarts <= RESET_MAX_RESTARTS;
}

// any agent whose nearest neighbour is farther than hi, or (lo > 0) closer than lo (mpe_connect_spread.py:54-68)
__device__ __forceinline__ bool badly_spaced(const float* p, int n, float lo, float hi, int lane) {
  bool bad = false;
  for (int i = lane; i < n; i += 32) {
    float mind = INFINITY;
    for (int j = 0; j < n; ++j) {
      float d = norm2(fsub(p[2 * i], p[2 * j]), fsub(p[2 * i + 1], p[2 * j + 1]));
      d = fadd(d, (i == j) ? 1e6f : 0.f);
      mind = fminf(mind, d);
    }
    bad |= (mind > hi) || (mind < lo);
  }
  return __any_sync(0xffffffffu, bad);
}

__global__ void __launch_bounds__(RESET_WARPS * 32)
reset_landmark_kernel(LandmarkConsts k, const unsigned long long* __restrict__ keys, float* __restrict__ agent,
                      float* __restrict__ goal, float* __restrict__ obst, int* __restrict__ n_draws, int b) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * RESET_WARPS + warp;
  if (env >= b) return;
  const int n = k.n;
  float* st = smem + (size_t)warp * (6 * n + 4 + DGPPO_OBS_STRIDE);   // agent xy
  float* gl = st + 2 * n;                                             // sampled goal xy
  float* eg = gl + 2 * n;                                             // the n goal positions (landmark2goal)
  float* lm = eg + 2 * n;                                             // landmarks (2, 2)
  float* rec = lm + 4;                                                // one rectangle candidate
  Rng rng{keys[env], 0u};
  bool ok = true;
  const int sd = 4;
  float* ao = agent + (size_t)env * n * sd;
  const int g_nodes = n_goals_of(k.kind, n);
  float* go = goal + (size_t)env * g_nodes * sd;

  if (k.kind == DGPPO_ENV_MPE_CONNECT_SPREAD) {
    // redraw the whole placement until agents and goals are connected and the agents apart (:52-90)
    int tries = 0;
    do {
      ok = sample_nodes(rng, st, gl, n, k.node_min_dist, k.area, k.side_y, k.area, lane);
      for (int j = lane; j < n; j += 32) gl[2 * j + 1] = fadd(gl[2 * j + 1], k.goal_shift_y);
      __syncwarp();
    } while (ok && ++tries < 4096 &&
             (badly_spaced(st, n, fmul(k.car, 2.f), k.connect_r, lane) || badly_spaced(gl, n, 0.f, k.connect_r, lane)));
    if (tries >= 4096) ok = false;
    const float2 u = rng.next2();                                    // the one large obstacle (:93-97)
    if (lane == 0) {
      float* oo = obst + (size_t)env * 4;
      oo[0] = fadd(fmul(u.x, fsub(fsub(k.area, k.obs_r), k.obs_r)), k.obs_r);
      oo[1] = fdiv(k.area, 2.f); oo[2] = 0.f; oo[3] = 0.f;
    }
    for (int i = lane; i < n; i += 32) {
      ao[i * sd] = st[2 * i]; ao[i * sd + 1] = st[2 * i + 1]; ao[i * sd + 2] = 0.f; ao[i * sd + 3] = 0.f;
      go[i * sd] = gl[2 * i]; go[i * sd + 1] = gl[2 * i + 1]; go[i * sd + 2] = 0.f; go[i * sd + 3] = 0.f;
    }
    if (n_draws && lane == 0) n_draws[env] = ok ? (int)rng.ctr : -1;
    return;
  }

  ok = sample_nodes(rng, st, gl, n, k.node_min_dist, k.area, k.area, k.area, lane);   // goals discarded
  // ---- landmarks
  if (k.kind == DGPPO_ENV_MPE_FORMATION) {                            // mpe_formation.py:48-53
    const float2 u = rng.next2();
    if (lane == 0) {
      lm[0] = fadd(fmul(u.x, fsub(k.lm_hi, k.lm_lo)), k.lm_lo);
      lm[1] = fadd(fmul(u.y, fsub(k.lm_hi, k.lm_lo)), k.lm_lo);
      lm[2] = 0.f; lm[3] = 0.f;
    }
  } else {
    const bool short_line = k.kind == DGPPO_ENV_MPE_LINE && n <= 3;
    float2 u = rng.next2();
    float l0x, l0y;
    if (short_line) {                                                 // mpe_line.py:54-55
      l0x = fmul(u.x, k.area); l0y = fmul(u.y, k.area);
    } else {                                                          // lidar_line.py:51-66, mpe_line.py:57-72
      float cx = fmul(u.x, fsub(k.area, k.side)), cy = fmul(u.y, k.side);
      const float half = fdiv(k.area, 2.f);
      cx = fadd(fsub(cx, half), 0.f);
      cy = fadd(fsub(cy, 0.f), fsub(half, k.side));
      const float2 r = rng.next2();
      const int region = min(3, (int)(r.x * 4.f));
      float rx, ry;                                                   // exact quarter turns
      if (region == 0) { rx = cx; ry = cy; } else if (region == 1) { rx = -cy; ry = cx; }
      else if (region == 2) { rx = -cx; ry = -cy; } else { rx = cy; ry = -cx; }
      l0x = fadd(rx, half); l0y = fadd(ry, half);
    }
    u = rng.next2();
    float l1x = fmul(u.x, k.area), l1y = fmul(u.y, k.area);
    int guard = 0;
    while (guard < (1 << 16) && norm2(fsub(l1x, l0x), fsub(l1y, l0y)) < k.lm_min_dist) {
      ++guard; u = rng.next2(); l1x = fmul(u.x, k.area); l1y = fmul(u.y, k.area);
    }
    if (guard >= (1 << 16)) ok = false;
    if (lane == 0) { lm[0] = l0x; lm[1] = l0y; lm[2] = l1x; lm[3] = l1y; }
  }
  __syncwarp();
  // ---- the n goal positions (landmark2goal), the same arithmetic as K1's reward
  for (int q = lane; q < n; q += 32) {
    if (k.kind == DGPPO_ENV_MPE_FORMATION) {
      eg[2 * q] = fadd(lm[0], __ldg(k.goal_table + 2 * q));
      eg[2 * q + 1] = fadd(lm[1], __ldg(k.goal_table + 2 * q + 1));
    } else {
      const bool short_line = k.kind == DGPPO_ENV_MPE_LINE && n <= 3;
      const float kq = (float)(short_line ? q + 1 : q), ni = (float)(short_line ? n + 1 : n - 1);
      eg[2 * q] = fadd(lm[0], fdiv(fmul(kq, fsub(lm[2], lm[0])), ni));
      eg[2 * q + 1] = fadd(lm[1], fdiv(fmul(kq, fsub(lm[3], lm[1])), ni));
    }
  }
  __syncwarp();
  // ---- obstacles
  if (k.kind == DGPPO_ENV_LIDAR_LINE) {                               // lidar_line.py:84-121: redraw (pos, length, theta)
    const float r = fmul(k.car, 1.1f);                               // until no agent / goal lies inside (margin 1.1 car)
    for (int o = 0; o < k.n_obs; ++o) {
      int guard = 0;
      bool bad;
      do {
        const float2 c = rng.next2(), l = rng.next2(), t = rng.next2();
        if (lane == 0) {
          const float cx = fmul(c.x, k.area), cy = fmul(c.y, k.area);
          const float w = fadd(k.len_lo, fmul(l.x, fsub(k.len_hi, k.len_lo)));
          const float h = fadd(k.len_lo, fmul(l.y, fsub(k.len_hi, k.len_lo)));
          const float th = fmul(t.x, 3.14159274101257324f);           // U[0, pi]
          const float cs = cosf(th), sn = sinf(th), hw = fdiv(w, 2.f), hh = fdiv(h, 2.f);
          rec[0] = cx; rec[1] = cy; rec[2] = w; rec[3] = h; rec[4] = th; rec[5] = cs; rec[6] = sn; rec[7] = 0.f;
          const float bx[4] = {hw, -hw, -hw, hw}, by[4] = {hh, hh, -hh, -hh};
          for (int p = 0; p < 4; ++p) {
            rec[8 + 2 * p] = fadd(fadd(fmul(cs, bx[p]), fmul(-sn, by[p])), cx);
            rec[9 + 2 * p] = fadd(fadd(fmul(sn, bx[p]), fmul(cs, by[p])), cy);
          }
        }
        __syncwarp();
        bad = false;
        for (int p = lane; p < 2 * n; p += 32) {                      // points = [agents; goals]
          const float* q = (p < n) ? st + 2 * p : eg + 2 * (p - n);
          const float rel_x = fsub(q[0], rec[0]), rel_y = fsub(q[1], rec[1]);
          const float xx = fsub(fabsf(fadd(fmul(rel_x, rec[5]), fmul(rel_y, rec[6]))), fdiv(rec[2], 2.f));
          const float yy = fsub(fabsf(fsub(fmul(rel_x, rec[6]), fmul(rel_y, rec[5]))), fdiv(rec[3], 2.f));
          const bool in_down = (xx < r) && (yy < 0.f), in_up = (xx < 0.f) && (yy < r);
          const bool corner = (xx > 0.f) && (yy > 0.f) && (fsqrt(fadd(fmul(xx, xx), fmul(yy, yy))) < r);
          bad |= in_down || in_up || corner;
        }
        bad = __any_sync(0xffffffffu, bad);
      } while (bad && ++guard < (1 << 16));
      if (guard >= (1 << 16)) ok = false;
      float* oo = obst + ((size_t)env * k.n_obs + o) * DGPPO_OBS_STRIDE;
      for (int j = lane; j < DGPPO_OBS_STRIDE; j += 32) oo[j] = rec[j];
      __syncwarp();
    }
  } else {                                                            // circles: mpe_line.py:87-108, mpe_formation.py:56-78
    const float lo = fmul(k.car, 3.f), hi = fsub(k.area, fmul(k.car, 3.f));
    float* oo = obst + (size_t)env * k.n_obs * 4;
    for (int o = 0; o < k.n_obs; ++o) {
      float2 u = rng.next2();
      float ox = fmul(u.x, k.area), oy = fmul(u.y, k.area);
      int guard = 0;
      while (guard < (1 << 20) &&
             (collides(ox, oy, st, n, fadd(k.car, k.obs_r), lane) ||
              collides(ox, oy, eg, n, fadd(fmul(k.car, 2.f), k.obs_r), lane) ||
              ox < lo || oy < lo || ox > hi || oy > hi)) {
        ++guard; u = rng.next2();
        ox = fadd(lo, fmul(u.x, fsub(hi, lo))); oy = fadd(lo, fmul(u.y, fsub(hi, lo)));
      }
      if (guard >= (1 << 20)) ok = false;
      if (lane == 0) { oo[4 * o] = ox; oo[4 * o + 1] = oy; oo[4 * o + 2] = 0.f; oo[4 * o + 3] = 0.f; }
    }
  }
  for (int i = lane; i < n; i += 32) {
    ao[i * sd] = st[2 * i]; ao[i * sd + 1] = st[2 * i + 1]; ao[i * sd + 2] = 0.f; ao[i * sd + 3] = 0.f;
  }
  for (int q = lane; q < g_nodes; q += 32) {
    go[q * sd] = lm[2 * q]; go[q * sd + 1] = lm[2 * q + 1]; go[q * sd + 2] = 0.f; go[q * sd + 3] = 0.f;
  }
  if (n_draws && lane == 0) n_draws[env] = ok ? (int)rng.ctr : -1;
}

}  // namespace dgppo

using namespace dgppo;

extern "C" int dgppo_reset(void* stream, const DgppoEnvCfg* cfg, const uint64_t* keys,
                           double obs_len_lo, double obs_len_hi, double theta_lo, double theta_hi,
                           float* agent, float* goal, float* obstacles, int32_t* n_draws, int32_t b) {
  if (int rc = check_env_cfg(cfg)) return rc;
  if (b == 0) return 0;
  if (b < 0 || !keys || !agent || !goal || (cfg->n_obs > 0 && !obstacles)) return DGPPO_EINVAL;
  if (cfg->kind >= DGPPO_ENV_LIDAR_LINE) {                  // landmark families + connected spread
    LandmarkConsts k;
    const double A = cfg->area_size, car = cfg->car_radius, ro = cfg->obs_radius;
    const int n = cfg->n_agents;
    k.kind = cfg->kind; k.n = n; k.n_obs = cfg->n_obs;
    k.area = (float)A; k.car = (float)car; k.obs_r = (float)ro; k.connect_r = (float)cfg->connect_radius;
    k.len_lo = (float)obs_len_lo; k.len_hi = (float)obs_len_hi;
    k.node_min_dist = (float)((cfg->kind == DGPPO_ENV_MPE_CONNECT_SPREAD ? 2.3 : 2.0) * car);
    const bool short_line = cfg->kind == DGPPO_ENV_MPE_LINE && n <= 3;
    const double lm_min = short_line ? n * 5 * car : (n - 2) * 6 * car;      // mpe_line.py:49-52, lidar_line.py:49
    k.lm_min_dist = (float)lm_min;
    k.side = (float)(A - lm_min);
    if (is_line(cfg->kind) && !short_line && A - lm_min < 0) return DGPPO_EINVAL;   // "area size is too small"
    k.side_y = (float)((A - ro * 2) / 2 - 1.5 * car);                           // mpe_connect_spread.py:79
    k.goal_shift_y = (float)(A - (A - ro * 2) / 2 + 1.5 * car);                 // :82-84
    k.lm_lo = (float)(cfg->comm_radius + 2 * car);                              // mpe_formation.py:50-52
    k.lm_hi = (float)(A - cfg->comm_radius - 2 * car);
    k.goal_table = cfg->goal_table;
    const size_t smem = (size_t)RESET_WARPS * (6 * n + 4 + DGPPO_OBS_STRIDE) * sizeof(float);
    if (smem > 48 * 1024) return DGPPO_ENOTSUP;
    const int grid = (b + RESET_WARPS - 1) / RESET_WARPS;
    reset_landmark_kernel<<<grid, RESET_WARPS * 32, smem, (cudaStream_t)stream>>>(
        k, (const unsigned long long*)keys, agent, goal, obstacles, n_draws, b);
    return (int)cudaGetLastError();
  }
  ResetConsts k;
  k.kind = cfg->kind; k.n = cfg->n_agents; k.n_obs = cfg->n_obs;
  k.area = (float)cfg->area_size;
  k.area_y = k.area; k.goal_shift_y = 0.f;
  k.fixed_obs[0] = k.fixed_obs[1] = k.fixed_obs[2] = k.fixed_obs[3] = 0.f;
  if (cfg->kind == DGPPO_ENV_MPE_CORRIDOR) {                 // mpe_corridor.py:39-54, Python doubles rounded once
    const double A = cfg->area_size, ro = cfg->obs_radius, rc = cfg->car_radius;
    k.area_y = (float)((A - ro * 2) / 2 - 1.5 * rc);
    k.goal_shift_y = (float)(A - (A - ro * 2) / 2 + 1.5 * rc);
    k.fixed_obs[0] = (float)ro; k.fixed_obs[1] = (float)(A / 2);
    k.fixed_obs[2] = (float)(A - ro); k.fixed_obs[3] = (float)(A / 2);
  }
  k.min_dist = (float)((is_lidar(cfg->kind) ? 2.2 : 2.0) * cfg->car_radius);   // lidar_env/base.py:111 | mpe/base.py:88
  k.car = (float)cfg->car_radius; k.obs_r = (float)cfg->obs_radius;
  k.len_lo = (float)obs_len_lo; k.len_hi = (float)obs_len_hi; k.th_lo = (float)theta_lo; k.th_hi = (float)theta_hi;
  const size_t smem = (size_t)RESET_WARPS * (4 * k.n + k.n_obs * DGPPO_OBS_STRIDE) * sizeof(float);
  if (smem > 48 * 1024) return DGPPO_ENOTSUP;
  const int grid = (b + RESET_WARPS - 1) / RESET_WARPS;
  reset_kernel<<<grid, RESET_WARPS * 32, smem, (cudaStream_t)stream>>>(
      k, (const unsigned long long*)keys, agent, goal, obstacles, n_draws, b, is_bicycle(cfg->kind) ? 5 : 4);
  return (int)cudaGetLastError();
}
