/*
 * dgppo_abi.h - C ABI of libdgppo_b200.so, the sm_100a kernel library behind
 * the DGPPO rollout hot path.
 *
 * The reference (syzhang092218-source/dgppo) is pure Python/JAX: it has no
 * FFI, plugin or operator interface.  Its boundary for this path is the
 * Python API (env.reset/step/get_graph, algo.act/step/collect/update); every
 * entry point below names the reference function(s) (file:line under
 * /root/reference) whose jitted body it replaces.  A JAX host binds these
 * through XLA-FFI custom calls; this repo's host mirror
 * (dgppo_b200/) binds them with ctypes on torch device buffers.
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer is DEVICE memory owned by
 *    the caller; no allocation, no retention past the call (the rollout
 *    orchestrator takes an explicit workspace);
 *  - every launch is asynchronous on `stream` (a cudaStream_t passed as
 *    void*), no host synchronisation; re-entrant, no mutable globals;
 *  - return value: 0 on success, a cudaError_t (>0) from the launch, or a
 *    negative DGPPO_E* code for arguments the kernels do not support;
 *  - all arrays are dense row-major fp32 / int32, batched over a leading
 *    environment axis `b` (the reference batches with jax.vmap:
 *    dgppo/algo/informarl.py:183-184);
 *  - graph arrays may live inside a (b, pitch, ...) time-major-per-env
 *    record (pitch = T+1 for a Rollout buffer, 1 for a plain batch): the
 *    pointer addresses slot t of env 0 and consecutive envs are
 *    `pitch * per-graph-size` elements apart.
 */
#ifndef DGPPO_ABI_H_
#define DGPPO_ABI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGPPO_ABI_VERSION 7

/* negative error codes (positive values are cudaError_t) */
#define DGPPO_EINVAL   (-1)   /* inconsistent sizes / null pointer            */
#define DGPPO_ENOTSUP  (-2)   /* configuration outside what the kernels cover */

/* env kinds (dgppo/env/__init__.py:9-23): the ones on BASELINE.json configs + MPETarget, MPECorridor */
#define DGPPO_ENV_LIDAR_SPREAD          0  /* lidar_env/lidar_spread.py          */
#define DGPPO_ENV_LIDAR_TARGET          1  /* lidar_env/lidar_target.py          */
#define DGPPO_ENV_LIDAR_BICYCLE_TARGET  2  /* lidar_env/lidar_bicycle_target.py  */
#define DGPPO_ENV_MPE_SPREAD            3  /* mpe/mpe_spread.py                  */
#define DGPPO_ENV_MPE_TARGET            4  /* mpe/mpe_target.py (same kernels, paired goals) */
#define DGPPO_ENV_MPE_CORRIDOR          5  /* mpe/mpe_corridor.py: MPESpread + 2 fixed obstacles, y <= 2 area, obstacle edges always on */
/* the landmark families: the goal NODES are landmarks (2 for Line, 1 for Formation), the n goal positions
 * the reward uses are derived from them (landmark2goal: lidar_line.py:128-133, mpe_line.py:119-128,
 * mpe_formation.py:93-96); spread-type agent-goal edges to every landmark */
#define DGPPO_ENV_LIDAR_LINE            6  /* lidar_env/lidar_line.py            */
#define DGPPO_ENV_MPE_LINE              7  /* mpe/mpe_line.py                    */
#define DGPPO_ENV_MPE_FORMATION         8  /* mpe/mpe_formation.py               */
#define DGPPO_ENV_MPE_CONNECT_SPREAD    9  /* mpe/mpe_connect_spread.py: third cost column (connectivity), y <= 2 area, obstacle edges always on */

/* Static environment description: the PARAMS dicts (lidar_spread.py:13-22,
 * mpe_spread.py:12-19) plus dt / num_agents (env/__init__.py:47-53). */
typedef struct DgppoEnvCfg {
  int32_t kind;         /* DGPPO_ENV_*                                   */
  int32_t n_agents;     /* n                                             */
  int32_t n_obs;        /* rectangles (Lidar*) or circles (MPE*)         */
  int32_t n_rays;       /* LiDAR beams per agent (default 32)            */
  int32_t top_k;        /* LiDAR returns kept per agent (default 8)      */
  int32_t reserved_;    /* keeps the doubles 8-byte aligned              */
  /* Python floats are doubles: thresholds such as comm_radius - 1e-1,
   * comm_radius + 1, 2 * car_radius are formed in double and only then
   * rounded to fp32, as the reference's weak-typed scalars are.          */
  double comm_radius;   /* 0.5 (area*10 under --full-observation)        */
  double car_radius;    /* 0.05                                          */
  double obs_radius;    /* 0.05 (MPE only)                               */
  double area_size;     /* 1.5                                           */
  double dt;            /* 0.03                                          */
  double dist2goal;     /* 0.01                                          */
  double connect_radius; /* 0.45 (MPEConnectSpread only)                 */
  /* MPEFormation only, else NULL: DEVICE table (n, 2) f32 of the goal offsets comm_radius * [cos, sin](theta_k),
   * theta = linspace(0, 2 pi, n + 1)[:-1] (mpe_formation.py:93-96) - data, like K2's ray table, so that no
   * device libm enters a goal position */
  const float* goal_table;
} DgppoEnvCfg;

/* goal nodes of an env kind: 2 (Line), 1 (Formation), else n_agents; cost columns: 3 (ConnectSpread), else 2.
 * Every `goal` array below is (b, n_goals, state_dim) and every `cost` array (..., n, n_cost).              */
int dgppo_n_goals(const DgppoEnvCfg* cfg);
int dgppo_n_cost(const DgppoEnvCfg* cfg);

/* Derived graph sizes (utils/graph.py:212-247 + each env's edge_blocks). */
typedef struct DgppoGraphDims {
  int32_t state_dim;    /* 4, or 5 for the bicycle                        */
  int32_t node_dim;     /* state_dim + 3                                  */
  int32_t edge_dim;     /* 4                                              */
  int32_t n_obs_nodes;  /* top_k*n hit nodes (Lidar) or n_obs (MPE)       */
  int32_t n_nodes;      /* n + n_goal + n_obs_nodes + 1 (pad)             */
  int32_t n_edges;      /* n*n + n*n_ag + n*n_ao                          */
  int32_t n_ag;         /* goal senders per agent: n (Spread) or 1        */
  int32_t n_ao;         /* obstacle senders per agent: top_k | n_obs | 0  */
} DgppoGraphDims;

int dgppo_abi_version(void);
int dgppo_graph_dims(const DgppoEnvCfg* cfg, DgppoGraphDims* out);

/* Obstacle record, 16 floats per rectangle (Rectangle, env/obstacle.py:30-56):
 *   [cx, cy, width, height, theta, cos(theta), sin(theta), 0,
 *    p0x, p0y, p1x, p1y, p2x, p2y, p3x, p3y]
 * cos/sin are carried as data (obstacle.py:65-66 re-evaluates them per call). */
#define DGPPO_OBS_STRIDE 16

/* ---- K0: reset (sampling part) ----------------------------------------------
 * The random part of LidarEnv.reset / LidarBicycleTarget.reset / MPE.reset
 * (lidar_env/base.py:89-119, lidar_bicycle_target.py:60-85, mpe/base.py:81-125):
 * obstacle sampling + Rectangle.create (obstacle.py:39-56) and the rejection
 * sampler get_node_goal_rng (env/utils.py:139-244: <= 1024 tries per agent / goal,
 * restart from agent 0 on failure, unplaced slots repel from the origin).  The
 * caller then runs dgppo_lidar + dgppo_build_graph on the result, as reset does.
 * MPECorridor (mpe_corridor.py:39-54): agents / goals sampled in [0, area] x [0, side_length_y],
 * goals then shifted past the corridor, the two obstacles at their fixed places (no draws).
 *   keys      (b) u64 : one key per environment; draw c of key k is
 *             splitmix64(k + c * 0x9E3779B97F4A7C15) -> two 24-bit uniforms (jax's
 *             threefry streams are not reproduced; the accept / reject rules are)
 *   obs_len_lo/hi, theta_lo/hi : PARAMS["obs_len_range"], obstacle angle range
 *             ([0, 2 pi) LidarEnv, [-pi, pi) bicycle); ignored for MPE
 *   agent, goal (b, n, state_dim) out; obstacles (b, n_obs, DGPPO_OBS_STRIDE) out
 *             (Lidar) or (b, n_obs, 4) out (MPE); n_draws (b) out, nullable:
 *             random draws consumed, or -1 when the area cannot hold the agents
 *             (more than 16 restarts of the placement; the reference loops for
 *             ever in that case, env/utils.py:229-232).                            */
int dgppo_reset(void* stream, const DgppoEnvCfg* cfg, const uint64_t* keys,
                double obs_len_lo, double obs_len_hi, double theta_lo, double theta_hi,
                float* agent, float* goal, float* obstacles, int32_t* n_draws, int32_t b);

/* ---- K1: dynamics + reward + cost ------------------------------------
 * Replaces the arithmetic of LidarEnv.step / MPE.step minus LiDAR and graph
 * build: clip_action + agent_step_euler + clip_state
 * (lidar_env/base.py:142-149,160-161; mpe/base.py:129-135,146-147;
 * lidar_bicycle_target.py:92-111; env/base.py:80-86), get_reward
 * (lidar_spread.py:35-52, lidar_target.py:35-52, mpe_spread.py:32-49) and
 * get_cost (lidar_env/base.py:180-207, mpe/base.py:164-191), the last two on
 * the PRE-step state.
 *   agent      (b, n, state_dim)      current agent states
 *   goal       (b, n, state_dim)
 *   obs_nodes  Lidar: (b, n, top_k, 2) hit points stored in the current graph
 *              MPE:   (b, n_obs, 4) obstacle states; NULL iff n_obs == 0
 *   action     (b, io_pitch, n, 2)    unclipped; slot pointer (io_pitch = T
 *                                     inside a Rollout record, 1 for a batch)
 *   next_agent (b, n, state_dim)  out
 *   reward     (b, io_pitch)      out, slot pointer
 *   cost       (b, io_pitch, n, 2) out, slot pointer                        */
int dgppo_env_step(void* stream, const DgppoEnvCfg* cfg,
                   const float* agent, const float* goal, const float* obs_nodes,
                   const float* action,
                   float* next_agent, float* reward, float* cost,
                   int32_t io_pitch, int32_t b);

/* ---- K2: LiDAR ----------------------------------------------------------
 * get_lidar_data -> get_lidar -> raytracing -> Rectangle.raytracing/inside
 * (lidar_env/base.py:126-140, env/utils.py:49-79,82-136, obstacle.py:62-105).
 *   agent     (b, n, state_dim)  positions are columns 0:2
 *   obstacles (b, n_obs, DGPPO_OBS_STRIDE)
 *   ray_dirs  (n_rays, 2)  end-point offsets (cos, sin)(theta_r) * comm_radius
 *   hits      (b, n, top_k, 2) out: the top_k smallest-alpha hit points in
 *             stable ascending-alpha order (jnp.argsort, env/utils.py:132)    */
int dgppo_lidar(void* stream, const DgppoEnvCfg* cfg,
                const float* agent, const float* obstacles, const float* ray_dirs,
                float* hits, int32_t b);

/* ---- K3: radius graph ----------------------------------------------------
 * get_graph + edge_blocks + EdgeBlock.make_edges + GetGraph.to_padded
 * (lidar_env/base.py:227-271, mpe/base.py:211-241, lidar_spread.py:57-96,
 * lidar_target.py:57-96, lidar_bicycle_target.py:113-118, mpe_spread.py:51-81,
 * mpe_target.py:51-80, mpe_corridor.py:69-98,
 * utils/graph.py:35-44,212-247).  Outputs are the GraphsTuple array fields
 * (utils/graph.py:61-86) for slot `t` of a (b, pitch, ...) record:
 *   nodes (N, node_dim) f32, edges (E, 4) f32, states (N, state_dim) f32,
 *   receivers / senders (E) i32, node_type (N) i32, n_node / n_edge () i32
 * (n_node / n_edge may be NULL).                                            */
int dgppo_build_graph(void* stream, const DgppoEnvCfg* cfg,
                      const float* agent, const float* goal, const float* obs_nodes,
                      float* nodes, float* edges, float* states,
                      int32_t* receivers, int32_t* senders, int32_t* node_type,
                      int32_t* n_node, int32_t* n_edge,
                      int32_t pitch, int32_t b);

/* ---- networks ------------------------------------------------------------ */
#define DGPPO_NET_POLICY 0   /* PPOPolicy / TanhNormal   (algo/module/policy.py:20-78,132-212) */
#define DGPPO_NET_VH     1   /* DecRStateFn, per-agent   (algo/module/value.py:47-79)          */
#define DGPPO_NET_VL     2   /* RStateFn, mean-pooled    (algo/module/value.py:15-44)          */

typedef struct DgppoNetCfg {
  int32_t kind;       /* DGPPO_NET_*                                          */
  int32_t node_dim;   /* 7 | 8                                                */
  int32_t edge_dim;   /* 4                                                    */
  int32_t n_layers;   /* GraphTransformer layers: 2 (policy, Vl) | 1 (Vh)     */
  int32_t n_out;      /* action_dim (2) | n_cost (2) | 1                      */
} DgppoNetCfg;

/* Offsets (in floats) of each block in the packed parameter buffer.
 * Fixed widths follow the reference: msg_dim 32, gnn_out_dim 64, H = 3 heads,
 * head MLP (64,64), GRU 64 (policy.py:149-169, informarl.py:102-112,
 * dgppo.py:83-95).  The buffer is a DEVICE layout, produced from the
 * reference's flax pytree by dgppo_b200.algo.params.pack_params (or any host
 * following this description).  flax kernels are (in, out) row-major; with
 * IN = layer input width, D = layer output width, HD = H*D and the
 * GraphTransformer denses named as flax auto-names them (gnn.py:86-98,110:
 * Dense_0 query on receivers, Dense_1 key on senders, Dense_2 value on
 * senders, Dense_3 edge without bias, Dense_4 update), layer l holds
 *   wqk  [IN+1][H*INP] query and key merged (both are linear in the receiver row):
 *                        score(e) = q_h . k_h = x_r^T (Wq_h Wk_h^T) x_s + ... so with
 *                        INP = round_up(IN + 1, 4) and column (h, c):
 *                        wqk[c'][h*INP+c] = sum_j Dense_0.kernel[c'][h*D+j] * Dense_1.kernel[c][h*D+j]  (c < IN)
 *                                         = sum_j Dense_0.kernel[c'][h*D+j] * Dense_1.bias[h*D+j]       (c == IN)
 *                        row IN holds the same with Dense_0.bias in place of Dense_0.kernel[c'] (products
 *                        formed in double at pack time); columns c > IN are 0
 *   wagg [H][INA][D]   value / edge, grouped per head, INA = IN + 1 + 4:
 *                        wagg[h][c][j] = Dense_2.kernel[c][h*D+j]      (c < IN)
 *                                      = Dense_2.bias[h*D+j]           (c == IN)
 *                                      = Dense_3.kernel[c-IN-1][h*D+j] (c > IN)
 *   wu   [IN][D]       = Dense_4.kernel             bu [D] = Dense_4.bias
 * and, after the head/tail blocks, for the large-n fallback kernel only:
 *   wq   [IN][HD]      = Dense_0.kernel             bq [HD] = Dense_0.bias
 *   wkt  [H][D][INP]   wkt[h][j][c] = Dense_1.kernel[c][h*D+j] (c < IN), Dense_1.bias[h*D+j] (c == IN), 0
 * (the kernels evaluate the attention in the algebraically regrouped form
 *  score = x_r^T (Wq Wk^T) x_s + ...,  agg = sum_h Wv_h (sum_e a_e x_s) + ...,
 *  DESIGN.md "GNN regrouping").
 * Head (mlp.py:14-30): d0w [64][64], d0b, ln0s, ln0b, d1w, d1b, ln1s, ln1b.
 * GRU (flax GRUCell): wi = [ir|iz|in].kernel (64,192), bi = their biases
 *   (192), wh = [hr|hz|hn].kernel (64,192), bhn = hn.bias (64).
 * Tail: policy (policy.py:66-70): ScaleHid is a Dense WITHOUT activation feeding two Denses, so the
 *   three are merged at pack time (products in double, rounded once):
 *   out_w [64][4] = ScaleHid.kernel @ [OutputDenseMean.kernel | OutputDenseStdTrans.kernel] (n_out = 2),
 *   out_b [4]     = ScaleHid.bias @ [Mean.kernel | StdTrans.kernel] + [Mean.bias | StdTrans.bias].
 *   value: out_w [64][4] = Dense_0.kernel zero-padded to 4 columns, out_b [4].
 * tc_head: the B operands of the tensor-core head (tcgen05.mma kind::tf32, 3xTF32 split): for each of
 *   W = d0w, d1w (K = 64, N = 64), wi, wh (K = 64, N = 192), in that order, a block
 *   [hi | lo] with hi = TF32(W) (round to nearest), lo = TF32(W - hi), each copy laid out as
 *   [K/4][N][4]: block[kc][n][j] = W[4 kc + j][n] (the no-swizzle K-major canonical smem layout, so the
 *   kernel stages it with plain bulk copies).  65536 floats, 16-byte aligned. */
typedef struct DgppoNetLayout {
  int32_t wqk[2], wagg[2], wu[2], bu[2];
  int32_t wq[2], bq[2], wkt[2];                 /* fallback-kernel blocks, at the tail */
  int32_t in_dim[2], out_dim[2];
  int32_t d0w, d0b, ln0s, ln0b, d1w, d1b, ln1s, ln1b;
  int32_t wi, bi, wh, bhn;
  int32_t out_w, out_b;
  int32_t tc_head;    /* tensor-core copy of the head / GRU matrices (see above), 65536 floats */
  int32_t total;      /* floats in the packed buffer */
} DgppoNetLayout;

int dgppo_net_layout(const DgppoNetCfg* net, DgppoNetLayout* out);

/* ---- K4a: policy forward + sample -------------------------------------
 * InforMARL.step / act (algo/informarl.py:230-252) ->
 * PPOPolicy.sample_action / get_action (policy.py:191-203) ->
 * TanhNormal (policy.py:61-74) -> PolicyNet (policy.py:25-33) ->
 * GraphTransformerGNN (nn/gnn.py:78-142), MLP (nn/mlp.py:14-30),
 * RNN/GRUCell (nn/rnn.py:14-30), TanhTransformedDistribution
 * (algo/module/distribution.py:10-46).
 *   params  packed per dgppo_net_layout
 *   nodes/edges/receivers/senders  graph slot (see pitch)
 *   rnn_in  (b, rnn_pitch, n, 64) slot pointer;  rnn_out likewise
 *   eps     (b, eps_pitch, n, 2) slot pointer of N(0,1) draws, or NULL for
 *           the deterministic mode tanh(mean)
 *   action  (b, act_pitch, n, 2) slot pointer out
 *   log_pi  (b, act_pitch, n)   slot pointer out (NULL allowed when eps NULL) */
int dgppo_gnn_policy(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                     const float* params,
                     const float* nodes, const float* edges,
                     const int32_t* receivers, const int32_t* senders, int32_t pitch,
                     const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                     const float* eps, int32_t eps_pitch,
                     float* action, float* log_pi, int32_t act_pitch, int32_t b);

/* ---- K4b: value forward -------------------------------------------------
 * DGPPO.get_Vh -> ValueNet.get_value -> DecRStateFn (dgppo.py:128-134,
 * value.py:47-79,155-157) for kind VH: the GRU carry is the policy's stored
 * rnn state (n rows per env), output (b, out_pitch, n, n_out).
 * Kind VL (RStateFn, value.py:15-44): mean over agents, one GRU row per env:
 * rnn_in/rnn_out (b, rnn_pitch, 64), output (b, out_pitch).
 * The whole (b, pitch) record is evaluated in one call when n_slots > 1:
 * slots 0..n_slots-1 of every env (Vh over all (b,T): dgppo.py:219-220).     */
int dgppo_gnn_value(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                    const float* params,
                    const float* nodes, const float* edges,
                    const int32_t* receivers, const int32_t* senders, int32_t pitch,
                    const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                    float* value, int32_t out_pitch, int32_t n_slots, int32_t b);

/* ---- Vl scan ---------------------------------------------------------------
 * InforMARL.scan_Vl + the final Vl (informarl.py:281-293, dgppo.py:204-216): the
 * centralised value over slots 0..n_slots-1 of every env, the GRU carry threaded
 * through the slots.  Only the GRU is recurrent: the GNN layers of ALL b * n_slots
 * graphs run as one launch, then the head / GRU / output runs slot by slot.
 *   carry  (b, carry_pitch, 64), carry_pitch >= n_slots + 1: slot 0 = initial carry
 *          in, slot t + 1 = carry after slot t out (slots 1.. are scratch in between)
 *   value  (b, out_pitch) out, slot t at [e * out_pitch + t]                        */
int dgppo_vl_scan(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                  const float* params,
                  const float* nodes, const float* edges,
                  const int32_t* receivers, const int32_t* senders, int32_t pitch,
                  float* carry, int32_t carry_pitch, float* value, int32_t out_pitch,
                  int32_t n_slots, int32_t b);

/* ---- graph from state: the same three forwards WITHOUT a stored graph -------------------------------
 * The compact Rollout record (dgppo/trainer/data.py:8-16) keeps K3's INPUTS per slot - agent
 * states and LiDAR hit points - instead of the GraphsTuple arrays (2.9 KB instead of 10.7 KB per env-step at
 * C3).  These entry points take that record: the GNN layers form node features, slot masks, senders and edge
 * features of every graph in their staging phase with build_graph's own arithmetic
 * (utils/graph.py:212-247 + each env's edge_blocks), so outputs equal the graph-record calls bit for bit.  */
typedef struct DgppoStateRecord {
  const float* agent;      /* (b, pitch, n, state_dim) slot pointer                                          */
  const float* obs_nodes;  /* Lidar: hit points (b, pitch, n, top_k, 2) slot pointer; MPE: obstacle states
                              (b, n_obs, 4), static per env; NULL iff n_obs == 0                              */
  const float* goal;       /* (b, n_goals, state_dim), static per env                                        */
} DgppoStateRecord;

int dgppo_gnn_policy_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                                const float* params, const DgppoStateRecord* state, int32_t pitch,
                                const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                                const float* eps, int32_t eps_pitch,
                                float* action, float* log_pi, int32_t act_pitch, int32_t b);
int dgppo_gnn_value_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                               const float* params, const DgppoStateRecord* state, int32_t pitch,
                               const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                               float* value, int32_t out_pitch, int32_t n_slots, int32_t b);
int dgppo_vl_scan_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                             const float* params, const DgppoStateRecord* state, int32_t pitch,
                             float* carry, int32_t carry_pitch, float* value, int32_t out_pitch,
                             int32_t n_slots, int32_t b);

/* ---- K5: GAE ---------------------------------------------------------------
 * compute_dec_ocp_gae (algo/utils.py:11-79; callers dgppo.py:232-237,268-273).
 *   hs (b,T,n,nh) costs, l (b,T) = -rewards, Vh (b,T+1,n,nh), Vl (b,T+1)
 *   -> Qh (b,T,n,nh), Ql (b,T)                                               */
int dgppo_gae(void* stream, const float* hs, const float* l, const float* Vh, const float* Vl,
              float gamma, float gae_lambda, float* Qh, float* Ql,
              int32_t b, int32_t T, int32_t n, int32_t nh);

/* ---- CBF residual + advantage merge (dgppo.py:239-259) --------------------
 *   Ql (b,T), Vl (b,T+1), Vh (b,T+1,n,nh)
 *   -> A (b,T,n) [= -(safe ? Al_hat : 0) - max_h Acbf * cbf_weight],
 *      cbf_deriv (b,T,n,nh), Acbf (b,T,n,nh) (either may be NULL),
 *      is_safe (b,T,n) u8 (may be NULL)                                      */
int dgppo_cbf_advantage(void* stream, const float* Ql, const float* Vl, const float* Vh,
                        float dt, float alpha, float cbf_eps, float cbf_weight,
                        float* A, float* cbf_deriv, float* Acbf, uint8_t* is_safe,
                        int32_t b, int32_t T, int32_t n, int32_t nh);

/* ---- fused rollout -----------------------------------------------------------
 * rollout / test_rollout scan (trainer/utils.py:45-57,70-86) behind
 * algo.collect / det_rollout_fn (informarl.py:177-186,254-256,
 * dgppo.py:108-117,141), from an already-reset batch: slot 0 of the graph
 * record and of `agent` hold the initial state.  Runs T x {policy, step,
 * LiDAR, graph} on `stream` with no host synchronisation.
 * The record is (b, T+1, ...) per graph field (graph = [:, :T], next_graph =
 * [:, 1:]); rnn (b, T+1, n, 64) (rollout emits [:, :T], test_rollout [:, 1:]);
 * eps (b, T, n, 2) or NULL (deterministic); actions (b,T,n,2), log_pis (b,T,n)
 * (NULL allowed iff eps NULL), rewards (b,T), costs (b,T,n,2).
 * agent_ws: workspace (2, b, n, state_dim) ping-pong agent states, slot 0
 * initialised by the caller; hits_ws (b, n, top_k, 2) (Lidar) initialised
 * with the hits of the initial state.  hits_ws2 (same shape, nullable): with a
 * second hit buffer the LiDAR of step t + 1 - which depends on the state before
 * step t only, the action never moves the position within a step - runs on an
 * internal side stream concurrently with the policy forward of step t.         */
typedef struct DgppoRolloutBuffers {
  float* nodes; float* edges; float* states;
  int32_t* receivers; int32_t* senders; int32_t* node_type;
  int32_t* n_node; int32_t* n_edge;
  float* rnn; const float* eps;
  float* actions; float* log_pis; float* rewards; float* costs;
  float* agent_ws; float* hits_ws;
  const float* goal; const float* obstacles; const float* ray_dirs;
  float* hits_ws2;
  /* COMPACT record: when agent_rec is non-NULL the graph fields above (nodes ... n_edge) and the
   * workspaces are not used (may be NULL); the rollout stores K3's inputs per slot instead -
   *   agent_rec (b, T+1, n, state_dim), slot 0 = the reset state (caller-initialised);
   *   hits_rec  (b, T+1, n, top_k, 2)   (Lidar with obstacles; slot 0 = the hits of the reset state)
   * - and the policy runs through dgppo_gnn_policy_from_state.  K3 is not launched at all: a GraphsTuple for any
   * slot is dgppo_build_graph of that slot's state (bit-identical to what the full record would hold).        */
  float* agent_rec; float* hits_rec;
} DgppoRolloutBuffers;

int dgppo_rollout(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                  const float* params, const DgppoRolloutBuffers* buf,
                  int32_t T, int32_t b, void* prof /* nullable */);

/* Captured rollout (the launch-free form of the scan, trainer/utils.py:45-57): the kernel sequence of
 * dgppo_rollout recorded once into a CUDA graph and replayed with ONE launch per rollout.  The graph bakes
 * in `params` and every pointer of `buf`: the caller keeps those buffers alive and refills them in place
 * (initial state into slot 0 / the workspaces, eps, goal, obstacles, packed parameters) before each launch.
 * create returns NULL on failure with the error in *rc_out (nullable); nodes = kernels in the graph.    */
void* dgppo_rollout_graph_create(const DgppoEnvCfg* env, const DgppoNetCfg* net, const float* params,
                                 const DgppoRolloutBuffers* buf, int32_t T, int32_t b, int32_t* rc_out);
int   dgppo_rollout_graph_launch(void* graph, void* stream);
int   dgppo_rollout_graph_nodes(void* graph);
void  dgppo_rollout_graph_destroy(void* graph);

/* Optional per-kernel timing of a rollout (measurement only).  `prof` records
 * cudaEvents on the rollout's stream around the 4 kernels of every step; it
 * never synchronises inside dgppo_rollout.  dgppo_prof_read waits for the last
 * event and returns, per kernel kind [policy, step, lidar, graph], the summed
 * (and optionally the maximum) device time in ms over the T steps of the most
 * recent rollout that used it.                                               */
void* dgppo_prof_create(int32_t T);
void  dgppo_prof_destroy(void* prof);
int   dgppo_prof_read(void* prof, float* ms_sum4, float* ms_max4 /* nullable */);

#ifdef __cplusplus
}
#endif
#endif  /* DGPPO_ABI_H_ */
