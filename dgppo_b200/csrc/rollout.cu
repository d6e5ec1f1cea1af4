// Rollout orchestrator: the T-step scan of rollout / test_rollout
// (dgppo/trainer/utils.py:45-57,70-86) as one asynchronous launch sequence on
// the caller's stream: T x {K4a policy, K1 step, K2 LiDAR, K3 graph}, each
// writing straight into slot t / t+1 of the caller's (b, T+1, ...) record.
// No host synchronisation; per-step GPU time exceeds the launch cost at the
// benchmark sizes, so the host runs ahead of the device.
#include "common.cuh"

#include <stdlib.h>

#include <vector>

using namespace dgppo;

// Optional per-kernel timing: cudaEvents recorded on the rollout's own stream
// around each of the 4 kernels of every step (no synchronisation until read).
struct DgppoProf {
  int T;
  std::vector<cudaEvent_t> ev;     // [T][5]: before policy, after policy, after step, after lidar, after graph
  bool recorded;
};

extern "C" void* dgppo_prof_create(int32_t T) {
  if (T < 1) return nullptr;
  DgppoProf* p = new DgppoProf{T, std::vector<cudaEvent_t>((size_t)T * 5), false};
  for (auto& e : p->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { delete p; return nullptr; }
  return p;
}

extern "C" void dgppo_prof_destroy(void* prof) {
  DgppoProf* p = (DgppoProf*)prof;
  if (!p) return;
  for (auto& e : p->ev) cudaEventDestroy(e);
  delete p;
}

extern "C" int dgppo_prof_read(void* prof, float* ms_sum4, float* ms_max4) {
  DgppoProf* p = (DgppoProf*)prof;
  if (!p || !ms_sum4 || !p->recorded) return DGPPO_EINVAL;
  cudaError_t err = cudaEventSynchronize(p->ev.back());
  if (err != cudaSuccess) return (int)err;
  for (int k = 0; k < 4; ++k) { ms_sum4[k] = 0.f; if (ms_max4) ms_max4[k] = 0.f; }
  for (int t = 0; t < p->T; ++t)
    for (int k = 0; k < 4; ++k) {
      float ms = 0.f;
      err = cudaEventElapsedTime(&ms, p->ev[(size_t)t * 5 + k], p->ev[(size_t)t * 5 + k + 1]);
      if (err != cudaSuccess) return (int)err;
      ms_sum4[k] += ms;
      if (ms_max4 && ms > ms_max4[k]) ms_max4[k] = ms;
    }
  return 0;
}

// Resources of the LiDAR look-ahead branch.  Outside a capture they are released as soon as the launches are
// enqueued (the runtime defers the release until the work has drained); inside a capture they must outlive
// cudaStreamEndCapture, so the capturing caller passes a holder and releases it afterwards.
struct AheadRes {
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev;
  void release() {
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    ev.clear();
    if (side) cudaStreamDestroy(side);
    side = nullptr;
  }
};

static int rollout_impl(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                        const float* params, const DgppoRolloutBuffers* B, int32_t T, int32_t b,
                        void* prof_, AheadRes* keep) {
  DgppoProf* prof = (DgppoProf*)prof_;
  if (prof && prof->T != T) return DGPPO_EINVAL;
  auto mark = [&](int t, int k) {
    if (prof) cudaEventRecord(prof->ev[(size_t)t * 5 + k], (cudaStream_t)stream);
  };
  if (int rc = check_env_cfg(env)) return rc;
  if (!net || !params || !B || T < 1 || b < 0) return DGPPO_EINVAL;
  if (b == 0) return 0;
  if (!B->rnn || !B->actions || !B->rewards || !B->costs || !B->goal) return DGPPO_EINVAL;
  if (B->eps && !B->log_pis) return DGPPO_EINVAL;
  const GraphDims d = graph_dims(*env);
  const bool lid = is_lidar(env->kind);
  const int n = d.n, P = T + 1;
  if (B->agent_rec) {
    // ---- compact record: K3's inputs per slot instead of the graph.  Per step: policy from the
    // state of slot t (the GNN layers build the graph in their staging phase), K1 slot t -> t + 1, K2 slot t + 1.
    if (d.n_on > 0 && (!B->obstacles || (lid && (!B->hits_rec || !B->ray_dirs)))) return DGPPO_EINVAL;
    const int nh = n_cost_of(env->kind), k2 = env->top_k * 2;
    // LiDAR look-ahead as a parallel graph branch (captured rollouts of small batches, as below): the hits of slot
    // t + 1 depend on the state of slot t only, so they are cast from it (predict = 1) while the policy of step t runs.
    static const char* ahead_env_c = getenv("DGPPO_LIDAR_AHEAD");
    const bool on_c = ahead_env_c && ahead_env_c[0] == '1', off_c = ahead_env_c && ahead_env_c[0] == '0';
    const bool ahead_c = lid && d.n_on > 0 && !prof && (keep ? (on_c || (!off_c && b < 1024)) : false);
    AheadRes unused_c;
    AheadRes& res_c = keep ? *keep : unused_c;      // only used when ahead_c (keep != nullptr then)
    if (ahead_c) {
      if (cudaStreamCreateWithFlags(&res_c.side, cudaStreamNonBlocking) != cudaSuccess) return DGPPO_EINVAL;
      res_c.ev.assign((size_t)2 * (T + 1), nullptr);   // [2 t] state of slot t ready (main), [2 t + 1] hits of slot t ready (side)
      for (auto& e : res_c.ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return DGPPO_EINVAL;
      cudaEventRecord(res_c.ev[0], (cudaStream_t)stream);
      cudaStreamWaitEvent(res_c.side, res_c.ev[0], 0);
      if (int rc = launch_lidar(res_c.side, env, B->agent_rec, B->obstacles, B->ray_dirs, B->hits_rec + (size_t)n * k2,
                                b, 1, P, P)) return rc;
      cudaEventRecord(res_c.ev[3], res_c.side);
    }
    for (int t = 0; t < T; ++t) {
      mark(t, 0);
      if (ahead_c && t >= 1) cudaStreamWaitEvent((cudaStream_t)stream, res_c.ev[2 * t + 1], 0);   // hits of slot t
      DgppoStateRecord st;
      st.agent = B->agent_rec + (size_t)t * n * d.sd;
      st.obs_nodes = (d.n_on == 0) ? nullptr : (lid ? B->hits_rec + (size_t)t * n * k2 : B->obstacles);
      st.goal = B->goal;
      int rc = dgppo_gnn_policy_from_state(stream, env, net, params, &st, P,
                                           B->rnn + (size_t)t * n * 64, B->rnn + (size_t)(t + 1) * n * 64, P,
                                           B->eps ? B->eps + (size_t)t * n * 2 : nullptr, T,
                                           B->actions + (size_t)t * n * 2,
                                           B->log_pis ? B->log_pis + (size_t)t * n : nullptr, T, b);
      if (rc) return rc;
      mark(t, 1);
      rc = launch_env_step(stream, env, st.agent, B->goal, st.obs_nodes, B->actions + (size_t)t * n * 2,
                           B->agent_rec + (size_t)(t + 1) * n * d.sd, B->rewards + t,
                           B->costs + (size_t)t * n * nh, T, b, P);
      if (rc) return rc;
      mark(t, 2);
      if (ahead_c) {
        if (t + 2 <= T) {        // state of slot t + 1 exists: cast the hits of slot t + 2 from it on the side branch
          cudaEventRecord(res_c.ev[2 * (t + 1)], (cudaStream_t)stream);
          cudaStreamWaitEvent(res_c.side, res_c.ev[2 * (t + 1)], 0);
          rc = launch_lidar(res_c.side, env, B->agent_rec + (size_t)(t + 1) * n * d.sd, B->obstacles, B->ray_dirs,
                            B->hits_rec + (size_t)(t + 2) * n * k2, b, 1, P, P);
          if (rc) return rc;
          cudaEventRecord(res_c.ev[2 * (t + 2) + 1], res_c.side);
        }
      } else if (lid && d.n_on > 0) {
        rc = launch_lidar(stream, env, B->agent_rec + (size_t)(t + 1) * n * d.sd, B->obstacles, B->ray_dirs,
                          B->hits_rec + (size_t)(t + 1) * n * k2, b, 0, P, P);
        if (rc) return rc;
      }
      mark(t, 3);
      mark(t, 4);                                // no graph kernel in the loop
    }
    if (ahead_c) cudaStreamWaitEvent((cudaStream_t)stream, res_c.ev[2 * T + 1], 0);   // join: hits of the last slot
    if (prof) prof->recorded = true;
    return 0;
  }
  if (!B->nodes || !B->edges || !B->states || !B->receivers || !B->senders || !B->node_type || !B->agent_ws)
    return DGPPO_EINVAL;
  if (d.n_on > 0 && (!B->obstacles || (lid && (!B->hits_ws || !B->ray_dirs)))) return DGPPO_EINVAL;
  const size_t agent_sz = (size_t)b * n * d.sd;
  cudaStream_t main_st = (cudaStream_t)stream;

  // LiDAR look-ahead: the hits of graph t + 1 depend on the state BEFORE step t only (the action changes
  // velocity / heading, not the position, within a step), so K2 runs on a side stream while the policy of
  // step t is still in flight.  Needs the second hit buffer (graph t reads one while t + 1 is produced);
  // Opt-in (DGPPO_LIDAR_AHEAD=1): measured -1.9 % with one rollout stream, but +-0 with the default four
  // env groups (their streams already overlap K2 with other groups' policy kernels) while the cross-stream
  // events triple the host-side submission cost; never under the per-kernel profiler (serial schedule).
  // Inside a captured rollout (keep != nullptr) the branch costs the host nothing - it is a parallel branch of
  // the graph - and pays when the kernels are too small to fill the GPU (measured, C3, B200: 512 envs 7.37 ->
  // 6.96 ms, 4 x 256 envs 9.91 -> 9.21 ms; 4 x 1024 envs 28.6 -> 29.3 ms): default on below 1024 envs per
  // group; DGPPO_LIDAR_AHEAD=0 / 1 forces it.  Launched directly (no graph) it stays opt-in (=1).
  static const char* ahead_env = getenv("DGPPO_LIDAR_AHEAD");          // read once per process
  const bool forced_on = ahead_env && ahead_env[0] == '1', forced_off = ahead_env && ahead_env[0] == '0';
  const bool want_ahead = keep ? (forced_on || (!forced_off && b < 1024)) : forced_on;
  const bool ahead = lid && d.n_on > 0 && B->hits_ws2 && !prof && want_ahead;
  AheadRes local_res;
  AheadRes& res = keep ? *keep : local_res;
  cudaStream_t& side_st = res.side;
  std::vector<cudaEvent_t>& ev = res.ev;       // [2 t] state t + 1 ready (main), [2 t + 1] hits t + 2 ready (side)
  auto fail = [&](int rc) {
    if (!keep) res.release();
    return rc;
  };
  if (ahead) {
    if (cudaStreamCreateWithFlags(&side_st, cudaStreamNonBlocking) != cudaSuccess) return DGPPO_EINVAL;
    ev.assign((size_t)2 * T + 2, nullptr);
    for (auto& e : ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(DGPPO_EINVAL);
    // hits of graph 1, from state 0 (already in agent_ws[0] when the caller's stream gets here)
    cudaEventRecord(ev[2 * T], main_st);
    cudaStreamWaitEvent(side_st, ev[2 * T], 0);
    if (int rc = launch_lidar(side_st, env, B->agent_ws, B->obstacles, B->ray_dirs, B->hits_ws2, b, 1)) return fail(rc);
    cudaEventRecord(ev[2 * T + 1], side_st);
  }
  float* hits_buf[2] = {B->hits_ws, ahead ? B->hits_ws2 : B->hits_ws};   // graph t reads hits_buf[t & 1]

  for (int t = 0; t < T; ++t) {
    const float* agent_cur = B->agent_ws + (size_t)(t & 1) * agent_sz;
    float* agent_nxt = B->agent_ws + (size_t)((t + 1) & 1) * agent_sz;
    const float* hits_cur = hits_buf[t & 1];
    float* hits_nxt = hits_buf[(t + 1) & 1];
    const float* obs_cur = (d.n_on == 0) ? nullptr : (lid ? hits_cur : B->obstacles);
    const float* obs_nxt = (d.n_on == 0) ? nullptr : (lid ? hits_nxt : B->obstacles);
    mark(t, 0);
    int rc = dgppo_gnn_policy(stream, env, net, params,
                              B->nodes + (size_t)t * d.N * d.nd, B->edges + (size_t)t * d.E * 4,
                              B->receivers + (size_t)t * d.E, B->senders + (size_t)t * d.E, P,
                              B->rnn + (size_t)t * n * 64, B->rnn + (size_t)(t + 1) * n * 64, P,
                              B->eps ? B->eps + (size_t)t * n * 2 : nullptr, T,
                              B->actions + (size_t)t * n * 2,
                              B->log_pis ? B->log_pis + (size_t)t * n : nullptr, T, b);
    if (rc) return fail(rc);
    mark(t, 1);
    rc = dgppo_env_step(stream, env, agent_cur, B->goal, obs_cur, B->actions + (size_t)t * n * 2,
                        agent_nxt, B->rewards + t, B->costs + (size_t)t * n * n_cost_of(env->kind), T, b);
    if (rc) return fail(rc);
    mark(t, 2);
    if (ahead) {
      // state t + 1 exists: the side stream may cast the rays of graph t + 2 into the buffer graph t used
      // (its last reader, K1 of step t, is the kernel just enqueued)
      if (t + 1 < T) {
        cudaEventRecord(ev[2 * t], main_st);
        cudaStreamWaitEvent(side_st, ev[2 * t], 0);
        rc = launch_lidar(side_st, env, agent_nxt, B->obstacles, B->ray_dirs, hits_buf[t & 1], b, 1);
        if (rc) return fail(rc);
        cudaEventRecord(ev[2 * t + 1], side_st);
      }
      // the hits of graph t + 1 were produced one step ago (or before the loop)
      cudaStreamWaitEvent(main_st, t == 0 ? ev[2 * T + 1] : ev[2 * (t - 1) + 1], 0);
    } else if (lid && d.n_on > 0) {
      rc = dgppo_lidar(stream, env, agent_nxt, B->obstacles, B->ray_dirs, hits_nxt, b);
      if (rc) return fail(rc);
    }
    mark(t, 3);
    rc = dgppo_build_graph(stream, env, agent_nxt, B->goal, obs_nxt,
                           B->nodes + (size_t)(t + 1) * d.N * d.nd, B->edges + (size_t)(t + 1) * d.E * 4,
                           B->states + (size_t)(t + 1) * d.N * d.sd,
                           B->receivers + (size_t)(t + 1) * d.E, B->senders + (size_t)(t + 1) * d.E,
                           B->node_type + (size_t)(t + 1) * d.N,
                           B->n_node ? B->n_node + (t + 1) : nullptr,
                           B->n_edge ? B->n_edge + (t + 1) : nullptr, P, b);
    if (rc) return fail(rc);
    mark(t, 4);
  }
  fail(0);                                     // events / side stream: released once their work has drained
  if (prof) prof->recorded = true;
  return 0;
}

extern "C" int dgppo_rollout(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                             const float* params, const DgppoRolloutBuffers* B, int32_t T, int32_t b,
                             void* prof_) {
  return rollout_impl(stream, env, net, params, B, T, b, prof_, nullptr);
}

// ---- captured rollout ---------------------------------------------------------------------------
// The launch sequence above, recorded ONCE into a CUDA graph and replayed with one launch per rollout:
// T x (2 policy + step + LiDAR + graph) kernels cost the host 3-4 us each when launched one by one, which
// is what bounds the rollout once the batch per GPU is small (512 envs per GPU in the 8-GPU configuration).
// The graph bakes in every pointer of `buf` and `params`: the caller keeps those buffers alive and
// refills them in place (dgppo_b200/trainer/rollout.py keeps them in the record).
struct DgppoRolloutGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;
};

extern "C" void* dgppo_rollout_graph_create(const DgppoEnvCfg* env, const DgppoNetCfg* net, const float* params,
                                            const DgppoRolloutBuffers* B, int32_t T, int32_t b, int32_t* rc_out) {
  auto set_rc = [&](int rc) { if (rc_out) *rc_out = rc; };
  set_rc(DGPPO_EINVAL);
  if (b < 1) return nullptr;
  cudaStream_t cs = nullptr;
  if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  DgppoRolloutGraph* G = new DgppoRolloutGraph();
  cudaError_t err = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
  AheadRes res;
  int rc = (err == cudaSuccess) ? rollout_impl(cs, env, net, params, B, T, b, nullptr, &res) : (int)err;
  cudaGraph_t g = nullptr;
  if (err == cudaSuccess) {
    const cudaError_t e2 = cudaStreamEndCapture(cs, &g);      // always end the capture, also after a failed launch
    if (rc == 0 && e2 != cudaSuccess) rc = (int)e2;
  }
  if (rc == 0) {
    G->graph = g;
    size_t nn = 0;
    cudaGraphGetNodes(g, nullptr, &nn);
    G->n_kernels = (int)nn;
    err = cudaGraphInstantiate(&G->exec, g, 0);
    if (err != cudaSuccess) rc = (int)err;
  }
  res.release();
  cudaStreamDestroy(cs);
  set_rc(rc);
  if (rc != 0) {
    if (g) cudaGraphDestroy(g);
    delete G;
    cudaGetLastError();
    return nullptr;
  }
  return G;
}

extern "C" int dgppo_rollout_graph_launch(void* graph, void* stream) {
  DgppoRolloutGraph* G = (DgppoRolloutGraph*)graph;
  if (!G || !G->exec) return DGPPO_EINVAL;
  return (int)cudaGraphLaunch(G->exec, (cudaStream_t)stream);
}

extern "C" int dgppo_rollout_graph_nodes(void* graph) {
  DgppoRolloutGraph* G = (DgppoRolloutGraph*)graph;
  return G ? G->n_kernels : DGPPO_EINVAL;
}

extern "C" void dgppo_rollout_graph_destroy(void* graph) {
  DgppoRolloutGraph* G = (DgppoRolloutGraph*)graph;
  if (!G) return;
  if (G->exec) cudaGraphExecDestroy(G->exec);
  if (G->graph) cudaGraphDestroy(G->graph);
  delete G;
}
