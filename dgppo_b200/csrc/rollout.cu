// Rollout orchestrator: the T-step scan of rollout / test_rollout
// (dgppo/trainer/utils.py:45-57,70-86) as one asynchronous launch sequence on
// the caller's stream: T x {K4a policy, K1 step, K2 LiDAR, K3 graph}, each
// writing straight into slot t / t+1 of the caller's (b, T+1, ...) record.
// No host synchronisation; per-step GPU time exceeds the launch cost at the
// benchmark sizes, so the host runs ahead of the device.
#include "common.cuh"

using namespace dgppo;

extern "C" int dgppo_rollout(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                             const float* params, const DgppoRolloutBuffers* B, int32_t T, int32_t b) {
  if (int rc = check_env_cfg(env)) return rc;
  if (!net || !params || !B || T < 1 || b < 0) return DGPPO_EINVAL;
  if (b == 0) return 0;
  if (!B->nodes || !B->edges || !B->states || !B->receivers || !B->senders || !B->node_type ||
      !B->rnn || !B->actions || !B->rewards || !B->costs || !B->agent_ws || !B->goal)
    return DGPPO_EINVAL;
  if (B->eps && !B->log_pis) return DGPPO_EINVAL;
  const GraphDims d = graph_dims(*env);
  const bool lid = is_lidar(env->kind);
  if (d.n_on > 0 && (!B->obstacles || (lid && (!B->hits_ws || !B->ray_dirs)))) return DGPPO_EINVAL;
  const int n = d.n, P = T + 1;
  const size_t agent_sz = (size_t)b * n * d.sd;
  for (int t = 0; t < T; ++t) {
    const float* agent_cur = B->agent_ws + (size_t)(t & 1) * agent_sz;
    float* agent_nxt = B->agent_ws + (size_t)((t + 1) & 1) * agent_sz;
    const float* obs_nodes = (d.n_on == 0) ? nullptr : (lid ? B->hits_ws : B->obstacles);
    int rc = dgppo_gnn_policy(stream, env, net, params,
                              B->nodes + (size_t)t * d.N * d.nd, B->edges + (size_t)t * d.E * 4,
                              B->receivers + (size_t)t * d.E, B->senders + (size_t)t * d.E, P,
                              B->rnn + (size_t)t * n * 64, B->rnn + (size_t)(t + 1) * n * 64, P,
                              B->eps ? B->eps + (size_t)t * n * 2 : nullptr, T,
                              B->actions + (size_t)t * n * 2,
                              B->log_pis ? B->log_pis + (size_t)t * n : nullptr, T, b);
    if (rc) return rc;
    rc = dgppo_env_step(stream, env, agent_cur, B->goal, obs_nodes, B->actions + (size_t)t * n * 2,
                        agent_nxt, B->rewards + t, B->costs + (size_t)t * n * 2, T, b);
    if (rc) return rc;
    if (lid && d.n_on > 0) {
      rc = dgppo_lidar(stream, env, agent_nxt, B->obstacles, B->ray_dirs, B->hits_ws, b);
      if (rc) return rc;
    }
    rc = dgppo_build_graph(stream, env, agent_nxt, B->goal, obs_nodes,
                           B->nodes + (size_t)(t + 1) * d.N * d.nd, B->edges + (size_t)(t + 1) * d.E * 4,
                           B->states + (size_t)(t + 1) * d.N * d.sd,
                           B->receivers + (size_t)(t + 1) * d.E, B->senders + (size_t)(t + 1) * d.E,
                           B->node_type + (size_t)(t + 1) * d.N,
                           B->n_node ? B->n_node + (t + 1) : nullptr,
                           B->n_edge ? B->n_edge + (t + 1) : nullptr, P, b);
    if (rc) return rc;
  }
  return 0;
}
