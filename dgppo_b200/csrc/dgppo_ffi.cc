// XLA-FFI custom-call handlers over the C ABI of libdgppo_b200.so (include/dgppo_abi.h): what a JAX host
// (the reference) binds with jax.ffi.register_ffi_target / jax.ffi.ffi_call.
//
// Built only where jaxlib's headers exist (this image has no jax: the file is then an empty translation unit and
// dgppo_b200/csrc/build_ffi.sh says so and exits 0):
//   g++ -std=c++17 -shared -fPIC -I"$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')" \
//       -I/usr/local/cuda/include -Iinclude dgppo_b200/csrc/dgppo_ffi.cc -Ldgppo_b200 -ldgppo_b200 \
//       -Wl,-rpath,'$ORIGIN' -o dgppo_b200/libdgppo_ffi.so
//
// Conventions: buffers are the batched arrays of the header, leading axis = environments (jax.vmap with
// vmap_method="broadcast_all" hands the handler the whole batch); the static env / net description travels as
// scalar attributes; results are FFI result buffers XLA allocates.  Ten handlers, one per compute entry point:
//   DgppoReset, DgppoEnvStep, DgppoLidar, DgppoBuildGraph, DgppoGnnPolicy, DgppoGnnValue, DgppoVlScan, DgppoGae,
//   DgppoCbfAdvantage, DgppoRollout.
// [3P-unverified]: written against the documented xla::ffi API (jaxlib >= 0.4.31); never compiled here.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define DGPPO_HAVE_XLA_FFI 1
#endif
#endif

#ifdef DGPPO_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <string>

#include "dgppo_abi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using U64 = ffi::Buffer<ffi::U64>;
using U8 = ffi::Buffer<ffi::U8>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

namespace {

ffi::Error status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  if (rc == DGPPO_EINVAL) return ffi::Error::InvalidArgument(std::string(what) + ": DGPPO_EINVAL");
  if (rc == DGPPO_ENOTSUP) return ffi::Error(ffi::ErrorCode::kUnimplemented, std::string(what) + ": DGPPO_ENOTSUP");
  return ffi::Error::Internal(std::string(what) + ": CUDA error " + std::to_string(rc));
}

// the env description as attributes (every handler takes the same set, bound by ENV_ATTRS below)
struct EnvAttrs {
  int32_t kind, n_agents, n_obs, n_rays, top_k;
  double comm_radius, car_radius, obs_radius, area_size, dt, dist2goal, connect_radius;
};
DgppoEnvCfg make_cfg(const EnvAttrs& a, const float* goal_table) {
  DgppoEnvCfg c{};
  c.kind = a.kind; c.n_agents = a.n_agents; c.n_obs = a.n_obs; c.n_rays = a.n_rays; c.top_k = a.top_k;
  c.comm_radius = a.comm_radius; c.car_radius = a.car_radius; c.obs_radius = a.obs_radius;
  c.area_size = a.area_size; c.dt = a.dt; c.dist2goal = a.dist2goal; c.connect_radius = a.connect_radius;
  c.goal_table = goal_table;
  return c;
}
int32_t leading(const ffi::AnyBuffer::Dimensions& d, size_t trailing) {      // product of the batch axes
  int64_t b = 1;
  for (size_t i = 0; i + trailing < d.size(); ++i) b *= d[i];
  return (int32_t)b;
}
const float* opt(const F32& b) { return b.element_count() ? b.typed_data() : nullptr; }   // size-0 buffer == NULL

#define ENV_ATTR_ARGS                                                                                          \
  int32_t kind, int32_t n_agents, int32_t n_obs, int32_t n_rays, int32_t top_k, double comm_radius,            \
      double car_radius, double obs_radius, double area_size, double dt, double dist2goal, double connect_radius
#define ENV_ATTR_PACK                                                                                          \
  EnvAttrs { kind, n_agents, n_obs, n_rays, top_k, comm_radius, car_radius, obs_radius, area_size, dt, dist2goal, \
             connect_radius }
#define ENV_ATTRS                                                                                              \
  .Attr<int32_t>("kind").Attr<int32_t>("n_agents").Attr<int32_t>("n_obs").Attr<int32_t>("n_rays")              \
      .Attr<int32_t>("top_k").Attr<double>("comm_radius").Attr<double>("car_radius").Attr<double>("obs_radius") \
      .Attr<double>("area_size").Attr<double>("dt").Attr<double>("dist2goal").Attr<double>("connect_radius")
#define NET_ATTR_ARGS int32_t net_kind, int32_t node_dim, int32_t n_layers, int32_t n_out
#define NET_ATTRS .Attr<int32_t>("net_kind").Attr<int32_t>("node_dim").Attr<int32_t>("n_layers").Attr<int32_t>("n_out")

// ---- K0 reset: keys (b) u64 -> agent, goal, obstacles, n_draws
ffi::Error ResetImpl(cudaStream_t st, U64 keys, F32 goal_table, RF32 agent, RF32 goal, RF32 obstacles, RS32 n_draws,
                     ENV_ATTR_ARGS, double obs_len_lo, double obs_len_hi, double theta_lo, double theta_hi) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, opt(goal_table));
  return status(dgppo_reset(st, &cfg, keys.typed_data(), obs_len_lo, obs_len_hi, theta_lo, theta_hi,
                            agent->typed_data(), goal->typed_data(), n_obs ? obstacles->typed_data() : nullptr,
                            n_draws->typed_data(), (int32_t)keys.element_count()), "dgppo_reset");
}

// ---- K1 env step: agent (b,n,sd), goal (b,g,sd), obs_nodes, action (b,n,2) -> next_agent, reward (b), cost (b,n,nh)
ffi::Error EnvStepImpl(cudaStream_t st, F32 agent, F32 goal, F32 obs_nodes, F32 action, F32 goal_table, RF32 next_agent,
                       RF32 reward, RF32 cost, ENV_ATTR_ARGS) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, opt(goal_table));
  return status(dgppo_env_step(st, &cfg, agent.typed_data(), goal.typed_data(), opt(obs_nodes), action.typed_data(),
                               next_agent->typed_data(), reward->typed_data(), cost->typed_data(), 1,
                               leading(agent.dimensions(), 2)), "dgppo_env_step");
}

// ---- K2 LiDAR: agent (b,n,sd), obstacles (b,n_obs,16), ray_dirs (n_rays,2) -> hits (b,n,top_k,2)
ffi::Error LidarImpl(cudaStream_t st, F32 agent, F32 obstacles, F32 ray_dirs, RF32 hits, ENV_ATTR_ARGS) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, nullptr);
  return status(dgppo_lidar(st, &cfg, agent.typed_data(), obstacles.typed_data(), ray_dirs.typed_data(),
                            hits->typed_data(), leading(agent.dimensions(), 2)), "dgppo_lidar");
}

// ---- K3 graph: -> the eight GraphsTuple array fields
ffi::Error BuildGraphImpl(cudaStream_t st, F32 agent, F32 goal, F32 obs_nodes, RF32 nodes, RF32 edges, RF32 states,
                          RS32 receivers, RS32 senders, RS32 node_type, RS32 n_node, RS32 n_edge, ENV_ATTR_ARGS) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, nullptr);
  return status(dgppo_build_graph(st, &cfg, agent.typed_data(), goal.typed_data(), opt(obs_nodes), nodes->typed_data(),
                                  edges->typed_data(), states->typed_data(), receivers->typed_data(),
                                  senders->typed_data(), node_type->typed_data(), n_node->typed_data(),
                                  n_edge->typed_data(), 1, leading(agent.dimensions(), 2)), "dgppo_build_graph");
}

// ---- K4a policy: packed params, graph arrays, rnn (b,n,64), eps (b,n,2) or size 0 -> rnn_out, action, log_pi
ffi::Error GnnPolicyImpl(cudaStream_t st, F32 params, F32 nodes, F32 edges, S32 receivers, S32 senders, F32 rnn_in,
                         F32 eps, RF32 rnn_out, RF32 action, RF32 log_pi, ENV_ATTR_ARGS, NET_ATTR_ARGS) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, nullptr);
  const DgppoNetCfg net{net_kind, node_dim, 4, n_layers, n_out};
  return status(dgppo_gnn_policy(st, &cfg, &net, params.typed_data(), nodes.typed_data(), edges.typed_data(),
                                 receivers.typed_data(), senders.typed_data(), 1, rnn_in.typed_data(),
                                 rnn_out->typed_data(), 1, opt(eps), 1, action->typed_data(),
                                 eps.element_count() ? log_pi->typed_data() : nullptr, 1,
                                 leading(nodes.dimensions(), 2)), "dgppo_gnn_policy");
}

// ---- K4b value (Vh per agent | Vl pooled): graph record (b, n_slots, ...), rnn record -> value, rnn_out scratch
ffi::Error GnnValueImpl(cudaStream_t st, F32 params, F32 nodes, F32 edges, S32 receivers, S32 senders, F32 rnn_in,
                        RF32 rnn_out, RF32 value, ENV_ATTR_ARGS, NET_ATTR_ARGS, int32_t n_slots) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, nullptr);
  const DgppoNetCfg net{net_kind, node_dim, 4, n_layers, n_out};
  const int32_t b = leading(nodes.dimensions(), 2) / n_slots;
  return status(dgppo_gnn_value(st, &cfg, &net, params.typed_data(), nodes.typed_data(), edges.typed_data(),
                                receivers.typed_data(), senders.typed_data(), n_slots, rnn_in.typed_data(),
                                rnn_out->typed_data(), n_slots, value->typed_data(), n_slots, n_slots, b),
                "dgppo_gnn_value");
}

// ---- Vl scan: graph record (b, n_slots, ...); carry (b, n_slots + 1, 64) result with slot 0 = carry_in
ffi::Error VlScanImpl(cudaStream_t st, F32 params, F32 nodes, F32 edges, S32 receivers, S32 senders, F32 carry_in,
                      RF32 carry, RF32 value, ENV_ATTR_ARGS, NET_ATTR_ARGS, int32_t n_slots) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, nullptr);
  const DgppoNetCfg net{net_kind, node_dim, 4, n_layers, n_out};
  const int32_t b = leading(nodes.dimensions(), 2) / n_slots;
  cudaError_t e = cudaMemcpy2DAsync(carry->typed_data(), (size_t)(n_slots + 1) * 64 * sizeof(float), carry_in.typed_data(),
                                    64 * sizeof(float), 64 * sizeof(float), (size_t)b, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return status((int)e, "dgppo_vl_scan (carry copy)");
  return status(dgppo_vl_scan(st, &cfg, &net, params.typed_data(), nodes.typed_data(), edges.typed_data(),
                              receivers.typed_data(), senders.typed_data(), n_slots, carry->typed_data(), n_slots + 1,
                              value->typed_data(), n_slots, n_slots, b), "dgppo_vl_scan");
}

// ---- K5 GAE: hs (b,T,n,nh), l (b,T), Vh (b,T+1,n,nh), Vl (b,T+1) -> Qh, Ql
ffi::Error GaeImpl(cudaStream_t st, F32 hs, F32 l, F32 Vh, F32 Vl, RF32 Qh, RF32 Ql, float gamma, float gae_lambda) {
  const auto d = hs.dimensions();
  if (d.size() < 4) return ffi::Error::InvalidArgument("dgppo_gae: hs must be (b, T, n, nh)");
  const size_t r = d.size();
  return status(dgppo_gae(st, hs.typed_data(), l.typed_data(), Vh.typed_data(), Vl.typed_data(), gamma, gae_lambda,
                          Qh->typed_data(), Ql->typed_data(), leading(d, 3), (int32_t)d[r - 3], (int32_t)d[r - 2],
                          (int32_t)d[r - 1]), "dgppo_gae");
}

// ---- CBF residual + advantage merge
ffi::Error CbfAdvantageImpl(cudaStream_t st, F32 Ql, F32 Vl, F32 Vh, RF32 A, RF32 cbf_deriv, RF32 Acbf, RU8 is_safe,
                            float dt, float alpha, float cbf_eps, float cbf_weight) {
  const auto d = Vh.dimensions();
  if (d.size() < 4) return ffi::Error::InvalidArgument("dgppo_cbf_advantage: Vh must be (b, T+1, n, nh)");
  const size_t r = d.size();
  return status(dgppo_cbf_advantage(st, Ql.typed_data(), Vl.typed_data(), Vh.typed_data(), dt, alpha, cbf_eps, cbf_weight,
                                    A->typed_data(), cbf_deriv->typed_data(), Acbf->typed_data(), is_safe->typed_data(),
                                    leading(d, 3), (int32_t)d[r - 3] - 1, (int32_t)d[r - 2], (int32_t)d[r - 1]),
                "dgppo_cbf_advantage");
}

// ---- fused rollout: the lax.scan of trainer/utils.py:45-57 as ONE call.  Inputs: packed policy params, the reset
// graph (slot 0) and state, eps (b,T,n,2) or size 0.  Results: the (b, T+1, ...) record + (b, T, ...) outputs and the
// workspaces (declared as results so XLA owns them).
ffi::Error RolloutImpl(cudaStream_t st, F32 params, F32 nodes0, F32 edges0, F32 states0, S32 receivers0, S32 senders0,
                       S32 node_type0, F32 rnn0, F32 agent0, F32 hits0, F32 goal, F32 obstacles, F32 ray_dirs, F32 eps,
                       F32 goal_table, RF32 nodes, RF32 edges, RF32 states, RS32 receivers, RS32 senders, RS32 node_type,
                       RS32 n_node, RS32 n_edge, RF32 rnn, RF32 actions, RF32 log_pis, RF32 rewards, RF32 costs,
                       RF32 agent_ws, RF32 hits_ws, ENV_ATTR_ARGS, NET_ATTR_ARGS, int32_t T) {
  const DgppoEnvCfg cfg = make_cfg(ENV_ATTR_PACK, opt(goal_table));
  const DgppoNetCfg net{net_kind, node_dim, 4, n_layers, n_out};
  const int32_t b = leading(nodes0.dimensions(), 2);
  const size_t P = (size_t)T + 1;
  // slot 0 of every record field <- the reset graph; workspaces <- the reset state (strided device copies)
  auto put0 = [&](void* dst, const void* src, size_t row_bytes) {
    return cudaMemcpy2DAsync(dst, P * row_bytes, src, row_bytes, row_bytes, (size_t)b, cudaMemcpyDeviceToDevice, st);
  };
  auto per_env = [&](const ffi::AnyBuffer::Dimensions& d, size_t elt) { size_t m = elt; for (size_t i = d.size() - 2; i < d.size(); ++i) m *= (size_t)d[i]; return m; };
  cudaError_t e = put0(nodes->typed_data(), nodes0.typed_data(), per_env(nodes0.dimensions(), 4));
  if (e == cudaSuccess) e = put0(edges->typed_data(), edges0.typed_data(), per_env(edges0.dimensions(), 4));
  if (e == cudaSuccess) e = put0(states->typed_data(), states0.typed_data(), per_env(states0.dimensions(), 4));
  const size_t E = (size_t)receivers0.dimensions().back(), N = (size_t)node_type0.dimensions().back();
  if (e == cudaSuccess) e = put0(receivers->typed_data(), receivers0.typed_data(), E * 4);
  if (e == cudaSuccess) e = put0(senders->typed_data(), senders0.typed_data(), E * 4);
  if (e == cudaSuccess) e = put0(node_type->typed_data(), node_type0.typed_data(), N * 4);
  if (e == cudaSuccess) e = put0(rnn->typed_data(), rnn0.typed_data(), per_env(rnn0.dimensions(), 4));
  if (e == cudaSuccess) e = cudaMemcpyAsync(agent_ws->typed_data(), agent0.typed_data(), agent0.size_bytes(), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && hits0.element_count())
    e = cudaMemcpyAsync(hits_ws->typed_data(), hits0.typed_data(), hits0.size_bytes(), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return status((int)e, "dgppo_rollout (slot 0 copies)");
  DgppoRolloutBuffers B{};
  B.nodes = nodes->typed_data(); B.edges = edges->typed_data(); B.states = states->typed_data();
  B.receivers = receivers->typed_data(); B.senders = senders->typed_data(); B.node_type = node_type->typed_data();
  B.n_node = n_node->typed_data(); B.n_edge = n_edge->typed_data(); B.rnn = rnn->typed_data(); B.eps = opt(eps);
  B.actions = actions->typed_data(); B.log_pis = eps.element_count() ? log_pis->typed_data() : nullptr;
  B.rewards = rewards->typed_data(); B.costs = costs->typed_data(); B.agent_ws = agent_ws->typed_data();
  B.hits_ws = hits0.element_count() ? hits_ws->typed_data() : nullptr;
  B.goal = goal.typed_data(); B.obstacles = opt(obstacles); B.ray_dirs = opt(ray_dirs); B.hits_ws2 = nullptr;
  return status(dgppo_rollout(st, &cfg, &net, params.typed_data(), &B, T, b, nullptr), "dgppo_rollout");
}

}  // namespace

#define STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoReset, ResetImpl,
    STREAM.Arg<U64>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>() ENV_ATTRS
        .Attr<double>("obs_len_lo").Attr<double>("obs_len_hi").Attr<double>("theta_lo").Attr<double>("theta_hi"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoEnvStep, EnvStepImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>() ENV_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoLidar, LidarImpl, STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>() ENV_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoBuildGraph, BuildGraphImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<S32>().Ret<S32>().Ret<S32>()
        .Ret<S32>() ENV_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoGnnPolicy, GnnPolicyImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        ENV_ATTRS NET_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoGnnValue, GnnValueImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Ret<F32>().Ret<F32>() ENV_ATTRS NET_ATTRS
        .Attr<int32_t>("n_slots"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoVlScan, VlScanImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Ret<F32>().Ret<F32>() ENV_ATTRS NET_ATTRS
        .Attr<int32_t>("n_slots"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoGae, GaeImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Attr<float>("gamma").Attr<float>("gae_lambda"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoCbfAdvantage, CbfAdvantageImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U8>().Attr<float>("dt").Attr<float>("alpha")
        .Attr<float>("cbf_eps").Attr<float>("cbf_weight"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DgppoRollout, RolloutImpl,
    STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<S32>().Ret<S32>().Ret<S32>().Ret<S32>().Ret<F32>().Ret<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>() ENV_ATTRS NET_ATTRS.Attr<int32_t>("T"));

#endif  // DGPPO_HAVE_XLA_FFI
