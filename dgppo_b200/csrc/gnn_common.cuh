// Shared definitions of the GNN forward kernels (v1 fused kernel in
// gnn_kernels.cu, v2 weight-stationary pair in gnn_v2.cu).
#pragma once
#include "common.cuh"

namespace dgppo {

constexpr int R = 64;            // agent rows per tile
constexpr int RS = 68;           // row stride of transposed buffers (floats)
constexpr int NT = 256;          // threads per CTA
constexpr int H = 3;             // attention heads
constexpr int HID = 64;          // head / GRU width
constexpr int X0S = 8;           // node-major stride of input node features
constexpr int X1S = 36;          // node-major stride of layer-1 outputs (32 + pad)

struct LayerP { const float *wq, *bq, *wkt, *wagg, *wu, *bu, *wqk; int in, d; };

struct NetP {
  LayerP L[2];
  int n_layers, kind, n_out;
  const float *d0w, *d0b, *ln0s, *ln0b, *d1w, *d1b, *ln1s, *ln1b;
  const float *wi, *bi, *wh, *bhn, *out_w, *out_b;
};

struct GnnArgs {
  const float* nodes; const float* edges; const int* recv; const int* send;
  int pitch, n_slots;
  const float* rnn_in; float* rnn_out; int rnn_pitch;
  const float* eps; int eps_pitch;
  float* action; float* log_pi; int act_pitch;
  float* value; int out_pitch;
  int n_graphs;                 // b * n_slots
  int n, N, E, nd, n_ag, n_ao, G;
  // "graph from state" mode (c_agent != nullptr): nodes / edges / recv / send are not read; the tile's node
  // features, slot masks and edge features are formed from K3's INPUTS with K3's arithmetic (csrc/env_kernels.cu
  // build_graph_kernel), so the policy sees bit for bit the graph K3 would have written to the record.
  const float* c_agent;         // (b, c_pitch, n, sd) slot pointer
  const float* c_obs;           // Lidar: hits (b, c_pitch, n, top_k, 2) slot pointer; MPE: obstacles (b, n_obs, 4), static
  const float* c_goal;          // (b, g, sd), static per env
  int c_pitch, g_nodes, sd, c_lidar, c_paired;
  float cR, cR_diag, cR_obs, cR_mpe_obs;
};

// one node row of a graph in the layout of the x0 tile ([state | one-hot(obstacle, goal, agent)], stride X0S),
// from the state arrays (lidar_env/base.py:234-264, mpe/base.py:214-233); row in [0, N - 1)
__device__ __forceinline__ void node_row_from_state(const GnnArgs& g, int env, int slot, int row, float* dst) {
  const int n = g.n, gn = g.g_nodes, sd = g.sd;
  float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;      // columns 0-3, 4-7 (no indexed local array: registers)
  int type;                                                  // 0 agent, 1 goal, 2 obstacle
  if (row < n + gn) {
    type = row < n ? 0 : 1;
    const float* s = type == 0 ? g.c_agent + (((size_t)env * g.c_pitch + slot) * n + row) * sd
                               : g.c_goal + ((size_t)env * gn + (row - n)) * sd;
    if (sd == 4) {
      lo = __ldg(reinterpret_cast<const float4*>(s));
    } else {
      lo = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
      hi.x = __ldg(s + 4);
    }
  } else {
    type = 2;
    const int o = row - n - gn, n_on = g.N - 1 - n - gn;
    if (g.c_lidar) {
      const float2 h = __ldg(reinterpret_cast<const float2*>(g.c_obs + (((size_t)env * g.c_pitch + slot) * (size_t)n_on + o) * 2));
      lo.x = h.x; lo.y = h.y;
    } else {
      lo = __ldg(reinterpret_cast<const float4*>(g.c_obs + ((size_t)env * n_on + o) * 4));
    }
  }
  // one-hot at column sd + (2 - type): [obstacle, goal, agent] = columns sd, sd + 1, sd + 2
  const int c1 = sd + 2 - type;
  if (c1 == 4) hi.x = 1.f; else if (c1 == 5) hi.y = 1.f; else if (c1 == 6) hi.z = 1.f; else hi.w = 1.f;
  reinterpret_cast<float4*>(dst)[0] = lo;
  reinterpret_cast<float4*>(dst)[1] = hi;
}

// state2feat of a staged node row (identity, or [x, y, v cos, v sin] for the bicycle: lidar_bicycle_target.py:113-118)
__device__ __forceinline__ float4 feat_of_row(const float* x, bool bic) {
  return bic ? make_float4(x[0], x[1], fmul(x[4], x[2]), fmul(x[4], x[3])) : make_float4(x[0], x[1], x[2], x[3]);
}
// slot t of receiver agent i: sender node (graph-local) and whether the slot is live
// (lidar_spread.py:57-96, lidar_target.py:57-96, mpe_spread.py:51-81; K3's edge loop)
__device__ __forceinline__ int slot_sender_from_state(const GnnArgs& g, const float* xg /* the graph's x0 rows */,
                                                      int i, int t, bool& live) {
  const int n = g.n;
  const float* a = xg + i * X0S;
  if (t < n) {
    const float* c = xg + t * X0S;
    float dist = norm2(fsub(a[0], c[0]), fsub(a[1], c[1]));
    dist = fadd(dist, (i == t) ? g.cR_diag : 0.f);
    live = dist < g.cR;
    return t;
  }
  if (t < n + g.n_ag) {
    live = true;
    return g.c_paired ? n + i : n + (t - n);
  }
  const int q = t - n - g.n_ag;
  const int s = n + g.g_nodes + (g.c_lidar ? i * g.n_ao + q : q);
  const float* c = xg + s * X0S;
  live = norm2(fsub(a[0], c[0]), fsub(a[1], c[1])) < (g.c_lidar ? g.cR_obs : g.cR_mpe_obs);
  return s;
}
// edge feature of slot t (receiver row a, sender row c of the x0 tile), K3's expressions
__device__ __forceinline__ float4 edge_from_state(const GnnArgs& g, const float* a, const float* c, int t) {
  if (t >= g.n + g.n_ag && g.c_lidar) return make_float4(fsub(a[0], c[0]), fsub(a[1], c[1]), 0.f, 0.f);
  const bool bic = g.sd == 5;
  const float4 fa = feat_of_row(a, bic), fc = feat_of_row(c, bic && t < g.n + g.n_ag);
  return make_float4(fsub(fa.x, fc.x), fsub(fa.y, fc.y), fsub(fa.z, fc.z), fsub(fa.w, fc.w));
}

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// Tensor-core head (head_tc.cu): B operands of the 3xTF32 products, hi copy then lo copy, each in the
// no-swizzle K-major canonical layout [K/4 chunks][n][4]:  block[kc][n][j] = W[4 kc + j][n].
constexpr int TC_D_FL = 2 * 64 * 64;      // one Dense64 block (floats, hi + lo): 32 KB
constexpr int TC_G_FL = 2 * 64 * 192;     // one GRU block (x part [ir|iz|in] or h part [hr|hz|hn]): 96 KB
constexpr int TC_HEAD_FL = 2 * TC_D_FL + 2 * TC_G_FL;   // [Dense0 | Dense1 | GRU x | GRU h] = 256 KB


// GRU gate activations on the SFU (ex2.approx / rcp.approx): absolute error below 3e-7 on outputs in
// (0, 1) / (-1, 1), inside the fp32 tolerance of the path (rtol 1e-5).
__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) {
  const float ax = fminf(fabsf(x), 15.f);                  // tanh(15) == 1 in fp32; keeps exp finite
  const float t = 1.f - __fdividef(2.f, __expf(2.f * ax) + 1.f);
  return copysignf(t, x);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// tfp special_math.log_ndtr, float32 segments (lower -10, upper 5, 3-term series)
__device__ __forceinline__ float ndtrf_(float x) {
  const float hs2 = 0.70710678118654752440f;
  const float w = x * hs2, z = fabsf(w);
  const float y = (z < hs2) ? 1.f + erff(w) : ((w > 0.f) ? 2.f - erfcf(z) : erfcf(z));
  return 0.5f * y;
}
__device__ __forceinline__ float log_ndtrf_(float x) {
  if (x > 5.f) return -ndtrf_(-x);
  if (x > -10.f) return logf(ndtrf_(fmaxf(x, -10.f)));
  const float xl = fminf(x, -10.f), x2 = xl * xl;
  const float series = 1.f - 1.f / x2 + 3.f / (x2 * x2) - 15.f / (x2 * x2 * x2);
  return -0.5f * x2 - logf(-xl) - 0.91893853320467274178f + logf(series);
}
// TanhTransformedDistribution.log_prob for one action component
// (algo/module/distribution.py:25-35; tfd.Normal.log_prob; tfb.Tanh fldj)
__device__ __forceinline__ float tanh_normal_logp(float value, float loc, float scale) {
  const float thr = 0.999f;
  const float inv_thr = atanhf(thr);
  const float log_eps = (float)-6.907755278982136;        // np.log(1.0 - 0.999)
  const float v = fminf(fmaxf(value, -thr), thr);
  if (v <= -thr) return log_ndtrf_((-inv_thr - loc) / scale) - log_eps;
  if (v >= thr) return log_ndtrf_(-((inv_thr - loc) / scale)) - log_eps;
  const float x = atanhf(v);
  const float fldj = 2.f * (0.69314718055994530942f - x - softplusf_(-2.f * x));
  const float d = x / scale - loc / scale;
  const float lp = -0.5f * d * d - (0.91893853320467274178f + logf(scale));
  return lp - fldj;
}

// Policy tail for one agent row: TanhNormal heads -> action (+ log_pi)
// (policy.py:61-74,191-203).  o4 = [mean0, mean1, std_trans0, std_trans1].
__device__ __forceinline__ void policy_tail(const GnnArgs& g, float m0, float m1, float t0, float t1,
                                            int env, int slot, int i, int n) {
  const size_t ao = ((size_t)env * g.act_pitch + slot) * n + i;
  if (g.eps) {
    const float inv = -0.43275212956718856f;     // log(exp(0.5) - 1)  (policy.py:54-59)
    const float s0 = softplusf_(t0 + inv) + 1e-5f;
    const float s1 = softplusf_(t1 + inv) + 1e-5f;
    const float* ep = g.eps + (((size_t)env * g.eps_pitch + slot) * n + i) * 2;
    const float a0 = tanhf(fmaf(s0, __ldg(ep), m0));
    const float a1 = tanhf(fmaf(s1, __ldg(ep + 1), m1));
    g.action[ao * 2] = a0; g.action[ao * 2 + 1] = a1;
    if (g.log_pi) g.log_pi[ao] = tanh_normal_logp(a0, m0, s0) + tanh_normal_logp(a1, m1, s1);
  } else {
    g.action[ao * 2] = tanhf(m0); g.action[ao * 2 + 1] = tanhf(m1);   // mode (distribution.py:45-46)
  }
}

// v2 launcher (gnn_v2.cu): returns DGPPO_V2_UNSUPPORTED when the shape does not
// fit its shared-memory plan; the caller then uses the v1 fused kernel.
#define DGPPO_V2_UNSUPPORTED (-100)
// phase: 0 = GNN layers + head, 1 = GNN layers only (embeddings left in rnn_out), 2 = head only
// (embeddings expected in rnn_out); 1 and 2 serve the Vl scan, whose GNN part has no recurrence.
int launch_gnn_v2(void* stream, const NetP& P, const DgppoNetLayout& L, const float* params,
                  const GnnArgs& g, int sms, int phase = 0);

// Tensor-core head (head_tc.cu) over the rows of g (embeddings in rnn_out); tcw = params + layout.tc_head.
int launch_head_tc(cudaStream_t st, const NetP& P, const GnnArgs& g, const float* tcw, int sms, bool pdl);
size_t head_tc_smem_bytes();

}  // namespace dgppo
