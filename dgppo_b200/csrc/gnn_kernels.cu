// K4: GraphTransformer policy / value forward for sm_100a (FP32 FFMA path).
//
// One CTA processes a TILE of G graphs at a time (G*n <= 64 agent rows) and
// keeps every intermediate in shared memory; the grid is persistent over
// tiles.  All dense contractions run as register-tiled (4 rows x 4 cols)
// GEMMs whose A operand is a transposed [feature][row] shared-memory buffer
// and whose B operand is the packed weight matrix streamed through L1
// (weights are shared by every CTA and stay L2-resident).
//
// GNN regrouping: the reference (nn/gnn.py:85-117) projects
// q/k/v/e per EDGE.  Receivers are agents only, so per receiver i and head h
//     score(e) = q_h[i] . (Wk_h x_s + bk_h) = (Wk_h^T q_h[i]) . x_s + q_h[i] . bk_h
//     agg[i]   = 1/H sum_h [ Wv_h (sum_e a_e x_s) + bv_h (sum_e a_e) + We_h (sum_e a_e edge_e) ]
// which needs no per-sender key/value projection at all: ~6x fewer FLOPs at
// LidarSpread n=8 and no (E, 3, d) intermediates.  Masked edges (recv = pad)
// only feed the pad node, whose output no agent row ever reads.
#include <stdlib.h>

#include "gnn_common.cuh"

namespace dgppo {

// ---------------------------------------------------------------- tile GEMM
// acc[i][j] (row r0+i, col c0+j) = scale1 * sum_k A1[k][r] W1(k)[c] + sum_k A2[k][r] W2(k)[c]
struct RowPtr {                      // contiguous weight rows
  const float* base; int ld;
  __device__ __forceinline__ const float* operator()(int k) const { return base + (size_t)k * ld; }
};

template <class Epi>
__device__ __forceinline__ void tile_gemm(const float* A1, int K1, RowPtr W1, float scale1,
                                          const float* A2, int K2, RowPtr W2,
                                          int N4, Epi epi) {
  const int items = (R / 4) * N4;
  for (int item = threadIdx.x; item < items; item += NT) {
    const int cg = item % N4, rg = item / N4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* a_ptr = A1 + rg * 4;
#pragma unroll 4
    for (int k = 0; k < K1; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(a_ptr + k * RS);
      const float4 w = __ldg(reinterpret_cast<const float4*>(W1(k)) + cg);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (K2 > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] *= scale1;
      a_ptr = A2 + rg * 4;
#pragma unroll 4
      for (int k = 0; k < K2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(a_ptr + k * RS);
        const float4 w = __ldg(reinterpret_cast<const float4*>(W2(k)) + cg);
        const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
    }
    epi(rg * 4, cg * 4, acc);
  }
}

// store acc (+bias, optional relu) into a transposed [col][row] buffer
struct StoreT {
  float* out; const float* bias; bool relu;
  __device__ __forceinline__ void operator()(int r0, int c0, float (&acc)[4][4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float bj = bias ? __ldg(bias + c0 + j) : 0.f;
      float4 v = make_float4(acc[0][j] + bj, acc[1][j] + bj, acc[2][j] + bj, acc[3][j] + bj);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(out + (c0 + j) * RS + r0) = v;
    }
  }
};

// ------------------------------------------------------------ attention
// One thread per (row, head): two passes over the row's static edge slots
// (max, then exp-sum / weighted sums), reading sender features node-major.
// Writes z[h*INA + c][r]: c < IN  sum_e a_e x_s[c];  c == IN  sum_e a_e;
//                        c > IN  sum_e a_e edge_e[c-IN-1].
template <int IN_MAX>
__device__ __forceinline__ void attention(const GnnArgs& g, int tile0, int rows,
                                          const float* qt /*[H][INP][RS]*/, int IN, int INP, float inv_sqrt_d,
                                          const float* X, int XS, float* z /*[H*INA][RS]*/) {
  const int INA = IN + 5;
  const int n = g.n, deg = n + g.n_ag + g.n_ao, pad = g.N - 1, nodes_per = g.N - 1;
  for (int item = threadIdx.x; item < R * H; item += NT) {
    const int r = item / H, h = item - r * H;
    float* zc = z + (h * INA) * RS + r;
    if (r >= rows) {
      for (int c = 0; c < INA; ++c) zc[c * RS] = 0.f;
      continue;
    }
    const int gl = r / n, i = r - gl * n;
    const int gi = tile0 + gl;
    const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
    const size_t gslot = (size_t)env * g.pitch + slot;
    const int* rc = g.recv + gslot * g.E;
    const int* sn = g.send + gslot * g.E;
    const float4* ed = reinterpret_cast<const float4*>(g.edges + gslot * g.E * 4);
    const float* Xg = X + (size_t)gl * nodes_per * XS;

    float qv[IN_MAX];
    const float* qh = qt + (h * INP) * RS + r;
#pragma unroll
    for (int c = 0; c < IN_MAX; ++c) qv[c] = (c < IN) ? qh[c * RS] : 0.f;
    const float qc = qh[IN * RS];

    auto slot_of = [&](int t) {
      if (t < n) return i * n + t;
      if (t < n + g.n_ag) return n * n + i * g.n_ag + (t - n);
      return n * n + n * g.n_ag + i * g.n_ao + (t - n - g.n_ag);
    };
    auto score = [&](int s) {
      const float* x = Xg + s * XS;
      float acc = qc;
#pragma unroll
      for (int c = 0; c < IN_MAX; c += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + c);
        acc = fmaf(qv[c], xv.x, acc);
        if (c + 1 < IN_MAX) acc = fmaf(qv[c + 1], xv.y, acc);
        if (c + 2 < IN_MAX) acc = fmaf(qv[c + 2], xv.z, acc);
        if (c + 3 < IN_MAX) acc = fmaf(qv[c + 3], xv.w, acc);
      }
      return acc * inv_sqrt_d;
    };

    float mx = -INFINITY;
    for (int t = 0; t < deg; ++t) {
      const int e = slot_of(t);
      if (rc[e] == pad) continue;
      mx = fmaxf(mx, score(sn[e]));
    }
    float xb[IN_MAX];
#pragma unroll
    for (int c = 0; c < IN_MAX; ++c) xb[c] = 0.f;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, l = 0.f;
    for (int t = 0; t < deg; ++t) {
      const int e = slot_of(t);
      if (rc[e] == pad) continue;
      const int s = sn[e];
      const float p = expf(score(s) - mx);
      l += p;
      const float* x = Xg + s * XS;
#pragma unroll
      for (int c = 0; c < IN_MAX; c += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + c);
        xb[c] = fmaf(p, xv.x, xb[c]);
        if (c + 1 < IN_MAX) xb[c + 1] = fmaf(p, xv.y, xb[c + 1]);
        if (c + 2 < IN_MAX) xb[c + 2] = fmaf(p, xv.z, xb[c + 2]);
        if (c + 3 < IN_MAX) xb[c + 3] = fmaf(p, xv.w, xb[c + 3]);
      }
      const float4 ef = __ldg(ed + e);
      e0 = fmaf(p, ef.x, e0); e1 = fmaf(p, ef.y, e1); e2 = fmaf(p, ef.z, e2); e3 = fmaf(p, ef.w, e3);
    }
    const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
#pragma unroll
    for (int c = 0; c < IN_MAX; ++c)
      if (c < IN) zc[c * RS] = xb[c] * inv_l;
    zc[IN * RS] = (l > 0.f) ? 1.f : 0.f;
    zc[(IN + 1) * RS] = e0 * inv_l; zc[(IN + 2) * RS] = e1 * inv_l;
    zc[(IN + 3) * RS] = e2 * inv_l; zc[(IN + 4) * RS] = e3 * inv_l;
  }
}

// LayerNorm (flax: eps 1e-6, fast variance) + ReLU over the 64 features of
// each row of a transposed buffer, in place.  One thread per row.
__device__ __forceinline__ void layernorm_relu(float* y, const float* scale, const float* bias) {
  if (threadIdx.x < R) {
    float* col = y + threadIdx.x;
    float s = 0.f, s2 = 0.f;
#pragma unroll 8
    for (int c = 0; c < HID; ++c) { const float v = col[c * RS]; s += v; s2 = fmaf(v, v, s2); }
    const float mean = s * (1.f / HID), mean2 = s2 * (1.f / HID);
    const float var = fmaxf(0.f, mean2 - mean * mean);
    const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll 8
    for (int c = 0; c < HID; ++c) {
      const float mul = rstd * __ldg(scale + c);
      col[c * RS] = fmaxf((col[c * RS] - mean) * mul + __ldg(bias + c), 0.f);
    }
  }
}

// ------------------------------------------------------------------ kernel
// Shared memory map (floats):
//   x0   [M][X0S]   input node features, node-major (M = G*(N-1) <= m_cap)
//   x1   [M][X1S]   layer-1 outputs (2-layer nets only)
//   bufA..bufD      transposed row buffers (sizes below)
struct Smem {
  float *x0, *x1, *xr, *q, *qt, *z, *y0, *y1, *hbuf;
};

template <int NL>
__global__ void __launch_bounds__(NT, 1)
gnn_forward_kernel(NetP net, GnnArgs g, int m_cap) {
  extern __shared__ __align__(16) float smem[];
  // carve (region sizes must match launch_gnn):
  //   x0r  max(m_cap*X0S, HID*RS)  input node features (node-major); reused as y1 by the head
  //   x1   m_cap*X1S               layer-1 outputs, node-major (2-layer nets only)
  //   xr   [32][RS]                agent-row inputs of the current layer (transposed)
  //   q    [192][RS]               queries; reused as z (attention sums) and as the tail output
  //   qt   [H*36][RS]              regrouped keys; reused as the GRU carry hbuf
  //   y0   [64][RS]
  const size_t x0_fl = (size_t)m_cap * X0S > (size_t)HID * RS ? (size_t)m_cap * X0S : (size_t)HID * RS;
  float* x0 = smem;
  float* x1 = x0 + x0_fl;
  float* xr = x1 + (NL == 2 ? (size_t)m_cap * X1S : 0);
  float* q = xr + 32 * RS;
  float* qt = q + 192 * RS;
  float* y0 = qt + H * 36 * RS;
  float* y1 = x0;                                     // x0 is dead once layer 0 is done
  float* hbuf = qt;                                   // qt is dead once the last attention is done
  float* z = q;                                       // q is dead once qt is built

  const int n = g.n, N = g.N, nd = g.nd, G = g.G;
  const int nodes_per = N - 1;
  const int n_tiles = (g.n_graphs + G - 1) / G;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tile0 = tile * G;
    const int gcount = min(G, g.n_graphs - tile0);
    const int rows = gcount * n;
    const int M = gcount * nodes_per;
    __syncthreads();                                   // previous tile fully consumed

    // ---- load node features (node-major, stride X0S)
    for (int idx = threadIdx.x; idx < M * X0S; idx += NT) {
      const int s = idx / X0S, c = idx - s * X0S;
      const int gl = s / nodes_per, node = s - gl * nodes_per;
      const int gi = tile0 + gl;
      const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
      const float* src = g.nodes + ((size_t)env * g.pitch + slot) * N * nd;
      x0[idx] = (c < nd) ? __ldg(src + node * nd + c) : 0.f;
    }
    __syncthreads();

    // ---- GNN layers
    const float* X = x0; int XS = X0S;
    float* gnn_out = y0;                               // [64][RS] final agent embeddings
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const LayerP& P = net.L[l];
      const int IN = P.in, D = P.d, HD = H * D, INP = round4(IN + 1), INA = IN + 5;
      const bool last = (l == NL - 1);
      // agent-row inputs, transposed: xr[c][r] = X[node(r)][c]
      for (int idx = threadIdx.x; idx < 32 * R; idx += NT) {
        const int c = idx / R, r = idx - c * R;
        float v = 0.f;
        if (r < rows && c < IN) { const int gl = r / n, i = r - gl * n; v = X[((size_t)gl * nodes_per + i) * XS + c]; }
        xr[c * RS + r] = v;
      }
      __syncthreads();
      // q = xr Wq + bq                                 (gnn.py:86-88, on receivers)
      tile_gemm(xr, IN, RowPtr{P.wq, HD}, 1.f, nullptr, 0, RowPtr{nullptr, 0}, HD / 4,
                StoreT{q, P.bq, false});
      __syncthreads();
      // qt_h = Wk_h^T q_h  (+ q_h . bk_h in column IN)
      for (int h = 0; h < H; ++h)
        tile_gemm(q + (h * D) * RS, D, RowPtr{P.wkt + (size_t)h * D * INP, INP}, 1.f,
                  nullptr, 0, RowPtr{nullptr, 0}, INP / 4, StoreT{qt + (h * INP) * RS, nullptr, false});
      __syncthreads();
      // segment softmax + weighted sums               (gnn.py:100-107,114)
      const float isd = 1.f / sqrtf((float)D);
      if (l == 0) attention<X0S>(g, tile0, rows, qt, IN, INP, isd, X, XS, z);
      else        attention<32>(g, tile0, rows, qt, IN, INP, isd, X, XS, z);
      __syncthreads();
      // x' = relu(x Wu + bu + 1/H * z Wagg)            (gnn.py:109-111)
      if (last) {
        tile_gemm(z, H * INA, RowPtr{P.wagg, D}, 1.f / H, xr, IN, RowPtr{P.wu, D}, D / 4,
                  StoreT{gnn_out, P.bu, true});
      } else {
        // D == 32: agent rows go back node-major into x1; other nodes get relu(x Wu + bu)
        auto epi = [&](int r0, int c0, float (&acc)[4][4]) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i;
            if (r < rows) {
              const int gl = r / n, ia = r - gl * n;
              float* dst = x1 + ((size_t)gl * nodes_per + ia) * X1S + c0;
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = fmaxf(acc[i][j] + __ldg(P.bu + c0 + j), 0.f);
            }
          }
        };
        tile_gemm(z, H * INA, RowPtr{P.wagg, D}, 1.f / H, xr, IN, RowPtr{P.wu, D}, D / 4, epi);
        // non-agent nodes: 8 threads (4 cols each) per node
        for (int idx = threadIdx.x; idx < M * 8; idx += NT) {
          const int s = idx >> 3, c0 = (idx & 7) * 4;
          const int node = s % nodes_per;
          if (node < n) continue;
          const float* x = X + (size_t)s * XS;
          float a0 = __ldg(P.bu + c0), a1 = __ldg(P.bu + c0 + 1), a2 = __ldg(P.bu + c0 + 2), a3 = __ldg(P.bu + c0 + 3);
          for (int c = 0; c < IN; ++c) {
            const float xv = x[c];
            const float4 w = __ldg(reinterpret_cast<const float4*>(P.wu + c * D + c0));
            a0 = fmaf(xv, w.x, a0); a1 = fmaf(xv, w.y, a1); a2 = fmaf(xv, w.z, a2); a3 = fmaf(xv, w.w, a3);
          }
          float* dst = x1 + (size_t)s * X1S + c0;
          dst[0] = fmaxf(a0, 0.f); dst[1] = fmaxf(a1, 0.f); dst[2] = fmaxf(a2, 0.f); dst[3] = fmaxf(a3, 0.f);
        }
        X = x1; XS = X1S;
      }
      __syncthreads();
    }

    // ---- centralised Vl: mean over the agents of each graph (value.py:28-31);
    //      from here on there is one row per graph.
    const int nr = (net.kind == DGPPO_NET_VL) ? 1 : n;   // head rows per graph
    const int hrows = gcount * nr;
    if (net.kind == DGPPO_NET_VL) {
      for (int idx = threadIdx.x; idx < HID * R; idx += NT) {
        const int c = idx / R, gl = idx - c * R;
        float sacc = 0.f;
        if (gl < gcount) {
          for (int i = 0; i < n; ++i) sacc += gnn_out[c * RS + gl * n + i];
          sacc = sacc / (float)n;
        }
        hbuf[c * RS + gl] = sacc;
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < HID * R; idx += NT) {
        const int c = idx / R, r = idx - c * R;
        gnn_out[c * RS + r] = hbuf[c * RS + r];
      }
      __syncthreads();
    }

    // ---- head MLP: 2 x [Dense64 -> LayerNorm -> ReLU]   (mlp.py:14-30)
    tile_gemm(gnn_out, HID, RowPtr{net.d0w, HID}, 1.f, nullptr, 0, RowPtr{nullptr, 0}, HID / 4,
              StoreT{y1, net.d0b, false});
    __syncthreads();
    layernorm_relu(y1, net.ln0s, net.ln0b);
    __syncthreads();
    tile_gemm(y1, HID, RowPtr{net.d1w, HID}, 1.f, nullptr, 0, RowPtr{nullptr, 0}, HID / 4,
              StoreT{y0, net.d1b, false});
    __syncthreads();
    layernorm_relu(y0, net.ln1s, net.ln1b);
    __syncthreads();

    // GRU carry, transposed into hbuf (aliases qt)
    for (int idx = threadIdx.x; idx < R * HID; idx += NT) {
      const int r = idx / HID, c = idx - r * HID;
      float v = 0.f;
      if (r < hrows) {
        const int gl = r / nr, i = r - gl * nr;
        const int gi = tile0 + gl;
        const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
        v = __ldg(g.rnn_in + (((size_t)env * g.rnn_pitch + slot) * nr + i) * HID + c);
      }
      hbuf[c * RS + r] = v;
    }
    __syncthreads();

    // ---- GRU cell (flax GRUCell; rnn.py:19-21): one 4x4 (row, unit) tile per thread
    {
      const int cg = threadIdx.x & 15, rg = threadIdx.x >> 4;
      float ai[3][4][4], ah[3][4][4];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) { ai[t][i][j] = 0.f; ah[t][i][j] = 0.f; }
#pragma unroll 2
      for (int k = 0; k < HID; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(y0 + k * RS + rg * 4);
        const float4 ha = *reinterpret_cast<const float4*>(hbuf + k * RS + rg * 4);
        const float xv[4] = {xa.x, xa.y, xa.z, xa.w}, hv[4] = {ha.x, ha.y, ha.z, ha.w};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const float4 wi = __ldg(reinterpret_cast<const float4*>(net.wi + k * 192 + t * 64) + cg);
          const float4 wh = __ldg(reinterpret_cast<const float4*>(net.wh + k * 192 + t * 64) + cg);
          const float wiv[4] = {wi.x, wi.y, wi.z, wi.w}, whv[4] = {wh.x, wh.y, wh.z, wh.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ai[t][i][j] = fmaf(xv[i], wiv[j], ai[t][i][j]);
              ah[t][i][j] = fmaf(hv[i], whv[j], ah[t][i][j]);
            }
        }
      }
      __syncthreads();                                 // all reads of hbuf / y0 done
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cg * 4 + j;
        const float bir = __ldg(net.bi + c), biz = __ldg(net.bi + 64 + c), bin = __ldg(net.bi + 128 + c);
        const float bhn = __ldg(net.bhn + c);
        float hn[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float hprev = hbuf[c * RS + rg * 4 + i];
          const float rgate = sigmoidf_(ai[0][i][j] + bir + ah[0][i][j]);
          const float zgate = sigmoidf_(ai[1][i][j] + biz + ah[1][i][j]);
          const float cand = tanhf(ai[2][i][j] + bin + rgate * (ah[2][i][j] + bhn));
          hn[i] = (1.f - zgate) * cand + zgate * hprev;
        }
        *reinterpret_cast<float4*>(y1 + c * RS + rg * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
    __syncthreads();
    // new carry -> global (policy only)
    if (g.rnn_out) {
      for (int idx = threadIdx.x; idx < hrows * HID; idx += NT) {
        const int r = idx / HID, c = idx - r * HID;
        const int gl = r / nr, i = r - gl * nr;
        const int gi = tile0 + gl;
        const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
        g.rnn_out[(((size_t)env * g.rnn_pitch + slot) * nr + i) * HID + c] = y1[c * RS + r];
      }
    }

    // ---- tails
    const float* feat = y1;                           // ScaleHid is folded into out_w at pack time
    tile_gemm(feat, HID, RowPtr{net.out_w, 4}, 1.f, nullptr, 0, RowPtr{nullptr, 0}, 1,
              StoreT{q, net.out_b, false});           // [4][RS]: mean0, mean1, std0, std1 | value cols
    __syncthreads();
    if (threadIdx.x < hrows) {
      const int r = threadIdx.x;
      const int gl = r / nr, i = r - gl * nr;
      const int gi = tile0 + gl;
      const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
      if (net.kind == DGPPO_NET_POLICY) {
        policy_tail(g, q[0 * RS + r], q[1 * RS + r], q[2 * RS + r], q[3 * RS + r], env, slot, i, n);
      } else {
        float* vo = g.value + (((size_t)env * g.out_pitch + slot) * nr + i) * net.n_out;
        for (int c = 0; c < net.n_out; ++c) vo[c] = q[c * RS + r];
      }
    }
  }
}

static int fill_layout(const DgppoNetCfg* net, DgppoNetLayout* L) {
  if (!net || !L) return DGPPO_EINVAL;
  if (net->kind < 0 || net->kind > 2) return DGPPO_ENOTSUP;
  if (net->n_layers < 1 || net->n_layers > 2) return DGPPO_ENOTSUP;
  if (net->node_dim < 1 || net->node_dim > X0S || net->edge_dim != 4) return DGPPO_ENOTSUP;
  if (net->kind == DGPPO_NET_POLICY ? (net->n_out != 2) : (net->n_out < 1 || net->n_out > 4)) return DGPPO_ENOTSUP;
  int off = 0;
  auto take = [&](int nfl) { int o = off; off += round4(nfl); return o; };
  for (int l = 0; l < 2; ++l) {
    if (l >= net->n_layers) {
      L->wqk[l] = L->wagg[l] = L->wu[l] = L->bu[l] = -1;
      L->in_dim[l] = L->out_dim[l] = 0;
      continue;
    }
    const int IN = (l == 0) ? net->node_dim : 32;
    const int D = (l == net->n_layers - 1) ? 64 : 32;      // gnn.py:136
    L->in_dim[l] = IN; L->out_dim[l] = D;
    L->wqk[l] = take((IN + 1) * H * round4(IN + 1));
    L->wagg[l] = take(H * (IN + 5) * D);
    L->wu[l] = take(IN * D); L->bu[l] = take(D);
  }
  L->d0w = take(64 * 64); L->d0b = take(64); L->ln0s = take(64); L->ln0b = take(64);
  L->d1w = take(64 * 64); L->d1b = take(64); L->ln1s = take(64); L->ln1b = take(64);
  L->wi = take(64 * 192); L->bi = take(192); L->wh = take(64 * 192); L->bhn = take(64);
  L->out_w = take(64 * 4); L->out_b = take(4);
  for (int l = 0; l < 2; ++l) {                            // fallback-kernel blocks
    if (l >= net->n_layers) { L->wq[l] = L->bq[l] = L->wkt[l] = -1; continue; }
    const int IN = L->in_dim[l], D = L->out_dim[l];
    L->wq[l] = take(IN * H * D); L->bq[l] = take(H * D);
    L->wkt[l] = take(H * D * round4(IN + 1));
  }
  L->tc_head = take(TC_HEAD_FL);                           // tensor-core head operands, last
  L->total = off;
  return 0;
}

static int launch_gnn(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                      const float* params, GnnArgs g, int phase = 0) {
  DgppoNetLayout L;
  if (int rc = fill_layout(net, &L)) return rc;
  if (int rc = check_env_cfg(env)) return rc;
  const GraphDims d = graph_dims(*env);
  if (d.nd != net->node_dim) return DGPPO_EINVAL;
  if (d.n > R) return DGPPO_ENOTSUP;
  if (g.n_graphs == 0) return 0;
  g.n = d.n; g.N = d.N; g.E = d.E; g.nd = d.nd; g.n_ag = d.n_ag; g.n_ao = d.n_ao;
  if (g.c_agent) {                                    // graph-from-state mode: K3's constants (env_kernels.cu make_consts)
    g.g_nodes = d.g; g.sd = d.sd; g.c_lidar = is_lidar(env->kind) ? 1 : 0; g.c_paired = is_target(env->kind) ? 1 : 0;
    g.cR = (float)env->comm_radius; g.cR_diag = (float)(env->comm_radius + 1.0);
    g.cR_obs = (float)(env->comm_radius - 1e-1);
    g.cR_mpe_obs = (float)(is_tall_mpe(env->kind) ? env->comm_radius * 100 : env->comm_radius);
    if (d.n_on > 0 && !g.c_obs) return DGPPO_EINVAL;
    if (!g.c_goal) return DGPPO_EINVAL;
  }
  g.G = R / d.n;
  const int m_cap = g.G * (d.N - 1);

  NetP P;
  P.n_layers = net->n_layers; P.kind = net->kind; P.n_out = net->n_out;
  for (int l = 0; l < net->n_layers; ++l) {
    P.L[l] = LayerP{params + L.wq[l], params + L.bq[l], params + L.wkt[l], params + L.wagg[l],
                    params + L.wu[l], params + L.bu[l], params + L.wqk[l], L.in_dim[l], L.out_dim[l]};
  }
  if (net->n_layers == 1) P.L[1] = P.L[0];
  P.d0w = params + L.d0w; P.d0b = params + L.d0b; P.ln0s = params + L.ln0s; P.ln0b = params + L.ln0b;
  P.d1w = params + L.d1w; P.d1b = params + L.d1b; P.ln1s = params + L.ln1s; P.ln1b = params + L.ln1b;
  P.wi = params + L.wi; P.bi = params + L.bi; P.wh = params + L.wh; P.bhn = params + L.bhn;
  P.out_w = params + L.out_w; P.out_b = params + L.out_b;

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // v2 (weight-stationary pair of kernels) when the shape fits; v1 fused kernel otherwise
  const char* force_v1 = getenv("DGPPO_FORCE_V1");
  if (!(force_v1 && force_v1[0] == '1')) {
    const int rc2 = launch_gnn_v2(stream, P, L, params, g, sms, phase);
    if (rc2 != DGPPO_V2_UNSUPPORTED) return rc2;
  }
  if (phase != 0) return DGPPO_V2_UNSUPPORTED;      // the fused fallback kernel cannot be split
  if (g.c_agent) return DGPPO_ENOTSUP;              // ... nor build the graph from the state

  const size_t x0_fl = (size_t)m_cap * X0S > (size_t)HID * RS ? (size_t)m_cap * X0S : (size_t)HID * RS;
  const size_t fl = x0_fl + (net->n_layers == 2 ? (size_t)m_cap * X1S : 0) +
                    (size_t)(32 + 192 + H * 36 + HID) * RS;
  const size_t smem = fl * sizeof(float);
  if (smem > 227 * 1024) return DGPPO_ENOTSUP;
  const int n_tiles = (g.n_graphs + g.G - 1) / g.G;
  const int grid = n_tiles < sms ? n_tiles : sms;
  cudaError_t err;
  if (net->n_layers == 2) {
    err = cudaFuncSetAttribute(gnn_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    gnn_forward_kernel<2><<<grid, NT, smem, (cudaStream_t)stream>>>(P, g, m_cap);
  } else {
    err = cudaFuncSetAttribute(gnn_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    gnn_forward_kernel<1><<<grid, NT, smem, (cudaStream_t)stream>>>(P, g, m_cap);
  }
  return (int)cudaGetLastError();
}

}  // namespace dgppo

using namespace dgppo;

extern "C" int dgppo_net_layout(const DgppoNetCfg* net, DgppoNetLayout* out) {
  return fill_layout(net, out);
}

static void set_graph_source(GnnArgs& g, const float* nodes, const float* edges, const int32_t* receivers,
                             const int32_t* senders, const DgppoStateRecord* st, int32_t pitch) {
  g.nodes = nodes; g.edges = edges; g.recv = receivers; g.send = senders; g.pitch = pitch;
  g.c_agent = st ? st->agent : nullptr; g.c_obs = st ? st->obs_nodes : nullptr; g.c_goal = st ? st->goal : nullptr;
  g.c_pitch = pitch;
}

static int policy_impl(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net, const float* params,
                       const float* nodes, const float* edges, const int32_t* receivers, const int32_t* senders,
                       const DgppoStateRecord* st, int32_t pitch, const float* rnn_in, float* rnn_out,
                       int32_t rnn_pitch, const float* eps, int32_t eps_pitch, float* action, float* log_pi,
                       int32_t act_pitch, int32_t b) {
  if (!net || net->kind != DGPPO_NET_POLICY) return DGPPO_EINVAL;
  if (b < 0 || !params || !rnn_in || !rnn_out || !action) return DGPPO_EINVAL;
  if (st ? !st->agent : (!nodes || !edges || !receivers || !senders)) return DGPPO_EINVAL;
  if (pitch < 1 || rnn_pitch < 1 || act_pitch < 1 || (eps && eps_pitch < 1)) return DGPPO_EINVAL;
  GnnArgs g{};
  set_graph_source(g, nodes, edges, receivers, senders, st, pitch);
  g.n_slots = 1;
  g.rnn_in = rnn_in; g.rnn_out = rnn_out; g.rnn_pitch = rnn_pitch;
  g.eps = eps; g.eps_pitch = eps_pitch; g.action = action; g.log_pi = log_pi; g.act_pitch = act_pitch;
  g.value = nullptr; g.out_pitch = 1; g.n_graphs = b;
  return launch_gnn(stream, env, net, params, g);
}

static int value_impl(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net, const float* params,
                      const float* nodes, const float* edges, const int32_t* receivers, const int32_t* senders,
                      const DgppoStateRecord* st, int32_t pitch, const float* rnn_in, float* rnn_out,
                      int32_t rnn_pitch, float* value, int32_t out_pitch, int32_t n_slots, int32_t b) {
  if (!net || net->kind == DGPPO_NET_POLICY) return DGPPO_EINVAL;
  if (b < 0 || !params || !rnn_in || !value) return DGPPO_EINVAL;
  if (st ? !st->agent : (!nodes || !edges || !receivers || !senders)) return DGPPO_EINVAL;
  if (pitch < 1 || rnn_pitch < 1 || out_pitch < 1 || n_slots < 1 || n_slots > pitch) return DGPPO_EINVAL;
  GnnArgs g{};
  set_graph_source(g, nodes, edges, receivers, senders, st, pitch);
  g.n_slots = n_slots;
  g.rnn_in = rnn_in; g.rnn_out = rnn_out; g.rnn_pitch = rnn_pitch;
  g.eps = nullptr; g.eps_pitch = 1; g.action = nullptr; g.log_pi = nullptr; g.act_pitch = 1;
  g.value = value; g.out_pitch = out_pitch; g.n_graphs = b * n_slots;
  return launch_gnn(stream, env, net, params, g);
}

static int vl_scan_impl(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net, const float* params,
                        const float* nodes, const float* edges, const int32_t* receivers, const int32_t* senders,
                        const DgppoStateRecord* st, int32_t pitch, float* carry, int32_t carry_pitch, float* value,
                        int32_t out_pitch, int32_t n_slots, int32_t b) {
  if (!net || net->kind != DGPPO_NET_VL) return DGPPO_EINVAL;
  if (b < 0 || !params || !carry || !value) return DGPPO_EINVAL;
  if (st ? !st->agent : (!nodes || !edges || !receivers || !senders)) return DGPPO_EINVAL;
  if (pitch < 1 || out_pitch < 1 || n_slots < 1 || n_slots > pitch || n_slots > out_pitch ||
      carry_pitch < n_slots + 1) return DGPPO_EINVAL;
  if (b == 0) return 0;
  GnnArgs g{};
  set_graph_source(g, nodes, edges, receivers, senders, st, pitch);
  g.eps = nullptr; g.eps_pitch = 1; g.action = nullptr; g.log_pi = nullptr; g.act_pitch = 1;
  g.rnn_pitch = carry_pitch; g.out_pitch = out_pitch;
  // phase 1: the GNN part of every slot has no recurrence: one launch over all b * n_slots graphs,
  // embeddings parked in the carry rows they will be overwritten in (slot t -> carry slot t + 1)
  g.n_slots = n_slots; g.n_graphs = b * n_slots;
  g.rnn_in = carry; g.rnn_out = carry + HID; g.value = value;
  int rc = launch_gnn(stream, env, net, params, g, 1);
  if (rc == DGPPO_V2_UNSUPPORTED) {                   // shape outside the split kernels: slot by slot
    if (st) return DGPPO_ENOTSUP;
    const GraphDims d = graph_dims(*env);
    for (int t = 0; t < n_slots; ++t) {
      GnnArgs s = g;
      s.nodes = nodes + (size_t)t * d.N * d.nd;
      s.edges = edges + (size_t)t * d.E * 4;
      s.recv = receivers + (size_t)t * d.E; s.send = senders + (size_t)t * d.E;
      s.n_slots = 1; s.n_graphs = b;
      s.rnn_in = carry + (size_t)t * HID; s.rnn_out = carry + (size_t)(t + 1) * HID; s.value = value + t;
      if ((rc = launch_gnn(stream, env, net, params, s, 0))) return rc;
    }
    return 0;
  }
  if (rc) return rc;
  // phase 2: the recurrent head, slot by slot (carry t -> carry t + 1, value t)
  for (int t = 0; t < n_slots; ++t) {
    GnnArgs s = g;
    s.n_slots = 1; s.n_graphs = b;
    s.rnn_in = carry + (size_t)t * HID; s.rnn_out = carry + (size_t)(t + 1) * HID; s.value = value + t;
    if ((rc = launch_gnn(stream, env, net, params, s, 2))) return rc;
  }
  return 0;
}

extern "C" int dgppo_gnn_policy(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                                const float* params, const float* nodes, const float* edges,
                                const int32_t* receivers, const int32_t* senders, int32_t pitch,
                                const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                                const float* eps, int32_t eps_pitch, float* action, float* log_pi,
                                int32_t act_pitch, int32_t b) {
  return policy_impl(stream, env, net, params, nodes, edges, receivers, senders, nullptr, pitch, rnn_in, rnn_out,
                     rnn_pitch, eps, eps_pitch, action, log_pi, act_pitch, b);
}

extern "C" int dgppo_gnn_policy_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                                           const float* params, const DgppoStateRecord* st, int32_t pitch,
                                           const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                                           const float* eps, int32_t eps_pitch, float* action, float* log_pi,
                                           int32_t act_pitch, int32_t b) {
  if (!st) return DGPPO_EINVAL;
  return policy_impl(stream, env, net, params, nullptr, nullptr, nullptr, nullptr, st, pitch, rnn_in, rnn_out,
                     rnn_pitch, eps, eps_pitch, action, log_pi, act_pitch, b);
}

extern "C" int dgppo_gnn_value(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                               const float* params, const float* nodes, const float* edges,
                               const int32_t* receivers, const int32_t* senders, int32_t pitch,
                               const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                               float* value, int32_t out_pitch, int32_t n_slots, int32_t b) {
  return value_impl(stream, env, net, params, nodes, edges, receivers, senders, nullptr, pitch, rnn_in, rnn_out,
                    rnn_pitch, value, out_pitch, n_slots, b);
}

extern "C" int dgppo_gnn_value_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                                          const float* params, const DgppoStateRecord* st, int32_t pitch,
                                          const float* rnn_in, float* rnn_out, int32_t rnn_pitch,
                                          float* value, int32_t out_pitch, int32_t n_slots, int32_t b) {
  if (!st) return DGPPO_EINVAL;
  return value_impl(stream, env, net, params, nullptr, nullptr, nullptr, nullptr, st, pitch, rnn_in, rnn_out,
                    rnn_pitch, value, out_pitch, n_slots, b);
}

extern "C" int dgppo_vl_scan(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                             const float* params, const float* nodes, const float* edges,
                             const int32_t* receivers, const int32_t* senders, int32_t pitch,
                             float* carry, int32_t carry_pitch, float* value, int32_t out_pitch,
                             int32_t n_slots, int32_t b) {
  return vl_scan_impl(stream, env, net, params, nodes, edges, receivers, senders, nullptr, pitch, carry, carry_pitch,
                      value, out_pitch, n_slots, b);
}

extern "C" int dgppo_vl_scan_from_state(void* stream, const DgppoEnvCfg* env, const DgppoNetCfg* net,
                                        const float* params, const DgppoStateRecord* st, int32_t pitch,
                                        float* carry, int32_t carry_pitch, float* value, int32_t out_pitch,
                                        int32_t n_slots, int32_t b) {
  if (!st) return DGPPO_EINVAL;
  return vl_scan_impl(stream, env, net, params, nullptr, nullptr, nullptr, nullptr, st, pitch, carry, carry_pitch,
                      value, out_pitch, n_slots, b);
}
