// K4 v2: weight-stationary GraphTransformer forward for sm_100a (FP32 FFMA).
//
// The v1 kernel (gnn_kernels.cu) streams every weight through L1/L2 inside the
// GEMM k-loops and is latency-bound on those loads (profiles/r1_policy_v1.ncu.txt:
// 40% long-scoreboard stalls, 15% FMA-pipe utilisation).  v2 splits the forward
// into two persistent kernels whose weights are loaded ONCE per CTA into shared
// memory and stay there for every tile:
//
//   gnn_layers_kernel     : GNN layers for a tile of G graphs (G*n <= 16 agent rows),
//                           256 threads, 2 CTAs per SM; smem = GNN weights (~57 KB) +
//                           node/row activations.  Output: agent embeddings (rows x 64)
//                           into the caller's rnn_out buffer, used as scratch.
//   gnn_layers_big_kernel : the same for n > 16: one graph per tile, rows in chunks of 16,
//                           attention over compacted lists of live slots.
//   head_kernel_wide      : head MLP + LayerNorm + GRU + tails for a tile of 128 rows
//                           (head_kernel<WR>: 64 rows); smem = head/GRU/tail weights
//                           (~148 KB) + activations.  Reads the embeddings back from
//                           rnn_out, then overwrites them with the new GRU carry.
//
// Same arithmetic as v1 (GNN regrouping); attention is parallelised
// over (row, head, edge slot) instead of (row, head).
#include <stdlib.h>
#include <string.h>

#include "gnn_common.cuh"

namespace dgppo {

constexpr int R2 = 16;           // agent rows per GNN tile (two CTAs per SM)
constexpr int RS2 = 16;          // row stride of the GNN tile's transposed buffers
constexpr int QTS = 36;          // per-(row, head) stride of the regrouped keys

__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Register-tiled GEMM over a tile of ROWS rows: thread tile RT rows x 4 cols,
// A operands transposed [k][row] (stride AS) in smem, W row-major [k][ld] in smem.
// A warp covers 8 row groups x 4 col groups.  Warps [w0, w0 + nw) of the CTA take
// part; the reduction runs over k in [k_lo, k_hi) of source 1 (+ all of source 2
// when K2 > 0):  acc = scale1 * A1 W1 + A2 W2.
template <int ROWS, int AS, int RT, class Epi>
__device__ __forceinline__ void gemm_ws(const float* A1, int k_lo, int k_hi, const float* W1, int ld1, float scale1,
                                        const float* A2, int K2, const float* W2, int ld2,
                                        int N4, int w0, int nw, Epi epi) {
  constexpr int NRG = ROWS / RT;             // row groups
  constexpr int NRGB = NRG / 8;              // blocks of 8 row groups
  const int ncgb = (N4 + 3) >> 2;
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) - w0;
  if (warp < 0 || warp >= nw) return;
  for (int bi = warp; bi < NRGB * ncgb; bi += nw) {
    const int rgb = bi % NRGB, cgb = bi / NRGB;
    const int rg = rgb * 8 + (lane & 7), cg = cgb * 4 + (lane >> 3);
    if (cg >= N4) continue;
    // packed FP32 FMA (FFMA2): accumulators are (col 2q, col 2q+1) pairs, the W float4 supplies the
    // pairs directly and only the RT row values are duplicated
    float2 acc2[RT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i) { acc2[i][0] = make_float2(0.f, 0.f); acc2[i][1] = make_float2(0.f, 0.f); }
    auto run = [&](const float* A, int ka, int kb, const float* W, int ld) {
      const float* ap = A + rg * RT;
      const float* wp = W + cg * 4;
#pragma unroll 8
      for (int k = ka; k < kb; ++k) {
        float av[RT];
        if constexpr (RT == 4) {
          const float4 a = *reinterpret_cast<const float4*>(ap + k * AS);
          av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
        } else {
          const float2 a = *reinterpret_cast<const float2*>(ap + k * AS);
          av[0] = a.x; av[1] = a.y;
        }
        const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)k * ld);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const float2 ad = make_float2(av[i], av[i]);
          acc2[i][0] = __ffma2_rn(ad, w01, acc2[i][0]);
          acc2[i][1] = __ffma2_rn(ad, w23, acc2[i][1]);
        }
      }
    };
    run(A1, k_lo, k_hi, W1, ld1);
    float acc[RT][4];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      acc[i][0] = acc2[i][0].x; acc[i][1] = acc2[i][0].y; acc[i][2] = acc2[i][1].x; acc[i][3] = acc2[i][1].y;
    }
    if (scale1 != 1.f) {
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] *= scale1;
    }
    if (K2 > 0) {
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        acc2[i][0] = make_float2(acc[i][0], acc[i][1]); acc2[i][1] = make_float2(acc[i][2], acc[i][3]);
      }
      run(A2, 0, K2, W2, ld2);
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        acc[i][0] = acc2[i][0].x; acc[i][1] = acc2[i][0].y; acc[i][2] = acc2[i][1].x; acc[i][3] = acc2[i][1].y;
      }
    }
    epi(rg * RT, cg * 4, acc);
  }
}

struct GnnV2Plan {
  int w_off, w_fl;        // GNN weights: [w_off, w_off + w_fl) floats of the packed buffer
  int m_cap, deg, degp;   // degp: slots per row padded to a multiple of 32
  int x0_fl, x1_fl, sidx_fl, scr_fl;
  int threads;            // CTA size of gnn_layers_kernel (512 or 256)
  size_t smem_bytes;
  int hw_off, hw_fl;      // head weights
  size_t head_smem_bytes;
};

// exp of a non-positive softmax argument: ex2.approx(x * log2 e).  Relative error <= 2^-22 + |x| 2^-24,
// i.e. below 1e-6 wherever exp(x) still matters; the weights that dominate the sum have x near 0.
__device__ __forceinline__ float softmax_exp(float x) { return __expf(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Attention of ONE receiver row by ONE warp (gnn.py:100-107,114 regrouped):
//   lane t scores slot t for the 3 heads -> warp softmax -> the ACTIVE slots are
//   compacted (ballot) into the warp's scratch as (a_0, a_1, a_2, offset of x_s)
//   -> lanes turn into feature columns and accumulate the weighted sender features
//   over the compacted list; 12 lanes accumulate the weighted edge features.
// Writes column r of z[(h*INA + c)][r].
template <int INX, int J>
__device__ __forceinline__ void attention_row(int r, bool live, int lane, const int* srow, int deg,
                                              const float* qrow /* qt + r*H*QTS */, int IN, float isd,
                                              const float* X, int XS, const float4* ed, int i_agent,
                                              int n, int n_ag, int n_ao, float* scr, float* z,
                                              const GnnArgs* cg = nullptr, const float* x0g = nullptr, int node0 = 0) {
  const int INA = IN + 5;
  float* zc = z + r;
  if (!live) {                                   // rows past the tile's graphs: zero column
    for (int c = lane; c < H * INA; c += 32) zc[c * RS2] = 0.f;
    return;
  }
  float4* al = reinterpret_cast<float4*>(scr);                  // [count]: (a_0, a_1, a_2, bits(s * XS))
  float4* eft = reinterpret_cast<float4*>(scr + J * 32 * 4);    // [count]: edge features of the slot
  int sv[J];
  float sc[J][H];
  float4 ef[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int t = lane + 32 * j;
    sv[j] = (t < deg) ? srow[t] : -1;
#pragma unroll
    for (int h = 0; h < H; ++h) sc[j][h] = -INFINITY;
    ef[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sv[j] >= 0) {
      const float* xp = X + (size_t)sv[j] * XS;
      float acc[H], acd[H];                       // two partial sums per head: half the dependent-FMA depth
#pragma unroll
      for (int h = 0; h < H; ++h) { acc[h] = qrow[h * QTS + IN]; acd[h] = 0.f; }
#pragma unroll
      for (int c = 0; c < INX; c += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + c);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 qv = *reinterpret_cast<const float4*>(qrow + h * QTS + c);
          acc[h] = fmaf(qv.x, xv.x, acc[h]); acd[h] = fmaf(qv.y, xv.y, acd[h]);
          acc[h] = fmaf(qv.z, xv.z, acc[h]); acd[h] = fmaf(qv.w, xv.w, acd[h]);
        }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) sc[j][h] = (acc[h] + acd[h]) * isd;
      const int e = (t < n) ? i_agent * n + t
                            : ((t < n + n_ag) ? n * n + i_agent * n_ag + (t - n)
                                              : n * n + n * n_ag + i_agent * n_ao + (t - n - n_ag));
      // edge features: from the record, or (graph-from-state mode) from the staged node rows with K3's arithmetic
      ef[j] = cg ? edge_from_state(*cg, x0g + i_agent * X0S, x0g + (sv[j] - node0) * X0S, t) : __ldg(ed + e);
    }
  }
  {                                               // jraph.segment_softmax over the row's slots;
    float mx[H], l[H];                            // the three heads' shuffle chains run interleaved
#pragma unroll
    for (int h = 0; h < H; ++h) {
      mx[h] = sc[0][h];
#pragma unroll
      for (int j = 1; j < J; ++j) mx[h] = fmaxf(mx[h], sc[j][h]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      l[h] = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) { sc[j][h] = (sv[j] >= 0) ? softmax_exp(sc[j][h] - mx[h]) : 0.f; l[h] += sc[j][h]; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] += __shfl_xor_sync(0xffffffffu, l[h], o);
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float inv_l = (l[h] > 0.f) ? __fdividef(1.f, l[h]) : 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) sc[j][h] *= inv_l;
    }
  }
  int count = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {                   // compact the active slots
    const unsigned m = __ballot_sync(0xffffffffu, sv[j] >= 0);
    if (sv[j] >= 0) {
      const int pos = count + __popc(m & ((1u << lane) - 1u));
      al[pos] = make_float4(sc[j][0], sc[j][1], sc[j][2], __int_as_float(sv[j] * XS));
      eft[pos] = ef[j];
    }
    count += __popc(m);
  }
  __syncwarp();
  {
    constexpr int NG = 32 / INX;                  // slot groups sharing the warp (4 for INX = 8)
    const int c = lane % INX, tq = lane / INX;
    const int eh = (lane >> 2) % H, ej = lane & 3;    // lanes < 12: (head, edge feature)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, ae = 0.f;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f, be = 0.f;   // second set of partial sums: two list entries in flight
    int t = tq;
    for (; t + NG < count; t += 2 * NG) {
      const float4 av = al[t], bv = al[t + NG];
      const float x = X[__float_as_int(av.w) + c], y = X[__float_as_int(bv.w) + c];
      a0 = fmaf(av.x, x, a0); a1 = fmaf(av.y, x, a1); a2 = fmaf(av.z, x, a2);
      b0 = fmaf(bv.x, y, b0); b1 = fmaf(bv.y, y, b1); b2 = fmaf(bv.z, y, b2);
      if (NG == 1) {
        ae = fmaf(scr[t * 4 + eh], scr[J * 32 * 4 + t * 4 + ej], ae);
        be = fmaf(scr[(t + 1) * 4 + eh], scr[J * 32 * 4 + (t + 1) * 4 + ej], be);
      }
    }
    if (t < count) {
      const float4 av = al[t];
      const float x = X[__float_as_int(av.w) + c];
      a0 = fmaf(av.x, x, a0); a1 = fmaf(av.y, x, a1); a2 = fmaf(av.z, x, a2);
      if (NG == 1) ae = fmaf(scr[t * 4 + eh], scr[J * 32 * 4 + t * 4 + ej], ae);
    }
    a0 += b0; a1 += b1; a2 += b2; ae += be;
    if (NG > 1) {
#pragma unroll
      for (int o = INX; o < 32; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (lane < H * 4)
        for (int t = 0; t < count; ++t) ae = fmaf(scr[t * 4 + eh], scr[J * 32 * 4 + t * 4 + ej], ae);
    }
    if (tq == 0 && c < IN) {
      zc[c * RS2] = a0; zc[(INA + c) * RS2] = a1; zc[(2 * INA + c) * RS2] = a2;
    }
    if (lane < H * 4) {
      zc[(eh * INA + IN + 1 + ej) * RS2] = ae;
      if (ej == 0) zc[(eh * INA + IN) * RS2] = (count > 0) ? 1.f : 0.f;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------ GNN layers
template <int NL>
__global__ void __launch_bounds__(256, 2)
gnn_layers_kernel(NetP net, GnnArgs g, GnnV2Plan pl, const float* __restrict__ params) {
  extern __shared__ __align__(16) float smem[];
  const int nth = blockDim.x, nwarps = nth >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ws = smem;                                    // GNN weights
  float* x0 = ws + pl.w_fl;                            // [M][8]
  float* x1 = x0 + pl.x0_fl;                           // [M][36]
  float* xr = x1 + pl.x1_fl;                           // [32][RS2]
  float* qt = xr + 32 * RS2;                           // [R2][H][QTS]; split-K partial after the attention
  float* z = qt + R2 * H * QTS;                        // [112][RS2]
  int* sidx = reinterpret_cast<int*>(z + 112 * RS2);   // [R2][degp]
  float* scr = reinterpret_cast<float*>(sidx) + pl.sidx_fl;   // per-warp attention scratch
  int* tab = reinterpret_cast<int*>(scr + pl.scr_fl);  // [0,32) row->graph, [32,64) row->agent, [64,96) gslot, [96,128) rslot
  unsigned char* nflag = reinterpret_cast<unsigned char*>(tab + 128);   // [m_cap] node is an active sender

  for (int i = threadIdx.x; i < pl.w_fl / 4; i += nth) cp_async16(ws + 4 * i, params + pl.w_off + 4 * i);
  cp_async_wait_all();
  pdl_wait();                     // weights staged while the predecessor (graph kernel) drains; now its output is visible
  pdl_launch_dependents();        // the head kernel may be scheduled as SMs free up: it stages its weights meanwhile
  __syncthreads();
  auto wptr = [&](const float* p) { return ws + ((p - params) - pl.w_off); };

  const int n = g.n, N = g.N, nd = g.nd, G = g.G, deg = pl.deg, degp = pl.degp;
  const int dshift = (degp == 32) ? 5 : 6;
  const int nodes_per = N - 1, pad = N - 1;
  const int n_tiles = (g.n_graphs + G - 1) / G;
  const int nr_out = (net.kind == DGPPO_NET_VL) ? 1 : n;
  float* my_scr = scr + (size_t)warp * degp * 8;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tile0 = tile * G;
    const int gcount = min(G, g.n_graphs - tile0);
    const int rows = gcount * n;
    const int M = gcount * nodes_per;
    __syncthreads();
    // ---- per-tile index tables (all later phases read these instead of dividing)
    if (threadIdx.x < R2) {
      const int r = threadIdx.x;
      tab[r] = (r < rows) ? r / n : -1;
      tab[32 + r] = (r < rows) ? r % n : 0;
      if (r < gcount) {
        const int gi = tile0 + r;
        const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
        tab[64 + r] = env * g.pitch + slot;
        tab[96 + r] = env * g.rnn_pitch + slot;
        tab[64 + R2 + r] = env;                      // graph-from-state mode: (env, slot) of the tile's graphs
        tab[96 + R2 + r] = slot;
      }
    }
    for (int i = threadIdx.x; i < (M + 3) / 4; i += nth) reinterpret_cast<int*>(nflag)[i] = 0;
    __syncthreads();

    const bool from_state = g.c_agent != nullptr;
    if (from_state) {
      // ---- graph from state: node rows with K3's layout, then slot masks / senders from the staged rows
      for (int idx = threadIdx.x; idx < gcount * nodes_per; idx += nth) {
        const int gl = idx / nodes_per, row = idx - gl * nodes_per;
        node_row_from_state(g, tab[64 + R2 + gl], tab[96 + R2 + gl], row, x0 + (size_t)idx * X0S);
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < R2 * degp; idx += nth) {
        const int r = idx >> dshift, t = idx & (degp - 1);
        int s = -1;
        const int gl = tab[r];
        if (gl >= 0 && t < deg) {
          bool live;
          const int sd = slot_sender_from_state(g, x0 + (size_t)gl * nodes_per * X0S, tab[32 + r], t, live);
          if (live) { s = gl * nodes_per + sd; nflag[s] = 1; }
        }
        sidx[idx] = s;
      }
      for (int idx = threadIdx.x; idx < 32 * R2; idx += nth) {       // layer 0: xr[c][r] = node row of agent(r)
        const int c = idx / R2, r = idx % R2;
        const int gl = tab[r];
        xr[c * RS2 + r] = (gl >= 0 && c < net.L[0].in) ? x0[((size_t)gl * nodes_per + tab[32 + r]) * X0S + c] : 0.f;
      }
    } else {
    // ---- stage node features (async) and the per-row sender table
    for (int gl = 0; gl < gcount; ++gl) {
      const float* src = g.nodes + (size_t)tab[64 + gl] * N * nd;
      float* dst = x0 + (size_t)gl * nodes_per * X0S;
      if (nd == X0S) {
        for (int j = threadIdx.x; j < nodes_per * X0S; j += nth) cp_async4(dst + j, src + j);
      } else {
        for (int j = threadIdx.x; j < nodes_per * 7; j += nth) {
          const int node = j / 7, c = j - node * 7;
          cp_async4(dst + node * X0S + c, src + j);
        }
        for (int j = threadIdx.x; j < nodes_per; j += nth) dst[j * X0S + 7] = 0.f;
      }
    }
    for (int idx = threadIdx.x; idx < R2 * degp; idx += nth) {
      const int r = idx >> dshift, t = idx & (degp - 1);  // degp is 32 or 64
      int s = -1;
      const int gl = tab[r];
      if (gl >= 0 && t < deg) {
        const int i = tab[32 + r];
        const size_t gslot = (size_t)tab[64 + gl];
        const int e = (t < n) ? i * n + t
                              : ((t < n + g.n_ag) ? n * n + i * g.n_ag + (t - n)
                                                  : n * n + n * g.n_ag + i * g.n_ao + (t - n - g.n_ag));
        const int rv = __ldg(g.recv + gslot * g.E + e), sd = __ldg(g.send + gslot * g.E + e);   // independent loads
        if (rv != pad) {
          s = gl * nodes_per + sd;
          nflag[s] = 1;
        }
      }
      sidx[idx] = s;
    }
    for (int idx = threadIdx.x; idx < 32 * R2; idx += nth) {       // layer 0: xr[c][r] = nodes[agent(r)][c]
      const int c = idx / R2, r = idx % R2;
      const int gl = tab[r];
      float v = 0.f;
      if (gl >= 0 && c < net.L[0].in)
        v = __ldg(g.nodes + ((size_t)tab[64 + gl] * N + tab[32 + r]) * nd + c);
      xr[c * RS2 + r] = v;
    }
    cp_async_wait_all();
    }
    __syncthreads();

    const float* X = x0; int XS = X0S;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const LayerP& P = net.L[l];
      const int IN = P.in, D = P.d, INP = round4(IN + 1), INA = IN + 5;
      const bool last = (l == NL - 1);
      const float *wqk = wptr(P.wqk), *wagg = wptr(P.wagg), *wu = wptr(P.wu), *bu = wptr(P.bu);

      // xr[c][r] = X[node(r)][c]: filled from global during staging (layer 0) / by the previous
      // layer's combine step (layer 1)
      // qt[r][h][c] = x_r (Wq_h Wk_h^T)[:, c] + bias row  (query and key merged at pack time)
      gemm_ws<R2, RS2, 2>(xr, 0, IN, wqk, H * INP, 1.f, nullptr, 0, nullptr, 0, H * INP / 4, 0, nwarps,
                          [&](int r0, int c0, float (&acc)[2][4]) {
                            const int h = c0 / INP, c = c0 - h * INP;
                            const float4 bb = *reinterpret_cast<const float4*>(wqk + IN * H * INP + c0);
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                              *reinterpret_cast<float4*>(qt + ((r0 + i) * H + h) * QTS + c) =
                                  make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
                          });
      __syncthreads();
      // ---- attention: one warp per receiver row
      {
        const float isd = 1.f / sqrtf((float)D);
        for (int r = warp; r < R2; r += nwarps) {
          const int gl = tab[r];
          const bool live = gl >= 0;
          const int ia = tab[32 + r];
          const float4* ed = reinterpret_cast<const float4*>(g.edges + (size_t)tab[64 + (live ? gl : 0)] * g.E * 4);
          const GnnArgs* cgp = from_state ? &g : nullptr;
          const int node0 = (live ? gl : 0) * nodes_per;
          const float* x0g = x0 + (size_t)node0 * X0S;
          if (l == 0) {
            if (degp == 32) attention_row<X0S, 1>(r, live, lane, sidx + r * degp, deg, qt + r * H * QTS, IN, isd,
                                                  X, XS, ed, ia, n, g.n_ag, g.n_ao, my_scr, z, cgp, x0g, node0);
            else            attention_row<X0S, 2>(r, live, lane, sidx + r * degp, deg, qt + r * H * QTS, IN, isd,
                                                  X, XS, ed, ia, n, g.n_ag, g.n_ao, my_scr, z, cgp, x0g, node0);
          } else {
            if (degp == 32) attention_row<32, 1>(r, live, lane, sidx + r * degp, deg, qt + r * H * QTS, IN, isd,
                                                 X, XS, ed, ia, n, g.n_ag, g.n_ao, my_scr, z, cgp, x0g, node0);
            else            attention_row<32, 2>(r, live, lane, sidx + r * degp, deg, qt + r * H * QTS, IN, isd,
                                                 X, XS, ed, ia, n, g.n_ag, g.n_ao, my_scr, z, cgp, x0g, node0);
          }
        }
      }
      __syncthreads();
      // ---- x' = relu(x Wu + bu + 1/H z Wagg): split-K, the two halves of the CTA run
      //      concurrently and leave partial sums in smem (scr / qt are dead by now):
      //      upper warps: k in [kh, H*INA) of z Wagg; lower warps: k in [0, kh) + x Wu.
      const int KZ = H * INA, kh = KZ / 2, hw = nwarps / 2;
      const int qshift = (D == 32) ? 3 : 4;
      float* part_hi = qt;                                            // [R2][64]
      float* part_lo = scr;                                           // [R2][64]
      auto store_part = [&](float* dstp) {
        return [=](int r0, int c0, float (&acc)[2][4]) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
            *reinterpret_cast<float4*>(dstp + (r0 + i) * 64 + c0) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        };
      };
      gemm_ws<R2, RS2, 2>(z, kh, KZ, wagg, D, 1.f / H, nullptr, 0, nullptr, 0, D / 4, hw, nwarps - hw,
                          store_part(part_hi));
      gemm_ws<R2, RS2, 2>(z, 0, kh, wagg, D, 1.f / H, xr, IN, wu, D, D / 4, 0, hw, store_part(part_lo));
      if (!last) {                          // non-agent nodes that send in the next layer (all threads)
        const int nn = nodes_per - n;
        for (int gl = 0; gl < gcount; ++gl) {
          // one (node, 8 output columns) item per thread; layer 0 only: X = x0, rows of X0S = 8 floats
          // (a non-final layer is 32 wide: gnn.py:136, dgppo_net_layout)
          for (int idx = threadIdx.x; idx < nn * 4; idx += nth) {
            const int s = gl * nodes_per + n + (idx >> 2), c0 = (idx & 3) * 8;
            if (!nflag[s]) continue;
            const float4 xa = *reinterpret_cast<const float4*>(x0 + (size_t)s * X0S);
            const float4 xb = *reinterpret_cast<const float4*>(x0 + (size_t)s * X0S + 4);
            const float xs[X0S] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            float2 acc[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[p] = make_float2(bu[c0 + 2 * p], bu[c0 + 2 * p + 1]);
#pragma unroll
            for (int c = 0; c < X0S; ++c) {
              if (c < IN) {
                const float4 w0 = *reinterpret_cast<const float4*>(wu + c * D + c0);
                const float4 w1 = *reinterpret_cast<const float4*>(wu + c * D + c0 + 4);
                const float2 xd = make_float2(xs[c], xs[c]);
                acc[0] = __ffma2_rn(xd, make_float2(w0.x, w0.y), acc[0]);
                acc[1] = __ffma2_rn(xd, make_float2(w0.z, w0.w), acc[1]);
                acc[2] = __ffma2_rn(xd, make_float2(w1.x, w1.y), acc[2]);
                acc[3] = __ffma2_rn(xd, make_float2(w1.z, w1.w), acc[3]);
              }
            }
            float* dst = x1 + (size_t)s * X1S + c0;
            *reinterpret_cast<float4*>(dst) =
                make_float4(fmaxf(acc[0].x, 0.f), fmaxf(acc[0].y, 0.f), fmaxf(acc[1].x, 0.f), fmaxf(acc[1].y, 0.f));
            *reinterpret_cast<float4*>(dst + 4) =
                make_float4(fmaxf(acc[2].x, 0.f), fmaxf(acc[2].y, 0.f), fmaxf(acc[3].x, 0.f), fmaxf(acc[3].y, 0.f));
          }
        }
      }
      __syncthreads();
      float* ob = x0;                                                 // Vl only: [64][RS2] (x0 is dead by then)
      for (int idx = threadIdx.x; idx < R2 * (D / 4); idx += nth) {   // combine + bias + relu
        const int r = idx >> qshift, c0 = (idx & ((D >> 2) - 1)) * 4;   // D / 4 is 8 or 16
        const float4 lo = *reinterpret_cast<const float4*>(part_lo + r * 64 + c0);
        const float4 hi = *reinterpret_cast<const float4*>(part_hi + r * 64 + c0);
        const float4 v = make_float4(fmaxf(lo.x + hi.x + bu[c0], 0.f), fmaxf(lo.y + hi.y + bu[c0 + 1], 0.f),
                                     fmaxf(lo.z + hi.z + bu[c0 + 2], 0.f), fmaxf(lo.w + hi.w + bu[c0 + 3], 0.f));
        const int gl = tab[r];
        if (last && net.kind == DGPPO_NET_VL) {
          ob[(c0 + 0) * RS2 + r] = v.x; ob[(c0 + 1) * RS2 + r] = v.y;
          ob[(c0 + 2) * RS2 + r] = v.z; ob[(c0 + 3) * RS2 + r] = v.w;
        } else if (gl >= 0) {
          const int ia = tab[32 + r];
          if (!last) *reinterpret_cast<float4*>(x1 + ((size_t)gl * nodes_per + ia) * X1S + c0) = v;
          else       *reinterpret_cast<float4*>(g.rnn_out + ((size_t)tab[96 + gl] * n + ia) * HID + c0) = v;   // scratch rows
        }
        if (!last) {                                                  // receiver features of the next layer (D == 32)
          const float4 w = (gl >= 0) ? v : make_float4(0.f, 0.f, 0.f, 0.f);
          xr[(c0 + 0) * RS2 + r] = w.x; xr[(c0 + 1) * RS2 + r] = w.y;
          xr[(c0 + 2) * RS2 + r] = w.z; xr[(c0 + 3) * RS2 + r] = w.w;
        }
      }
      if (!last) { X = x1; XS = X1S; }
      if (last && net.kind == DGPPO_NET_VL) {
        // centralised Vl: mean over the agents of each graph (value.py:28-31)
        __syncthreads();
        for (int idx = threadIdx.x; idx < HID * gcount; idx += nth) {
          const int gl = idx / HID, c = idx - gl * HID;
          float sacc = 0.f;
          for (int i = 0; i < n; ++i) sacc += ob[c * RS2 + gl * n + i];
          g.rnn_out[((size_t)tab[96 + gl] * nr_out) * HID + c] = sacc / (float)n;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------- GNN layers, large graphs
// n > R2 (e.g. LidarSpread n = 64): ONE graph per tile, its node features
// resident in smem, receiver rows processed in chunks of R2.  Attention works on
// the COMPACTED list of live slots (a row of C5 has 136 slots, about half of
// them masked): lanes first compact (slot, sender) pairs, then score the live
// entries 32 at a time, keep exp(score - max) unnormalised in the list and fold
// 1/sum into the column sums at the end.  Scratch per warp: degp float4.

// 16 per-lane values -> warp sums, 16 shuffles (transpose-reduce).  On return
// v[0] of lane L is the warp sum of value k(L) = the bits (L>>1) of L reversed:
//   k = 8*(L>>4 & 1) + 4*(L>>3 & 1) + 2*(L>>2 & 1) + (L>>1 & 1).
__device__ __forceinline__ void warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int INX>
__device__ __forceinline__ void attention_row_big(int r, bool live, int lane, const unsigned short* srow, int deg, int degp,
                                                  const float* qrow, int IN, float isd,
                                                  const float* X, int XS, const float4* ed, int i_agent,
                                                  int n, int n_ag, int n_ao, float4* al, float* z,
                                                  const GnnArgs* cg = nullptr, const float* x0g = nullptr) {
  const int INA = IN + 5;
  float* zc = z + r;
  if (!live) {
    for (int c = lane; c < H * INA; c += 32) zc[c * RS2] = 0.f;
    return;
  }
  // ---- compact the live slots: al[pos].w = (slot << 16) | (sender * XS)
  int count = 0;
  for (int t0 = 0; t0 < degp; t0 += 32) {
    const int t = t0 + lane;
    const int s = (t < deg) ? (int)srow[t] : 0xffff;             // 0xffff: masked slot
    const unsigned m = __ballot_sync(0xffffffffu, s != 0xffff);
    if (s != 0xffff) al[count + __popc(m & ((1u << lane) - 1u))].w = __int_as_float((t << 16) | (s * XS));
    count += __popc(m);
  }
  __syncwarp();
  // ---- scores of the live entries, running max per head
  float mx[H];
#pragma unroll
  for (int h = 0; h < H; ++h) mx[h] = -INFINITY;
  for (int p0 = 0; p0 < count; p0 += 32) {
    const int p = p0 + lane;
    if (p < count) {
      const int pk = __float_as_int(al[p].w);
      const float* xp = X + (pk & 0xffff);
      float acc[H];
#pragma unroll
      for (int h = 0; h < H; ++h) acc[h] = qrow[h * QTS + IN];
#pragma unroll
      for (int c = 0; c < INX; c += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + c);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 qv = *reinterpret_cast<const float4*>(qrow + h * QTS + c);
          acc[h] = fmaf(qv.x, xv.x, acc[h]); acc[h] = fmaf(qv.y, xv.y, acc[h]);
          acc[h] = fmaf(qv.z, xv.z, acc[h]); acc[h] = fmaf(qv.w, xv.w, acc[h]);
        }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) { acc[h] *= isd; mx[h] = fmaxf(mx[h], acc[h]); }
      al[p].x = acc[0]; al[p].y = acc[1]; al[p].z = acc[2];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));
  }
  // ---- exp, row sums, weighted edge features (unnormalised)
  float l[H] = {0.f, 0.f, 0.f};
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  auto edge_of = [&](int p) -> float4 {             // edge features of list entry p (global, L2)
    if (p >= count) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int t = __float_as_int(al[p].w) >> 16;
    if (cg)         // graph-from-state mode: K3's expressions on the staged node rows (al[p].w low half = sender * XS)
      return edge_from_state(*cg, x0g + i_agent * X0S, x0g + ((__float_as_int(al[p].w) & 0xffff) / XS) * X0S, t);
    const int e = (t < n) ? i_agent * n + t
                          : ((t < n + n_ag) ? n * n + i_agent * n_ag + (t - n)
                                            : n * n + n * n_ag + i_agent * n_ao + (t - n - n_ag));
    return __ldg(ed + e);
  };
  float4 ef_next = edge_of(lane);
  for (int p0 = 0; p0 < count; p0 += 32) {
    const int p = p0 + lane;
    const float4 ef = ef_next;
    ef_next = edge_of(p + 32);                        // in flight while this chunk is processed
    if (p < count) {
      float4 a = al[p];
      a.x = softmax_exp(a.x - mx[0]); a.y = softmax_exp(a.y - mx[1]); a.z = softmax_exp(a.z - mx[2]);
      al[p] = a;
      l[0] += a.x; l[1] += a.y; l[2] += a.z;
      v[0] = fmaf(a.x, ef.x, v[0]); v[1] = fmaf(a.x, ef.y, v[1]); v[2] = fmaf(a.x, ef.z, v[2]); v[3] = fmaf(a.x, ef.w, v[3]);
      v[4] = fmaf(a.y, ef.x, v[4]); v[5] = fmaf(a.y, ef.y, v[5]); v[6] = fmaf(a.y, ef.z, v[6]); v[7] = fmaf(a.y, ef.w, v[7]);
      v[8] = fmaf(a.z, ef.x, v[8]); v[9] = fmaf(a.z, ef.y, v[9]); v[10] = fmaf(a.z, ef.z, v[10]); v[11] = fmaf(a.z, ef.w, v[11]);
    }
  }
  float inv_l[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { l[h] = warp_sum(l[h]); inv_l[h] = (l[h] > 0.f) ? __fdividef(1.f, l[h]) : 0.f; }
  warp_reduce16(v, lane);
  __syncwarp();
  {
    const int k = 8 * ((lane >> 4) & 1) + 4 * ((lane >> 3) & 1) + 2 * ((lane >> 2) & 1) + ((lane >> 1) & 1);
    if (!(lane & 1) && k < H * 4) {
      const int eh = k >> 2, ej = k & 3;
      const float il = (eh == 0) ? inv_l[0] : ((eh == 1) ? inv_l[1] : inv_l[2]);
      zc[(eh * INA + IN + 1 + ej) * RS2] = v[0] * il;
      if (ej == 0) zc[(eh * INA + IN) * RS2] = (count > 0) ? 1.f : 0.f;
    }
  }
  // ---- weighted sender features: lanes = feature columns (x NG slot groups)
  {
    constexpr int NG = 32 / INX;
    const int c = lane % INX, tq = lane / INX;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int t = tq; t < count; t += NG) {
      const float4 av = al[t];
      const float x = X[(__float_as_int(av.w) & 0xffff) + c];
      a0 = fmaf(av.x, x, a0); a1 = fmaf(av.y, x, a1); a2 = fmaf(av.z, x, a2);
    }
    if (NG > 1) {
#pragma unroll
      for (int o = INX; o < 32; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
    }
    if (tq == 0 && c < IN) {
      zc[c * RS2] = a0 * inv_l[0]; zc[(INA + c) * RS2] = a1 * inv_l[1]; zc[(2 * INA + c) * RS2] = a2 * inv_l[2];
    }
  }
  __syncwarp();
}

template <int NL>
__global__ void __launch_bounds__(256, 1)
gnn_layers_big_kernel(NetP net, GnnArgs g, GnnV2Plan pl, const float* __restrict__ params) {
  extern __shared__ __align__(16) float smem[];
  const int nth = blockDim.x, nwarps = nth >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ws = smem;                                    // GNN weights
  float* x0 = ws + pl.w_fl;                            // [M][8]
  float* x1 = x0 + pl.x0_fl;                           // [M][36]
  float* xr = x1 + pl.x1_fl;                           // [32][RS2]
  float* qt = xr + 32 * RS2;                           // [R2][H][QTS]; split-K partial after the attention
  float* z = qt + R2 * H * QTS;                        // [112][RS2]; Vl: last-layer rows after the GEMM
  unsigned short* sidx = reinterpret_cast<unsigned short*>(z + 112 * RS2);   // [n][degp] sender of every slot (0xffff: masked)
  float* scr = reinterpret_cast<float*>(sidx) + pl.sidx_fl;   // per-warp compacted lists; split-K partial
  int* tab = reinterpret_cast<int*>(scr + pl.scr_fl);  // [64] gslot, [96] rslot
  float* vlsum = reinterpret_cast<float*>(tab + 128);  // [64] Vl: running sum over the agents
  unsigned char* nflag = reinterpret_cast<unsigned char*>(vlsum + HID);   // [M] node is an active sender

  for (int i = threadIdx.x; i < pl.w_fl / 4; i += nth) cp_async16(ws + 4 * i, params + pl.w_off + 4 * i);
  cp_async_wait_all();
  pdl_wait();                     // weights staged while the predecessor (graph kernel) drains; now its output is visible
  pdl_launch_dependents();        // the head kernel may be scheduled as SMs free up: it stages its weights meanwhile
  __syncthreads();
  auto wptr = [&](const float* p) { return ws + ((p - params) - pl.w_off); };

  const int n = g.n, N = g.N, nd = g.nd, deg = pl.deg, degp = pl.degp;
  const int M = N - 1, pad = N - 1;
  const int n_chunks = (n + R2 - 1) / R2;
  const bool vl = net.kind == DGPPO_NET_VL;
  float4* my_al = reinterpret_cast<float4*>(scr) + (size_t)warp * degp;

  for (int gi = blockIdx.x; gi < g.n_graphs; gi += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
      tab[64] = env * g.pitch + slot;
      tab[96] = env * g.rnn_pitch + slot;
    }
    for (int i = threadIdx.x; i < (M + 3) / 4; i += nth) reinterpret_cast<int*>(nflag)[i] = 0;
    if (threadIdx.x < HID) vlsum[threadIdx.x] = 0.f;
    __syncthreads();
    const size_t gslot = (size_t)tab[64];
    const bool from_state = g.c_agent != nullptr;
    if (from_state) {
      const int env = gi / g.n_slots, slot = gi - env * g.n_slots;
      for (int row = threadIdx.x; row < M; row += nth) node_row_from_state(g, env, slot, row, x0 + (size_t)row * X0S);
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * degp; idx += nth) {
        const int i = idx / degp, t = idx - i * degp;
        int s = 0xffff;
        if (t < deg) {
          bool live;
          const int sd = slot_sender_from_state(g, x0, i, t, live);
          if (live) { s = sd; if (NL > 1) nflag[sd] = 1; }
        }
        sidx[idx] = (unsigned short)s;
      }
    } else {
    {
      const float* src = g.nodes + gslot * N * nd;
      if (nd == X0S) {
        for (int j = threadIdx.x; j < M * X0S; j += nth) cp_async4(x0 + j, src + j);
      } else {
        for (int j = threadIdx.x; j < M * 7; j += nth) {
          const int node = j / 7, c = j - node * 7;
          cp_async4(x0 + node * X0S + c, src + j);
        }
        for (int j = threadIdx.x; j < M; j += nth) x0[j * X0S + 7] = 0.f;
      }
    }
    const int* recv = g.recv + gslot * g.E;
    const int* send = g.send + gslot * g.E;
    // ---- sender table of every receiver row, once per graph (loads are independent: no branch between them)
    for (int idx = threadIdx.x; idx < n * degp; idx += nth) {
      const int i = idx / degp, t = idx - i * degp;
      int s = 0xffff;
      if (t < deg) {
        const int e = (t < n) ? i * n + t
                              : ((t < n + g.n_ag) ? n * n + i * g.n_ag + (t - n)
                                                  : n * n + n * g.n_ag + i * g.n_ao + (t - n - g.n_ag));
        const int rv = __ldg(recv + e), sd = __ldg(send + e);
        if (rv != pad) { s = sd; if (NL > 1) nflag[sd] = 1; }
      }
      sidx[idx] = (unsigned short)s;
    }
    cp_async_wait_all();
    }
    const float4* ed = from_state ? nullptr : reinterpret_cast<const float4*>(g.edges + gslot * g.E * 4);
    const GnnArgs* cgp = from_state ? &g : nullptr;
    __syncthreads();

    const float* X = x0; int XS = X0S;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const LayerP& P = net.L[l];
      const int IN = P.in, D = P.d, INP = round4(IN + 1), INA = IN + 5;
      const bool last = (l == NL - 1);
      const float *wqk = wptr(P.wqk), *wagg = wptr(P.wagg), *wu = wptr(P.wu), *bu = wptr(P.bu);
      const float isd = 1.f / sqrtf((float)D);

      for (int rc = 0; rc < n_chunks; ++rc) {
        const int a0 = rc * R2, rows = min(R2, n - a0);
        // ---- transposed receiver features of the chunk's rows
        for (int idx = threadIdx.x; idx < 32 * R2; idx += nth) {       // xr[c][r] = X[a0 + r][c]
          const int c = idx / R2, r = idx % R2;
          xr[c * RS2 + r] = (r < rows && c < IN) ? X[(size_t)(a0 + r) * XS + c] : 0.f;
        }
        __syncthreads();
        gemm_ws<R2, RS2, 2>(xr, 0, IN, wqk, H * INP, 1.f, nullptr, 0, nullptr, 0, H * INP / 4, 0, nwarps,
                            [&](int r0, int c0, float (&acc)[2][4]) {
                              const int h = c0 / INP, c = c0 - h * INP;
                              const float4 bb = *reinterpret_cast<const float4*>(wqk + IN * H * INP + c0);
#pragma unroll
                              for (int i = 0; i < 2; ++i)
                                *reinterpret_cast<float4*>(qt + ((r0 + i) * H + h) * QTS + c) =
                                    make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
                            });
        __syncthreads();
        for (int r = warp; r < R2; r += nwarps) {
          if (l == 0) attention_row_big<X0S>(r, r < rows, lane, sidx + (a0 + r) * degp, deg, degp, qt + r * H * QTS, IN, isd,
                                             X, XS, ed, a0 + r, n, g.n_ag, g.n_ao, my_al, z, cgp, x0);
          else        attention_row_big<32>(r, r < rows, lane, sidx + (a0 + r) * degp, deg, degp, qt + r * H * QTS, IN, isd,
                                            X, XS, ed, a0 + r, n, g.n_ag, g.n_ao, my_al, z, cgp, x0);
        }
        __syncthreads();
        const int KZ = H * INA, kh = KZ / 2, hw = nwarps / 2;
        const int qshift = (D == 32) ? 3 : 4;
        float* part_hi = qt;                                            // [R2][64]
        float* part_lo = scr;                                           // [R2][64]
        auto store_part = [&](float* dstp) {
          return [=](int r0, int c0, float (&acc)[2][4]) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
              *reinterpret_cast<float4*>(dstp + (r0 + i) * 64 + c0) =
                  make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          };
        };
        gemm_ws<R2, RS2, 2>(z, kh, KZ, wagg, D, 1.f / H, nullptr, 0, nullptr, 0, D / 4, hw, nwarps - hw,
                            store_part(part_hi));
        gemm_ws<R2, RS2, 2>(z, 0, kh, wagg, D, 1.f / H, xr, IN, wu, D, D / 4, 0, hw, store_part(part_lo));
        __syncthreads();
        float* ob = z;                                                  // Vl only: [64][RS2]
        for (int idx = threadIdx.x; idx < R2 * (D / 4); idx += nth) {   // combine + bias + relu
          const int r = idx >> qshift, c0 = (idx & ((D >> 2) - 1)) * 4;   // D / 4 is 8 or 16
          const float4 lo = *reinterpret_cast<const float4*>(part_lo + r * 64 + c0);
          const float4 hi = *reinterpret_cast<const float4*>(part_hi + r * 64 + c0);
          const float4 v = make_float4(fmaxf(lo.x + hi.x + bu[c0], 0.f), fmaxf(lo.y + hi.y + bu[c0 + 1], 0.f),
                                       fmaxf(lo.z + hi.z + bu[c0 + 2], 0.f), fmaxf(lo.w + hi.w + bu[c0 + 3], 0.f));
          if (last && vl) {
            ob[(c0 + 0) * RS2 + r] = v.x; ob[(c0 + 1) * RS2 + r] = v.y;
            ob[(c0 + 2) * RS2 + r] = v.z; ob[(c0 + 3) * RS2 + r] = v.w;
          } else if (r < rows) {
            if (!last) *reinterpret_cast<float4*>(x1 + (size_t)(a0 + r) * X1S + c0) = v;
            else       *reinterpret_cast<float4*>(g.rnn_out + ((size_t)tab[96] * n + a0 + r) * HID + c0) = v;
          }
        }
        if (last && vl) {
          __syncthreads();
          if (threadIdx.x < HID) {                        // sequential over agents, as the small-graph kernel
            float sacc = vlsum[threadIdx.x];
            for (int r = 0; r < rows; ++r) sacc += ob[threadIdx.x * RS2 + r];
            vlsum[threadIdx.x] = sacc;
          }
        }
        __syncthreads();
      }
      if (!last) {                          // non-agent nodes that send in the next layer
        for (int idx = threadIdx.x; idx < (M - n) * 8; idx += nth) {
          const int s = n + (idx >> 3), c0 = (idx & 7) * 4;
          if (!nflag[s]) continue;
          const float* x = X + (size_t)s * XS;
          float b0 = bu[c0], b1 = bu[c0 + 1], b2 = bu[c0 + 2], b3 = bu[c0 + 3];
          for (int c = 0; c < IN; ++c) {
            const float xv = x[c];
            const float4 w = *reinterpret_cast<const float4*>(wu + c * D + c0);
            b0 = fmaf(xv, w.x, b0); b1 = fmaf(xv, w.y, b1); b2 = fmaf(xv, w.z, b2); b3 = fmaf(xv, w.w, b3);
          }
          *reinterpret_cast<float4*>(x1 + (size_t)s * X1S + c0) =
              make_float4(fmaxf(b0, 0.f), fmaxf(b1, 0.f), fmaxf(b2, 0.f), fmaxf(b3, 0.f));
        }
        __syncthreads();
        X = x1; XS = X1S;
      }
    }
    if (vl && threadIdx.x < HID)           // centralised Vl: mean over the agents (value.py:28-31)
      g.rnn_out[(size_t)tab[96] * HID + threadIdx.x] = vlsum[threadIdx.x] / (float)n;
  }
}

// ------------------------------------------------------------------ head
// Each WARP owns 8 rows of the tile end to end (head MLP -> LayerNorm -> GRU ->
// tails); lane l owns the output columns {2l, 2l + 1}.  The A operand (the
// warp's 8 rows of one feature) is a warp-uniform 2 x LDS.128, the weight pair one
// conflict-free LDS.64: 8 FFMA2 per three shared-memory instructions, and no
// block-wide barrier inside the tile loop (warps never exchange data).

// out[c][r0..r0+WR) = bias[c] + sum_k A[k][r0..r0+WR) * W[k][c]   for c = 2 lane, 2 lane + 1.
// Packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2, new on sm_100): accumulators are (row 2p, row 2p+1)
// pairs, the A float4 supplies the row pairs directly and only the two weights are duplicated, so
// each k costs 2 MOV + WR FFMA2 instead of 2*WR FFMA (B200: 65.9 vs 42.4 TFLOP/s measured,
// tools/micro/ffma2_bench.cu).  Each half is an IEEE fma: results equal scalar fmaf bit for bit.
template <int WR, int STR = RS>
__device__ __forceinline__ void warp_dense64(const float* A, const float* W, const float* bias,
                                             float* out, int r0, int lane) {
  float2 acc[WR / 2][2];
#pragma unroll
  for (int p = 0; p < WR / 2; ++p) { acc[p][0] = make_float2(0.f, 0.f); acc[p][1] = make_float2(0.f, 0.f); }
#pragma unroll 4
  for (int k = 0; k < HID; ++k) {
    float2 ap[WR / 2];
#pragma unroll
    for (int i4 = 0; i4 < WR / 4; ++i4) {
      const float4 a = *reinterpret_cast<const float4*>(A + k * STR + r0 + 4 * i4);
      ap[2 * i4] = make_float2(a.x, a.y); ap[2 * i4 + 1] = make_float2(a.z, a.w);
    }
    const float2 w = *reinterpret_cast<const float2*>(W + k * HID + 2 * lane);
    const float2 w0d = make_float2(w.x, w.x), w1d = make_float2(w.y, w.y);
#pragma unroll
    for (int p = 0; p < WR / 2; ++p) {
      acc[p][0] = __ffma2_rn(ap[p], w0d, acc[p][0]);
      acc[p][1] = __ffma2_rn(ap[p], w1d, acc[p][1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = 2 * lane + j;
    const float bj = bias[c];
#pragma unroll
    for (int i4 = 0; i4 < WR / 4; ++i4)
      *reinterpret_cast<float4*>(out + c * STR + r0 + 4 * i4) =
          make_float4(acc[2 * i4][j].x + bj, acc[2 * i4][j].y + bj, acc[2 * i4 + 1][j].x + bj, acc[2 * i4 + 1][j].y + bj);
  }
}

// LayerNorm (flax: eps 1e-6, fast variance) + ReLU over the 64 features of the
// warp's 8 rows; 4 lanes per row, features interleaved.
template <int WR, int STR = RS>
__device__ __forceinline__ void warp_layernorm_relu(float* y, const float* scale, const float* bias,
                                                    int r0, int lane) {
  constexpr int LPR = 32 / WR;                       // lanes per row (4 or 8)
  const int r = r0 + lane / LPR, part = lane % LPR;
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < HID / LPR; ++i) { const float v = y[(i * LPR + part) * STR + r]; s += v; s2 = fmaf(v, v, s2); }
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float mean = s * (1.f / HID), mean2 = s2 * (1.f / HID);
  const float var = fmaxf(0.f, mean2 - mean * mean);
  const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int i = 0; i < HID / LPR; ++i) {
    const int c = i * LPR + part;
    y[c * STR + r] = fmaxf((y[c * STR + r] - mean) * (rstd * scale[c]) + bias[c], 0.f);
  }
}

template <int WR>
__global__ void __launch_bounds__(64 / WR * 32, 1)
head_kernel(NetP net, GnnArgs g, GnnV2Plan pl, const float* __restrict__ params) {
  constexpr int NTH = 64 / WR * 32;                   // 64 rows per CTA: 8 warps x 8 rows or 16 warps x 4 rows
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                                   // head / GRU / tail weights
  float* y0 = ws + pl.hw_fl;                          // [64][RS]
  float* y1 = y0 + HID * RS;
  float* hbuf = y1 + HID * RS;
  float* o4 = hbuf + HID * RS;                        // [4][RS]
  for (int i = threadIdx.x; i < pl.hw_fl / 4; i += NTH) cp_async16(ws + 4 * i, params + pl.hw_off + 4 * i);
  cp_async_wait_all();
  pdl_wait();
  __syncthreads();
  auto wptr = [&](const float* p) { return ws + ((p - params) - pl.hw_off); };
  const float *d0w = wptr(net.d0w), *d0b = wptr(net.d0b), *ln0s = wptr(net.ln0s), *ln0b = wptr(net.ln0b);
  const float *d1w = wptr(net.d1w), *d1b = wptr(net.d1b), *ln1s = wptr(net.ln1s), *ln1b = wptr(net.ln1b);
  const float *wi = wptr(net.wi), *bi = wptr(net.bi), *wh = wptr(net.wh), *bhn = wptr(net.bhn);
  const float *out_w = wptr(net.out_w), *out_b = wptr(net.out_b);
  const bool policy = net.kind == DGPPO_NET_POLICY;

  const int n = g.n;
  const int nr = (net.kind == DGPPO_NET_VL) ? 1 : n;       // rows per graph
  const long total_rows = (long)g.n_graphs * nr;
  const long n_wtiles = (total_rows + WR - 1) / WR;         // 8-row warp tiles
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * WR;                                 // the warp's rows inside the CTA buffers
  // row -> flat offset of its 64-float record in rnn_in / rnn_out
  auto row_off = [&](long row) {
    const long gi = row / nr; const int i = (int)(row - gi * nr);
    const long env = gi / g.n_slots; const int slot = (int)(gi - env * g.n_slots);
    return (((size_t)env * g.rnn_pitch + slot) * nr + i) * HID;
  };

  for (long wt = (long)blockIdx.x * (NTH / 32) + warp; wt < n_wtiles; wt += (long)gridDim.x * (NTH / 32)) {
    const long row0 = wt * WR;
    const int rows = (int)min((long)WR, total_rows - row0);
    __syncwarp();
    // embeddings (scratch rows in rnn_out) -> y0, previous carry -> hbuf, transposed
    for (int i = 0; i < WR; ++i) {
      if (i < rows) {
        const size_t off = row_off(row0 + i);
        cp_async4(y0 + lane * RS + r0 + i, g.rnn_out + off + lane);
        cp_async4(y0 + (lane + 32) * RS + r0 + i, g.rnn_out + off + lane + 32);
        cp_async4(hbuf + lane * RS + r0 + i, g.rnn_in + off + lane);
        cp_async4(hbuf + (lane + 32) * RS + r0 + i, g.rnn_in + off + lane + 32);
      } else {
        y0[lane * RS + r0 + i] = 0.f; y0[(lane + 32) * RS + r0 + i] = 0.f;
        hbuf[lane * RS + r0 + i] = 0.f; hbuf[(lane + 32) * RS + r0 + i] = 0.f;
      }
    }
    cp_async_wait_all();
    __syncwarp();
    // head MLP: 2 x [Dense64 -> LayerNorm -> ReLU]   (mlp.py:14-30)
    warp_dense64<WR>(y0, d0w, d0b, y1, r0, lane);
    __syncwarp();
    warp_layernorm_relu<WR>(y1, ln0s, ln0b, r0, lane);
    __syncwarp();
    warp_dense64<WR>(y1, d1w, d1b, y0, r0, lane);
    __syncwarp();
    warp_layernorm_relu<WR>(y0, ln1s, ln1b, r0, lane);
    __syncwarp();
    // GRU cell (flax GRUCell; rnn.py:19-21): 8 rows x units {2 lane, 2 lane + 1} x 3 gates
    {
      float2 ai[3][WR / 2][2], ah[3][WR / 2][2];      // (row 2p, row 2p+1) pairs, see warp_dense64
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int p = 0; p < WR / 2; ++p) {
          ai[t][p][0] = ai[t][p][1] = make_float2(0.f, 0.f);
          ah[t][p][0] = ah[t][p][1] = make_float2(0.f, 0.f);
        }
#pragma unroll 2
      for (int k = 0; k < HID; ++k) {
        float2 xp[WR / 2], hp[WR / 2];
#pragma unroll
        for (int i4 = 0; i4 < WR / 4; ++i4) {
          const float4 xa = *reinterpret_cast<const float4*>(y0 + k * RS + r0 + 4 * i4);
          const float4 ha = *reinterpret_cast<const float4*>(hbuf + k * RS + r0 + 4 * i4);
          xp[2 * i4] = make_float2(xa.x, xa.y); xp[2 * i4 + 1] = make_float2(xa.z, xa.w);
          hp[2 * i4] = make_float2(ha.x, ha.y); hp[2 * i4 + 1] = make_float2(ha.z, ha.w);
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const float2 wiv = *reinterpret_cast<const float2*>(wi + k * 192 + t * 64 + 2 * lane);
          const float2 whv = *reinterpret_cast<const float2*>(wh + k * 192 + t * 64 + 2 * lane);
          const float2 wi0d = make_float2(wiv.x, wiv.x), wi1d = make_float2(wiv.y, wiv.y);
          const float2 wh0d = make_float2(whv.x, whv.x), wh1d = make_float2(whv.y, whv.y);
#pragma unroll
          for (int p = 0; p < WR / 2; ++p) {
            ai[t][p][0] = __ffma2_rn(xp[p], wi0d, ai[t][p][0]); ai[t][p][1] = __ffma2_rn(xp[p], wi1d, ai[t][p][1]);
            ah[t][p][0] = __ffma2_rn(hp[p], wh0d, ah[t][p][0]); ah[t][p][1] = __ffma2_rn(hp[p], wh1d, ah[t][p][1]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * lane + j;
        const float bir = bi[c], biz = bi[64 + c], bin = bi[128 + c], bh = bhn[c];
        float hn[WR];
#pragma unroll
        for (int i = 0; i < WR; ++i) {
          const int p = i >> 1;
          const float air = (i & 1) ? ai[0][p][j].y : ai[0][p][j].x, ahr = (i & 1) ? ah[0][p][j].y : ah[0][p][j].x;
          const float aiz = (i & 1) ? ai[1][p][j].y : ai[1][p][j].x, ahz = (i & 1) ? ah[1][p][j].y : ah[1][p][j].x;
          const float ain = (i & 1) ? ai[2][p][j].y : ai[2][p][j].x, ahn = (i & 1) ? ah[2][p][j].y : ah[2][p][j].x;
          const float hprev = hbuf[c * RS + r0 + i];
          const float rgate = sigmoidf_(air + bir + ahr);
          const float zgate = sigmoidf_(aiz + biz + ahz);
          const float cand = tanhf(ain + bin + rgate * (ahn + bh));
          hn[i] = (1.f - zgate) * cand + zgate * hprev;
        }
#pragma unroll
        for (int i4 = 0; i4 < WR / 4; ++i4)
          *reinterpret_cast<float4*>(y1 + c * RS + r0 + 4 * i4) =
              make_float4(hn[4 * i4], hn[4 * i4 + 1], hn[4 * i4 + 2], hn[4 * i4 + 3]);
        // new carry -> rnn_out (overwrites the scratch embeddings); coalesced across lanes
#pragma unroll
        for (int i = 0; i < WR; ++i)
          if (i < rows) g.rnn_out[row_off(row0 + i) + c] = hn[i];
      }
    }
    __syncwarp();
    const float* feat = y1;                                       // ScaleHid is folded into out_w at pack time
    {   // out: [64] -> 4 columns; lane = (row, column)
      const int rr = lane >> 2, col = lane & 3;
      if (rr < WR) {
        float acc = out_b[col];
#pragma unroll 8
        for (int k = 0; k < HID; ++k) acc = fmaf(feat[k * RS + r0 + rr], out_w[k * 4 + col], acc);
        o4[col * RS + r0 + rr] = acc;
      }
    }
    __syncwarp();
    if (lane < rows) {
      const int r = r0 + lane;
      const long row = row0 + lane;
      const long gi = row / nr; const int i = (int)(row - gi * nr);
      const long env = gi / g.n_slots; const int slot = (int)(gi - env * g.n_slots);
      if (policy) {
        policy_tail(g, o4[r], o4[RS + r], o4[2 * RS + r], o4[3 * RS + r], (int)env, slot, i, n);
      } else {
        float* vo = g.value + ((((size_t)env * g.out_pitch + slot) * nr + i) * net.n_out);
        for (int c = 0; c < net.n_out; ++c) vo[c] = o4[c * RS + r];
      }
    }
  }
}

// head_kernel_wide: 128 rows per CTA, 16 warps x 8 rows.  Compared with head_kernel<4> the weight rows
// are amortised over twice the rows per warp (4-6 FMA per shared-memory wavefront instead of 2.7-3.4;
// the kernel is smem-bandwidth-bound, profiles/), which 512 threads x 128 registers allow because
//   * the r and z gates accumulate x Wi + h Wh into ONE accumulator each (they are summed anyway),
//   * only two [64][R3S] activation buffers exist: the GRU carry is prefetched into registers at the
//     start of the tile and parked in the buffer the second Dense has just released.
constexpr int R3 = 128, R3S = 132, WR3 = 8;

__global__ void __launch_bounds__(512, 1)
head_kernel_wide(NetP net, GnnArgs g, GnnV2Plan pl, const float* __restrict__ params) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                                   // head / GRU / tail weights
  float* b0 = ws + pl.hw_fl;                          // [64][R3S]
  float* b1 = b0 + HID * R3S;                         // [64][R3S]
  float* o4 = b1 + HID * R3S;                         // [4][R3S]
  for (int i = threadIdx.x; i < pl.hw_fl / 4; i += 512) cp_async16(ws + 4 * i, params + pl.hw_off + 4 * i);
  cp_async_wait_all();
  pdl_wait();                     // weights staged while gnn_layers drains; its embeddings are visible from here on
  __syncthreads();
  auto wptr = [&](const float* p) { return ws + ((p - params) - pl.hw_off); };
  const float *d0w = wptr(net.d0w), *d0b = wptr(net.d0b), *ln0s = wptr(net.ln0s), *ln0b = wptr(net.ln0b);
  const float *d1w = wptr(net.d1w), *d1b = wptr(net.d1b), *ln1s = wptr(net.ln1s), *ln1b = wptr(net.ln1b);
  const float *wi = wptr(net.wi), *bi = wptr(net.bi), *wh = wptr(net.wh), *bhn = wptr(net.bhn);
  const float *out_w = wptr(net.out_w), *out_b = wptr(net.out_b);
  const bool policy = net.kind == DGPPO_NET_POLICY;

  const int n = g.n;
  const int nr = (net.kind == DGPPO_NET_VL) ? 1 : n;       // rows per graph
  const long total_rows = (long)g.n_graphs * nr;
  const long n_wtiles = (total_rows + WR3 - 1) / WR3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * WR3;
  // row -> (env, slot, agent): 32-bit divisions when the row count allows, and once per row (lane i decodes
  // the warp tile's row i; the others get the offset by shuffle) - the 64-bit form cost 10 % of the kernel
  const bool small_rows = total_rows < 0x7fffffffL;
  auto decode = [&](long row, int& env, int& slot, int& i) {
    if (small_rows) {
      const unsigned r = (unsigned)row, gi = r / (unsigned)nr;
      i = (int)(r - gi * (unsigned)nr);
      const unsigned e = (g.n_slots == 1) ? gi : gi / (unsigned)g.n_slots;
      env = (int)e; slot = (int)(gi - e * (unsigned)g.n_slots);
    } else {
      const long gi = row / nr; i = (int)(row - gi * nr);
      const long e = gi / g.n_slots; env = (int)e; slot = (int)(gi - e * g.n_slots);
    }
  };

  // warp tiles are dealt round-robin over the CTAs (tile wt -> CTA wt % grid), so a partial last round
  // thins out every SM evenly instead of leaving whole SMs idle
  for (long wt = (long)warp * gridDim.x + blockIdx.x; wt < n_wtiles; wt += (long)gridDim.x * 16) {
    const long row0 = wt * WR3;
    const int rows = (int)min((long)WR3, total_rows - row0);
    __syncwarp();
    int my_env = 0, my_slot = 0, my_i = 0;
    unsigned long long my_off = 0;
    if (lane < rows) {
      decode(row0 + lane, my_env, my_slot, my_i);
      my_off = (((size_t)my_env * g.rnn_pitch + my_slot) * nr + my_i) * HID;
    }
    // embeddings (scratch rows in rnn_out) -> b0 (transposed, async); previous carry -> registers
    float hreg[WR3][2];
#pragma unroll
    for (int i = 0; i < WR3; ++i) {
      if (i < rows) {
        const size_t off = __shfl_sync(0xffffffffu, my_off, i);
        cp_async4(b0 + lane * R3S + r0 + i, g.rnn_out + off + lane);
        cp_async4(b0 + (lane + 32) * R3S + r0 + i, g.rnn_out + off + lane + 32);
        const float2 hv = __ldg(reinterpret_cast<const float2*>(g.rnn_in + off) + lane);   // units 2 lane, 2 lane + 1
        hreg[i][0] = hv.x; hreg[i][1] = hv.y;
      } else {
        b0[lane * R3S + r0 + i] = 0.f; b0[(lane + 32) * R3S + r0 + i] = 0.f;
        hreg[i][0] = 0.f; hreg[i][1] = 0.f;
      }
    }
    cp_async_wait_all();
    __syncwarp();
    warp_dense64<WR3, R3S>(b0, d0w, d0b, b1, r0, lane);          // head MLP (mlp.py:14-30)
    __syncwarp();
    warp_layernorm_relu<WR3, R3S>(b1, ln0s, ln0b, r0, lane);
    __syncwarp();
    warp_dense64<WR3, R3S>(b1, d1w, d1b, b0, r0, lane);
    __syncwarp();
    warp_layernorm_relu<WR3, R3S>(b0, ln1s, ln1b, r0, lane);
    // park the carry (transposed) in b1, which the second Dense no longer needs
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i4 = 0; i4 < WR3 / 4; ++i4)
        *reinterpret_cast<float4*>(b1 + (2 * lane + j) * R3S + r0 + 4 * i4) =
            make_float4(hreg[4 * i4][j], hreg[4 * i4 + 1][j], hreg[4 * i4 + 2][j], hreg[4 * i4 + 3][j]);
    __syncwarp();
    // GRU cell (flax GRUCell; rnn.py:19-21): x = b0, h = b1; units {2 lane, 2 lane + 1} (weights: LDS.64)
    float hn[WR3][2];
    {
      float2 ar[WR3 / 2][2], az[WR3 / 2][2], an[WR3 / 2][2], ahn[WR3 / 2][2];
#pragma unroll
      for (int p = 0; p < WR3 / 2; ++p)
#pragma unroll
        for (int j = 0; j < 2; ++j) ar[p][j] = az[p][j] = an[p][j] = ahn[p][j] = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int k = 0; k < HID; ++k) {
        float2 xp[WR3 / 2], hp[WR3 / 2];
#pragma unroll
        for (int i4 = 0; i4 < WR3 / 4; ++i4) {
          const float4 xa = *reinterpret_cast<const float4*>(b0 + k * R3S + r0 + 4 * i4);
          const float4 ha = *reinterpret_cast<const float4*>(b1 + k * R3S + r0 + 4 * i4);
          xp[2 * i4] = make_float2(xa.x, xa.y); xp[2 * i4 + 1] = make_float2(xa.z, xa.w);
          hp[2 * i4] = make_float2(ha.x, ha.y); hp[2 * i4 + 1] = make_float2(ha.z, ha.w);
        }
        const float2* wik = reinterpret_cast<const float2*>(wi + k * 192) + lane;
        const float2* whk = reinterpret_cast<const float2*>(wh + k * 192) + lane;
        const float2 vir = wik[0], viz = wik[32], vin = wik[64], vhr = whk[0], vhz = whk[32], vhn = whk[64];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float sir = j ? vir.y : vir.x, siz = j ? viz.y : viz.x, sin_ = j ? vin.y : vin.x;
          const float shr = j ? vhr.y : vhr.x, shz = j ? vhz.y : vhz.x, shn = j ? vhn.y : vhn.x;
          const float2 wir = make_float2(sir, sir), whr = make_float2(shr, shr);
          const float2 wiz = make_float2(siz, siz), whz = make_float2(shz, shz);
          const float2 win = make_float2(sin_, sin_), whn = make_float2(shn, shn);
#pragma unroll
          for (int p = 0; p < WR3 / 2; ++p) {
            ar[p][j] = __ffma2_rn(xp[p], wir, ar[p][j]); ar[p][j] = __ffma2_rn(hp[p], whr, ar[p][j]);
            az[p][j] = __ffma2_rn(xp[p], wiz, az[p][j]); az[p][j] = __ffma2_rn(hp[p], whz, az[p][j]);
            an[p][j] = __ffma2_rn(xp[p], win, an[p][j]);
            ahn[p][j] = __ffma2_rn(hp[p], whn, ahn[p][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * lane + j;
        const float bir = bi[c], biz = bi[64 + c], bin = bi[128 + c], bh = bhn[c];
#pragma unroll
        for (int i = 0; i < WR3; ++i) {
          const int p = i >> 1;
          const float sr = (i & 1) ? ar[p][j].y : ar[p][j].x, sz = (i & 1) ? az[p][j].y : az[p][j].x;
          const float sn = (i & 1) ? an[p][j].y : an[p][j].x, sh = (i & 1) ? ahn[p][j].y : ahn[p][j].x;
          const float rgate = gate_sigmoid(sr + bir);
          const float zgate = gate_sigmoid(sz + biz);
          const float cand = gate_tanh(sn + bin + rgate * (sh + bh));
          hn[i][j] = (1.f - zgate) * cand + zgate * b1[c * R3S + r0 + i];   // previous carry (parked in b1)
        }
      }
    }
    __syncwarp();                                                   // every lane is done reading b0 / b1
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = 2 * lane + j;
#pragma unroll
      for (int i4 = 0; i4 < WR3 / 4; ++i4)
        *reinterpret_cast<float4*>(b0 + c * R3S + r0 + 4 * i4) =
            make_float4(hn[4 * i4][j], hn[4 * i4 + 1][j], hn[4 * i4 + 2][j], hn[4 * i4 + 3][j]);
    }
#pragma unroll
    for (int i = 0; i < WR3; ++i)                                   // new carry, one float2 per lane (coalesced)
      if (i < rows)
        reinterpret_cast<float2*>(g.rnn_out + __shfl_sync(0xffffffffu, my_off, i))[lane] = make_float2(hn[i][0], hn[i][1]);
    __syncwarp();
    const float* feat = b0;                                         // ScaleHid is folded into out_w at pack time
    {   // out: [64] -> 4 columns; lane = (row, column)
      const int rr = lane >> 2, col = lane & 3;
      float acc = out_b[col];
#pragma unroll 8
      for (int k = 0; k < HID; ++k) acc = fmaf(feat[k * R3S + r0 + rr], out_w[k * 4 + col], acc);
      o4[col * R3S + r0 + rr] = acc;
    }
    __syncwarp();
    if (lane < rows) {
      const int r = r0 + lane;
      const int env = my_env, slot = my_slot, i = my_i;
      if (policy) {
        policy_tail(g, o4[r], o4[R3S + r], o4[2 * R3S + r], o4[3 * R3S + r], env, slot, i, n);
      } else {
        float* vo = g.value + ((((size_t)env * g.out_pitch + slot) * nr + i) * net.n_out);
        for (int c = 0; c < net.n_out; ++c) vo[c] = o4[c * R3S + r];
      }
    }
  }
}


int launch_gnn_v2(void* stream, const NetP& P, const DgppoNetLayout& L, const float* params,
                  const GnnArgs& g_in, int sms, int phase) {
  if (!g_in.rnn_out || g_in.rnn_out == g_in.rnn_in) return DGPPO_V2_UNSUPPORTED;
  GnnArgs g = g_in;
  const bool big = g.n > R2;
  g.G = big ? 1 : R2 / g.n;
  GnnV2Plan pl;
  pl.w_off = L.wqk[0];
  pl.w_fl = L.d0w - L.wqk[0];
  pl.hw_off = L.d0w;
  pl.hw_fl = L.wq[0] - L.d0w;                       // up to the fallback-kernel blocks
  pl.m_cap = g.G * (g.N - 1);
  pl.deg = g.n + g.n_ag + g.n_ao;
  pl.x0_fl = pl.m_cap * X0S;
  if (P.kind == DGPPO_NET_VL && pl.x0_fl < HID * RS2) pl.x0_fl = HID * RS2;
  pl.x1_fl = (P.n_layers == 2) ? pl.m_cap * X1S : 0;
  pl.degp = ((pl.deg + 31) / 32) * 32;
  if (!big && pl.degp > 64) return DGPPO_V2_UNSUPPORTED;
  if (big && ((long)pl.m_cap * X1S > 0xffff || pl.degp > 0x7fff)) return DGPPO_V2_UNSUPPORTED;   // packed list entries
  pl.sidx_fl = big ? round4((g.n * pl.degp + 1) / 2) : R2 * pl.degp;    // large graphs: uint16, all rows
  const size_t base_fl = (size_t)pl.w_fl + pl.x0_fl + pl.x1_fl + 32 * RS2 + R2 * H * QTS + 112 * RS2 + pl.sidx_fl +
                         128 /* tab */ + (big ? HID : 0) /* vlsum */ + round4((pl.m_cap + 3) / 4) /* nflag */;
  pl.threads = 256;
  // per-warp attention scratch: small graphs (a, offset) + edge-feature lists, large graphs one float4 list
  pl.scr_fl = (pl.threads / 32) * pl.degp * (big ? 4 : 8);
  if (pl.scr_fl < R2 * 64) pl.scr_fl = R2 * 64;      // also hosts the split-K partial [R2][64]
  pl.smem_bytes = (base_fl + pl.scr_fl) * sizeof(float);
  pl.head_smem_bytes = ((size_t)pl.hw_fl + 3 * HID * RS + 4 * RS) * sizeof(float);
  if (pl.smem_bytes > 227 * 1024 || pl.head_smem_bytes > 227 * 1024) return DGPPO_V2_UNSUPPORTED;
  if ((pl.w_off & 3) || (pl.w_fl & 3) || (pl.hw_fl & 3)) return DGPPO_V2_UNSUPPORTED;

  cudaStream_t st = (cudaStream_t)stream;
  // Programmatic dependent launch (weights / TMEM / barriers set up while the predecessor drains): 1 gnn + head
  // (default), 0 off, 2 gnn only, 3 head only.  Round-2 measurements (captured rollouts, compact record, tcgen05 head;
  // C3): 512 envs 6.92 -> 5.99 ms per rollout (2: 6.09, 3: 6.82), 4096 envs 28.36 -> 28.16 ms.  (Round 1, FFMA head,
  // uncaptured, 4 streams: +2.9 %, which is why it used to be opt-in.)
  static const char* pdl_env = getenv("DGPPO_PDL");
  const int pdl_mode = pdl_env ? atoi(pdl_env) : 1;
  const bool pdl = pdl_mode == 1 || pdl_mode == 2, pdl_h = pdl_mode == 1 || pdl_mode == 3;
  const int n_tiles = (g.n_graphs + g.G - 1) / g.G;
  const int grid1 = n_tiles < 2 * sms ? n_tiles : 2 * sms;
  cudaError_t err = cudaSuccess;
  if (phase == 2) {
    // head only: the embeddings are already in rnn_out
  } else if (big) {
    const int gridb = g.n_graphs < sms ? g.n_graphs : sms;
    if (P.n_layers == 2) {
      err = cudaFuncSetAttribute(gnn_layers_big_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
      if (err != cudaSuccess) return (int)err;
      launch_pdl(pdl, gnn_layers_big_kernel<2>, gridb, pl.threads, pl.smem_bytes, st, P, g, pl, params);
    } else {
      err = cudaFuncSetAttribute(gnn_layers_big_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
      if (err != cudaSuccess) return (int)err;
      launch_pdl(pdl, gnn_layers_big_kernel<1>, gridb, pl.threads, pl.smem_bytes, st, P, g, pl, params);
    }
  } else if (P.n_layers == 2) {
    err = cudaFuncSetAttribute(gnn_layers_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (err != cudaSuccess) return (int)err;
    launch_pdl(pdl, gnn_layers_kernel<2>, grid1, pl.threads, pl.smem_bytes, st, P, g, pl, params);
  } else {
    err = cudaFuncSetAttribute(gnn_layers_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (err != cudaSuccess) return (int)err;
    launch_pdl(pdl, gnn_layers_kernel<1>, grid1, pl.threads, pl.smem_bytes, st, P, g, pl, params);
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) return (int)err;
  if (phase == 1) return 0;
  const int nr = (P.kind == DGPPO_NET_VL) ? 1 : g.n;
  const long total_rows = (long)g.n_graphs * nr;
  const long h_tiles = (total_rows + R - 1) / R;             // CTAs worth of 8-row warp tiles
  const int grid2 = h_tiles < sms ? (int)h_tiles : sms;
  // head variant: "tc" (default: tcgen05 3xTF32, head_tc.cu) | "wide" | "wr4" | "wr8" (FFMA kernels, A/B runs)
  static const char* hv_cached = getenv("DGPPO_HEAD");
  const char* hv = hv_cached;
  const size_t hvlen = hv ? strlen(hv) : 0;
  if (!hv || (hvlen >= 2 && hv[0] == 't' && hv[1] == 'c'))
    return launch_head_tc(st, P, g, params + L.tc_head, sms, pdl_h);
  const size_t wide_smem = ((size_t)pl.hw_fl + 2 * HID * R3S + 4 * R3S) * sizeof(float);
  const bool want_wide = hvlen >= 2 && hv[0] == 'w' && hv[1] == 'i';
  if (want_wide && wide_smem <= 227 * 1024) {
    const long w_tiles = (total_rows + R3 - 1) / R3;
    const int grid3 = w_tiles < sms ? (int)w_tiles : sms;      // no more CTAs than full 128-row tiles: spreading a
                                                               // small batch over all SMs blocks the other streams
    err = cudaFuncSetAttribute(head_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem);
    if (err != cudaSuccess) return (int)err;
    launch_pdl(pdl_h, head_kernel_wide, grid3, 512, wide_smem, st, P, g, pl, params);
    return (int)cudaGetLastError();
  }
  if (hvlen >= 3 && hv[2] == '8') {
    err = cudaFuncSetAttribute(head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.head_smem_bytes);
    if (err != cudaSuccess) return (int)err;
    launch_pdl(pdl_h, head_kernel<8>, grid2, 256, pl.head_smem_bytes, st, P, g, pl, params);
  } else {
    err = cudaFuncSetAttribute(head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.head_smem_bytes);
    if (err != cudaSuccess) return (int)err;
    launch_pdl(pdl_h, head_kernel<4>, grid2, 512, pl.head_smem_bytes, st, P, g, pl, params);
  }
  return (int)cudaGetLastError();
}

}  // namespace dgppo
