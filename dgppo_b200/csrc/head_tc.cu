// K4 head on the 5th-generation tensor cores (tcgen05, sm_100a): head MLP + LayerNorm + GRU + output
// tails for a tile of 128 rows (agent rows, or one pooled row per graph for Vl)
// (nn/mlp.py:14-30, nn/rnn.py:14-30, algo/module/policy.py:61-74, value.py:15-79).
//
// Precision: every GEMM is a 3xTF32 product (A_hi B_hi + A_lo B_hi + A_hi B_lo) accumulated in fp32 in
// TMEM, so the result stays at fp32 level (measured max |err| 2.5e-6 on |values| <= 2.7 for K = 64,
// tools/micro/tc_tf32_test.cu; the fp32 FMA loop beside it: 0.7e-6) - inside the path's rtol 1e-5.
//
// Mapping (512 threads, one CTA per SM): warp w owns TMEM lane quadrant q = w & 3 (rows 32 q .. 32 q + 31 of the
// tile, lane = row) and column group cg = w >> 2 (16 of the 64 feature columns): four threads share a row.
//   TMEM (512 columns): [0,256) accumulators, [256,320) x_hi, [320,384) x_lo, [384,448) h_hi, [448,512) h_lo.
//     The A operands live in TMEM (tcgen05.mma with A from TMEM): the threads that own a row write its
//     hi / lo split with tcgen05.st and read the accumulator row back with tcgen05.ld; no activation ever
//     goes through shared memory.  The per-row LayerNorm statistics and the 64 -> 4 output Dense are reduced
//     over the four column groups through spare TMEM columns (tcgen05.st, named barrier of the quadrant's
//     four warps, tcgen05.ld) - shared memory is full of weights.
//   smem (224 KB): the GRU weights hi + lo, resident for the whole kernel (2 x 96 KB, B operands in the
//     no-swizzle K-major canonical layout, pre-arranged at pack time), plus one 32 KB slot through which
//     the two Dense64 blocks are streamed from L2 by 1-D bulk copies (cp.async.bulk + mbarrier), each
//     fetched while the previous stage's epilogue runs.
//   Per tile: x = embedding, h = carry -> split -> TMEM | MMA Dense0 | LN+ReLU -> split -> TMEM | MMA Dense1 |
//     LN+ReLU -> split -> TMEM | MMA x [Wir Wiz Win] (N = 192), h [Whr Whz] (N = 128, accumulated onto r, z),
//     h Whn (N = 64) | gates, new carry, output Dense (64 x 4, FFMA), TanhNormal sample / log-prob or value.
//   One elected thread issues the MMAs (72 + 24 + 24 per tile) and commits them to an mbarrier the CTA
//   waits on.  (First version: 128 threads, thread = whole row: 48 us per 32768 rows, one warp per scheduler
//   and latency-bound, profiles/r2_head_tc1.ncu.txt; this version spreads a row over four warps.)
#include "gnn_common.cuh"
#include "tc_common.cuh"

namespace dgppo {

using namespace tc;

namespace {

constexpr int TC_ROWS = 128, TC_THREADS = 512;
constexpr unsigned COL_XH = 256, COL_XL = 320, COL_HH = 384, COL_HL = 448;
constexpr unsigned COL_STAT = 64;     // LayerNorm partial sums (s, s2) x 4 column groups: accumulator columns idle then
constexpr unsigned COL_OUT = COL_XH;  // output-Dense partial sums 4 x 4: x operand columns, dead after the GRU MMAs
constexpr int N_SMALL = 10 * 64;      // d0b ln0s ln0b d1b ln1s ln1b bi(192) bhn

struct __align__(128) HeadTcSmem {
  float gx[TC_G_FL];                  // x-part GRU weights [hi | lo], each [16 kc][192 n][4]
  float gh[TC_G_FL];                  // h-part
  float dslot[TC_D_FL];               // Dense64 block in flight [hi | lo], each [16 kc][64 n][4]
  float small_p[N_SMALL];
  unsigned long long bar_g, bar_d, bar_mma;
  unsigned tmem_base;
};
static_assert(sizeof(HeadTcSmem) <= 227 * 1024, "head_tc shared memory plan");

__device__ __forceinline__ void wait_or_trap(unsigned long long* bar, unsigned parity) {
  if (!mbar_wait(bar, parity)) __trap();      // a broken pipeline fails the launch instead of hanging the GPU
}

// 16 fp32 values of a row -> hi / lo TF32 split -> TMEM columns [col_hi, +16), [col_lo, +16) of this thread's lane.
// lo keeps the exact remainder (hi + lo == v in fp32; the tensor core reads its top 19 bits).
__device__ __forceinline__ void split_store16(unsigned addr_hi, unsigned addr_lo, const float (&v)[16]) {
  unsigned hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float h = tf32_rna(v[j]);
    hi[j] = __float_as_uint(h);
    lo[j] = __float_as_uint(v[j] - h);
  }
  tmem_st16(addr_hi, hi);
  tmem_st16(addr_lo, lo);
}

// 16 floats of a global row -> TMEM (zeros for rows past the end)
__device__ __forceinline__ void row16_to_tmem(const float* __restrict__ src, bool valid, unsigned addr_hi, unsigned addr_lo) {
  float v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = valid ? *reinterpret_cast<const float4*>(src + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  split_store16(addr_hi, addr_lo, v);
}

// quadrant exchange: every tcgen05.st of the quadrant's four warps is visible to their tcgen05.ld afterwards
__device__ __forceinline__ void quadrant_exchange(int q) {
  tmem_st_wait();
  tc_fence_before();
  named_bar_sync(1 + q, 128);
  tc_fence_after();
}

// Dense64 accumulator (TMEM columns [0,64)) + bias -> LayerNorm (flax: eps 1e-6, fast variance) -> ReLU
// -> split -> x operand columns; this thread: columns [c0, c0 + 16) of its row.
__device__ __forceinline__ void epilogue_ln(unsigned lane_addr, int q, int cg, const float* __restrict__ bias,
                                            const float* __restrict__ scale, const float* __restrict__ shift) {
  const int c0 = cg * 16;
  float v[16];
  {
    unsigned u[16];
    tmem_ld16(lane_addr + c0, u);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]) + bias[c0 + j];
  }
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) { s += v[j]; s2 = fmaf(v[j], v[j], s2); }
  tmem_st2(lane_addr + COL_STAT + 2 * cg, __float_as_uint(s), __float_as_uint(s2));
  quadrant_exchange(q);
  {
    unsigned u[8];
    tmem_ld8(lane_addr + COL_STAT, u);
    tmem_ld_wait();
    s = (__uint_as_float(u[0]) + __uint_as_float(u[2])) + (__uint_as_float(u[4]) + __uint_as_float(u[6]));
    s2 = (__uint_as_float(u[1]) + __uint_as_float(u[3])) + (__uint_as_float(u[5]) + __uint_as_float(u[7]));
  }
  const float mean = s * (1.f / HID), mean2 = s2 * (1.f / HID);
  const float var = fmaxf(0.f, mean2 - mean * mean);
  const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = fmaxf((v[j] - mean) * (rstd * scale[c0 + j]) + shift[c0 + j], 0.f);
  split_store16(lane_addr + COL_XH + c0, lane_addr + COL_XL + c0, v);
}

// 3xTF32 product of the K = 64 operand at TMEM columns (a_hi, a_lo) with a B block [hi | lo] in smem
// ([16 kc][ldn n][4] each; b_lo_off = float offset of the lo copy), N columns starting at row n0 of B,
// into accumulator columns d_col.  first: the first MMA overwrites the accumulator.
template <int N>
__device__ __forceinline__ void issue_3xtf32(unsigned tmem, unsigned a_hi, unsigned a_lo, const float* b_hi,
                                             int b_lo_off, int ldn, int n0, unsigned d_col, bool first) {
  constexpr unsigned idesc = make_idesc_tf32(TC_ROWS, N);
  const unsigned lbo = (unsigned)ldn * 16u;
  const unsigned long long dhi = make_sdesc(b_hi + n0 * 4, lbo, 128);
  const unsigned long long dlo = make_sdesc(b_hi + b_lo_off + n0 * 4, lbo, 128);
  const unsigned kstep = (2u * lbo) >> 4;                  // descriptor address units (16 B) per k-step of 8
#pragma unroll
  for (int part = 0; part < 3; ++part) {
    const unsigned a = tmem + (part == 1 ? a_lo : a_hi);
    const unsigned long long b = (part == 2) ? dlo : dhi;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      mma_tf32_ts(tmem + d_col, a + ks * 8, b + (unsigned long long)(ks * kstep), idesc, !(first && part == 0 && ks == 0));
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
head_tc_kernel(NetP net, GnnArgs g, const float* __restrict__ tcw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  HeadTcSmem& S = *reinterpret_cast<HeadTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = warp >> 2, c0 = cg * 16;
  const int n = g.n;
  const int nr = (net.kind == DGPPO_NET_VL) ? 1 : n;       // rows per graph
  const long total_rows = (long)g.n_graphs * nr;
  const long n_tiles = (total_rows + TC_ROWS - 1) / TC_ROWS;
  if ((long)blockIdx.x >= n_tiles) return;                 // (the launcher never over-provisions; belt and braces)

  if (tid == 0) {
    mbar_init(&S.bar_g, 1); mbar_init(&S.bar_d, 1); mbar_init(&S.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&S.tmem_base, 512);
  {
    const float* src[8] = {net.d0b, net.ln0s, net.ln0b, net.d1b, net.ln1s, net.ln1b, net.bi, net.bhn};
    const int len[8] = {64, 64, 64, 64, 64, 64, 192, 64};
    int o = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      if (tid < len[a]) S.small_p[o + tid] = __ldg(src[a] + tid);
      o += len[a];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = S.tmem_base;
  const unsigned lane_addr = tmem + ((unsigned)(q * 32) << 16);
  const float *d0b = S.small_p, *ln0s = d0b + 64, *ln0b = d0b + 128, *d1b = d0b + 192, *ln1s = d0b + 256,
              *ln1b = d0b + 320, *bi = d0b + 384, *bhn = d0b + 576;

  // weights: GRU blocks once (resident), Dense0 of the first tile
  if (tid == 0) {
    mbar_expect_tx(&S.bar_g, 2u * TC_G_FL * 4u);
#pragma unroll
    for (int p = 0; p < 6; ++p)                            // 32 KB pieces
      bulk_g2s(S.gx + p * 8192, tcw + 2 * TC_D_FL + p * 8192, 32768u, &S.bar_g);
    mbar_expect_tx(&S.bar_d, TC_D_FL * 4u);
    bulk_g2s(S.dslot, tcw, TC_D_FL * 4u, &S.bar_d);
  }
  pdl_wait();                     // weights in flight while gnn_layers drains; its embeddings are visible from here on

  const bool policy = net.kind == DGPPO_NET_POLICY;
  unsigned ph_d = 0, ph_mma = 0;
  bool g_ready = false;
  auto decode = [&](long row, int& env, int& slot, int& ai) -> size_t {
    const long gi = row / nr; ai = (int)(row - gi * nr);
    const long e = gi / g.n_slots; env = (int)e; slot = (int)(gi - e * g.n_slots);
    return (((size_t)env * g.rnn_pitch + slot) * nr + ai) * HID;
  };
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long row = tile * TC_ROWS + q * 32 + lane;
    const bool valid = row < total_rows;
    int env = 0, slot = 0, ai = 0;
    const size_t off = valid ? decode(row, env, slot, ai) : 0;
    {   // next tile's rows towards L1 while this tile computes (they sit in L2: gnn_layers has just written them)
      const long nrow = row + (long)gridDim.x * TC_ROWS;
      if (nrow < total_rows) {
        int e2, s2, a2;
        const size_t noff = decode(nrow, e2, s2, a2);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(g.rnn_out + noff + c0));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(g.rnn_in + noff + c0));
      }
    }
    // ---- operands: x = embedding (scratch rows of rnn_out), h = previous carry; this thread: 16 columns
    row16_to_tmem(g.rnn_out + off + c0, valid, lane_addr + COL_XH + c0, lane_addr + COL_XL + c0);
    row16_to_tmem(g.rnn_in + off + c0, valid, lane_addr + COL_HH + c0, lane_addr + COL_HL + c0);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    // ---- Dense0
    if (tid == 0) {
      wait_or_trap(&S.bar_d, ph_d);
      tc_fence_after();
      issue_3xtf32<64>(tmem, COL_XH, COL_XL, S.dslot, TC_D_FL / 2, 64, 0, 0, true);
      mma_commit(&S.bar_mma);
    }
    ph_d ^= 1;
    wait_or_trap(&S.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    if (tid == 0) {                                        // the slot is free: fetch Dense1 under the epilogue
      mbar_expect_tx(&S.bar_d, TC_D_FL * 4u);
      bulk_g2s(S.dslot, tcw + TC_D_FL, TC_D_FL * 4u, &S.bar_d);
    }
    epilogue_ln(lane_addr, q, cg, d0b, ln0s, ln0b);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    // ---- Dense1
    if (tid == 0) {
      wait_or_trap(&S.bar_d, ph_d);
      tc_fence_after();
      issue_3xtf32<64>(tmem, COL_XH, COL_XL, S.dslot, TC_D_FL / 2, 64, 0, 0, true);
      mma_commit(&S.bar_mma);
    }
    ph_d ^= 1;
    wait_or_trap(&S.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    if (tid == 0 && tile + gridDim.x < n_tiles) {          // Dense0 of the next tile
      mbar_expect_tx(&S.bar_d, TC_D_FL * 4u);
      bulk_g2s(S.dslot, tcw, TC_D_FL * 4u, &S.bar_d);
    }
    epilogue_ln(lane_addr, q, cg, d1b, ln1s, ln1b);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    // ---- GRU products: r | z | n_i at columns [0,192), h-part of r, z accumulated, n_h at [192,256)
    if (tid == 0) {
      if (!g_ready) wait_or_trap(&S.bar_g, 0);
      tc_fence_after();
      issue_3xtf32<192>(tmem, COL_XH, COL_XL, S.gx, TC_G_FL / 2, 192, 0, 0, true);
      issue_3xtf32<128>(tmem, COL_HH, COL_HL, S.gh, TC_G_FL / 2, 192, 0, 0, false);
      issue_3xtf32<64>(tmem, COL_HH, COL_HL, S.gh, TC_G_FL / 2, 192, 128, 192, true);
      mma_commit(&S.bar_mma);
    }
    g_ready = true;
    wait_or_trap(&S.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    // ---- gates (flax GRUCell) for units [c0, c0 + 16), new carry, partial output Dense (ScaleHid folded in)
    float hn[16];
    {
      unsigned ua[16], ub[16];
      tmem_ld16(lane_addr + c0, ua);                       // r
      tmem_ld16(lane_addr + 192 + c0, ub);                 // n_h
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        hn[j] = gate_sigmoid(__uint_as_float(ua[j]) + bi[c0 + j]) * (__uint_as_float(ub[j]) + bhn[c0 + j]);
      tmem_ld16(lane_addr + 128 + c0, ua);                 // n_i
      tmem_ld16(lane_addr + 64 + c0, ub);                  // z
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float cand = gate_tanh(__uint_as_float(ua[j]) + bi[128 + c0 + j] + hn[j]);
        const float zgate = gate_sigmoid(__uint_as_float(ub[j]) + bi[64 + c0 + j]);
        hn[j] = (1.f - zgate) * cand;
        ub[j] = __float_as_uint(zgate);
      }
      tmem_ld16(lane_addr + COL_HH + c0, ua);
      unsigned uc[16];
      tmem_ld16(lane_addr + COL_HL + c0, uc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)                         // previous carry = hi + lo exactly (lo is the remainder)
        hn[j] = fmaf(__uint_as_float(ub[j]), __uint_as_float(ua[j]) + __uint_as_float(uc[j]), hn[j]);
    }
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(net.out_w) + c0 + j);
      o0 = fmaf(hn[j], w.x, o0); o1 = fmaf(hn[j], w.y, o1); o2 = fmaf(hn[j], w.z, o2); o3 = fmaf(hn[j], w.w, o3);
    }
    tmem_st4(lane_addr + COL_OUT + 4 * cg, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
    if (valid) {
      float* hout = g.rnn_out + off + c0;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq)
        *reinterpret_cast<float4*>(hout + 4 * qq) = make_float4(hn[4 * qq], hn[4 * qq + 1], hn[4 * qq + 2], hn[4 * qq + 3]);
    }
    quadrant_exchange(q);
    if (cg == 0) {
      unsigned u[16];
      tmem_ld16(lane_addr + COL_OUT, u);
      tmem_ld_wait();
      if (valid) {
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          o[c] = __ldg(net.out_b + c) + ((__uint_as_float(u[c]) + __uint_as_float(u[4 + c])) +
                                         (__uint_as_float(u[8 + c]) + __uint_as_float(u[12 + c])));
        if (policy) {
          policy_tail(g, o[0], o[1], o[2], o[3], env, slot, ai, n);
        } else {
          float* vo = g.value + ((((size_t)env * g.out_pitch + slot) * nr + ai) * net.n_out);
          for (int c = 0; c < net.n_out; ++c) vo[c] = o[c];
        }
      }
    }
    tc_fence_before();            // TMEM reads done before the next tile's stores / MMAs (ordered by its first barrier)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

size_t head_tc_smem_bytes() { return sizeof(HeadTcSmem); }

// Launches the tensor-core head over all rows of `g` (embeddings expected in rnn_out, as after
// gnn_layers).  tcw: the packed [Dense0 | Dense1 | GRU x | GRU h] hi / lo blocks (DgppoNetLayout.tc_head).
int launch_head_tc(cudaStream_t st, const NetP& P, const GnnArgs& g, const float* tcw, int sms, bool pdl) {
  static bool attr_done[64] = {};               // the attribute is per device: set once per device and process
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t err = cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HeadTcSmem));
    if (err != cudaSuccess) return (int)err;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int nr = (P.kind == DGPPO_NET_VL) ? 1 : g.n;
  const long total_rows = (long)g.n_graphs * nr;
  const long n_tiles = (total_rows + TC_ROWS - 1) / TC_ROWS;
  const int grid = n_tiles < sms ? (int)n_tiles : sms;
  launch_pdl(pdl, head_tc_kernel, grid, TC_THREADS, sizeof(HeadTcSmem), st, P, g, tcw);
  return (int)cudaGetLastError();
}

}  // namespace dgppo
