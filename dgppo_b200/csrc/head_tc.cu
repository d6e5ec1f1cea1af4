// K4 head on the 5th-generation tensor cores (tcgen05, sm_100a): head MLP + LayerNorm + GRU + output
// tails for a tile of 128 rows (agent rows, or one pooled row per graph for Vl)
// (nn/mlp.py:14-30, nn/rnn.py:14-30, algo/module/policy.py:61-74, value.py:15-79).
//
// Precision: every GEMM is a 3xTF32 product (A_hi B_hi + A_lo B_hi + A_hi B_lo) accumulated in fp32 in
// TMEM, so the result stays at fp32 level (measured max |err| 2.5e-6 on |values| <= 2.7 for K = 64,
// tools/micro/tc_tf32_test.cu; the fp32 FMA loop beside it: 0.7e-6) - inside the path's rtol 1e-5.
//
// Mapping (thread = row = TMEM lane; 128 threads, one CTA per SM):
//   TMEM (512 columns): [0,256) accumulators, [256,320) x_hi, [320,384) x_lo, [384,448) h_hi, [448,512) h_lo.
//     The A operands live in TMEM (tcgen05.mma with A from TMEM): the thread that owns a row writes its
//     hi / lo split with tcgen05.st and reads the accumulator row back with tcgen05.ld, so LayerNorm is
//     thread-local and no activation ever goes through shared memory.
//   smem (224 KB): the GRU weights hi + lo, resident for the whole kernel (2 x 96 KB, B operands in the
//     no-swizzle K-major canonical layout, pre-arranged at pack time), plus one 32 KB slot through which
//     the two Dense64 blocks are streamed from L2 by 1-D bulk copies (cp.async.bulk + mbarrier), each
//     fetched while the previous stage's epilogue runs.
//   Per tile: x = embedding, h = carry -> split -> TMEM | MMA Dense0 | LN+ReLU -> split -> TMEM | MMA Dense1 |
//     LN+ReLU -> split -> TMEM | MMA x [Wir Wiz Win] (N = 192), h [Whr Whz] (N = 128, accumulated onto r, z),
//     h Whn (N = 64) | gates, new carry, output Dense (64 x 4, FFMA), TanhNormal sample / log-prob or value.
//   One elected thread issues the MMAs (72 + 24 + 24 per tile) and commits them to an mbarrier the CTA
//   waits on.
#include "gnn_common.cuh"
#include "tc_common.cuh"

namespace dgppo {

using namespace tc;

namespace {

constexpr int TC_ROWS = 128;
constexpr unsigned COL_XH = 256, COL_XL = 320, COL_HH = 384, COL_HL = 448;
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

// 16 fp32 values of this thread's row -> hi / lo TF32 split -> TMEM columns [col_hi + c0, +16), [col_lo + c0, +16).
// lo keeps the exact remainder (hi + lo == v in fp32; the tensor core reads its top 19 bits).
__device__ __forceinline__ void split_store16(unsigned lane_addr, unsigned col_hi, unsigned col_lo, int c0,
                                              const float (&v)[16]) {
  unsigned hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float h = tf32_rna(v[j]);
    hi[j] = __float_as_uint(h);
    lo[j] = __float_as_uint(v[j] - h);
  }
  tmem_st16(lane_addr + col_hi + c0, hi);
  tmem_st16(lane_addr + col_lo + c0, lo);
}

// one global row of 64 floats -> TMEM (zeros for rows past the end)
__device__ __forceinline__ void row_to_tmem(const float* __restrict__ src, bool valid, unsigned lane_addr,
                                            unsigned col_hi, unsigned col_lo) {
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = valid ? *reinterpret_cast<const float4*>(src + c0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    split_store16(lane_addr, col_hi, col_lo, c0, v);
  }
}

// Dense64 accumulator row (TMEM columns [0,64)) + bias -> LayerNorm (flax: eps 1e-6, fast variance) -> ReLU
// -> split -> x operand columns.
__device__ __forceinline__ void epilogue_ln(unsigned lane_addr, const float* __restrict__ bias,
                                            const float* __restrict__ scale, const float* __restrict__ shift) {
  float v[64];
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    unsigned u[16];
    tmem_ld16(lane_addr + c0, u);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c0 + j] = __uint_as_float(u[j]) + bias[c0 + j];
  }
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) { s += v[c]; s2 = fmaf(v[c], v[c], s2); }
  const float mean = s * (1.f / HID), mean2 = s2 * (1.f / HID);
  const float var = fmaxf(0.f, mean2 - mean * mean);
  const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = fmaxf((v[c0 + j] - mean) * (rstd * scale[c0 + j]) + shift[c0 + j], 0.f);
    split_store16(lane_addr, COL_XH, COL_XL, c0, y);
  }
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

__global__ void __launch_bounds__(TC_ROWS, 1)
head_tc_kernel(NetP net, GnnArgs g, const float* __restrict__ tcw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  HeadTcSmem& S = *reinterpret_cast<HeadTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
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
      for (int i = tid; i < len[a]; i += TC_ROWS) S.small_p[o + i] = __ldg(src[a] + i);
      o += len[a];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = S.tmem_base;
  const unsigned lane_addr = tmem + ((unsigned)(warp * 32) << 16);
  const float *d0b = S.small_p, *ln0s = d0b + 64, *ln0b = d0b + 128, *d1b = d0b + 192, *ln1s = d0b + 256,
              *ln1b = d0b + 320, *bi = d0b + 384, *bhn = d0b + 576;

  // weights: GRU blocks once (resident), Dense0 of the first tile
  if (tid == 0) {
    mbar_expect_tx(&S.bar_g, 2u * TC_G_FL * 4u);
#pragma unroll
    for (int p = 0; p < 6; ++p) {                          // 32 KB pieces
      bulk_g2s(S.gx + p * 8192, tcw + 2 * TC_D_FL + p * 8192, 32768u, &S.bar_g);
    }
    mbar_expect_tx(&S.bar_d, TC_D_FL * 4u);
    bulk_g2s(S.dslot, tcw, TC_D_FL * 4u, &S.bar_d);
  }
  pdl_wait();                     // weights in flight while gnn_layers drains; its embeddings are visible from here on

  const bool policy = net.kind == DGPPO_NET_POLICY;
  unsigned ph_d = 0, ph_mma = 0;
  bool g_ready = false;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long row = tile * TC_ROWS + tid;
    const bool valid = row < total_rows;
    int env = 0, slot = 0, ai = 0;
    size_t off = 0;
    if (valid) {
      const long gi = row / nr; ai = (int)(row - gi * nr);
      const long e = gi / g.n_slots; env = (int)e; slot = (int)(gi - e * g.n_slots);
      off = (((size_t)env * g.rnn_pitch + slot) * nr + ai) * HID;
    }
    // ---- operands: x = embedding (scratch rows of rnn_out), h = previous carry
    row_to_tmem(g.rnn_out + off, valid, lane_addr, COL_XH, COL_XL);
    row_to_tmem(g.rnn_in + off, valid, lane_addr, COL_HH, COL_HL);
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
    epilogue_ln(lane_addr, d0b, ln0s, ln0b);
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
    epilogue_ln(lane_addr, d1b, ln1s, ln1b);
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
    // ---- gates (flax GRUCell), new carry, output Dense (ScaleHid folded in at pack time)
    float o0 = __ldg(net.out_b), o1 = __ldg(net.out_b + 1), o2 = __ldg(net.out_b + 2), o3 = __ldg(net.out_b + 3);
    float* hout = g.rnn_out + off;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      unsigned ur[16], uz[16], uni[16], unh[16], uhh[16], uhl[16];
      tmem_ld16(lane_addr + c0, ur);
      tmem_ld16(lane_addr + 64 + c0, uz);
      tmem_ld16(lane_addr + 128 + c0, uni);
      tmem_ld16(lane_addr + 192 + c0, unh);
      tmem_ld16(lane_addr + COL_HH + c0, uhh);
      tmem_ld16(lane_addr + COL_HL + c0, uhl);
      tmem_ld_wait();
      float hn[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        const float rgate = gate_sigmoid(__uint_as_float(ur[j]) + bi[c]);
        const float zgate = gate_sigmoid(__uint_as_float(uz[j]) + bi[64 + c]);
        const float cand = gate_tanh(__uint_as_float(uni[j]) + bi[128 + c] + rgate * (__uint_as_float(unh[j]) + bhn[c]));
        const float hprev = __uint_as_float(uhh[j]) + __uint_as_float(uhl[j]);      // exact: lo is the remainder
        hn[j] = (1.f - zgate) * cand + zgate * hprev;
        const float4 w = __ldg(reinterpret_cast<const float4*>(net.out_w) + c);
        o0 = fmaf(hn[j], w.x, o0); o1 = fmaf(hn[j], w.y, o1); o2 = fmaf(hn[j], w.z, o2); o3 = fmaf(hn[j], w.w, o3);
      }
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(hout + c0 + 4 * q) = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
      }
    }
    if (valid) {
      if (policy) {
        policy_tail(g, o0, o1, o2, o3, env, slot, ai, n);
      } else {
        float* vo = g.value + ((((size_t)env * g.out_pitch + slot) * nr + ai) * net.n_out);
        const float o[4] = {o0, o1, o2, o3};
        for (int c = 0; c < net.n_out; ++c) vo[c] = o[c];
      }
    }
    tc_fence_before();            // accumulator reads done before the next tile's MMAs (ordered by its first barrier)
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
  launch_pdl(pdl, head_tc_kernel, grid, TC_ROWS, sizeof(HeadTcSmem), st, P, g, tcw);
  return (int)cudaGetLastError();
}

}  // namespace dgppo
