// Dense update / combine units on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
// Included by dense.cu (shares Group<>, stage_desc, the BatchNorm merges and the activation helpers).
//
// Reference: mp/layers.py:191-199, 303-325 (Linear -> BatchNorm -> act nets after propagate) and :290-293 (the
// coboundary message Linear, applied here as two per-cell products).
//
// Precision. fp32 parity (rtol 1e-5) rules out plain TF32 (10-bit mantissa). Every operand is split in registers into
// hi = tf32(v), lo = tf32(v - hi) (3xTF32 and the lo*lo term on top). What limits accuracy then is the tensor core's
// accumulator: it truncates (rounds toward zero) at every accumulation, a biased error that a chain of 24
// accumulations makes 3-4x larger than a sequential fp32 FMA chain (measured: profiles/r2_tc5_probe_accumulation.txt;
// the round-1 mma.sync variant failed the training-step parity test for this reason). So (a) the small cross terms
// never share an accumulator with hi*hi, and (b) the hi*hi chain is cut into `nb` k-ranges with one accumulator each,
// which the epilogue adds in fp32 round-to-nearest. With nb >= 2 the result is MORE accurate than the FMA chain (rms
// error 0.85x at nb = 2, 0.65x at nb = 4, K = 64); TMEM (512 columns) is what pays for it.
//
// One instruction per k-step. A tcgen05.mma costs ~80 cycles to issue whatever its shape up to 128 x 128 x 8
// (profiles/r2_tc5_probe_issue_latency.txt), and the issuing thread is the serial bottleneck of tiles this small. So
// the hi and lo parts are CONCATENATED along the operand's M / N extent instead of being issued as three products:
//     [A_hi ; A_lo] (M = 128)  x  [B_hi ; B_lo] (N = 2 n)   ->   D = | hi*hi  hi*lo |   rows 0..63    (TMEM lanes 0..63)
//                                                                   | lo*hi  lo*lo |   rows 64..127
// all four terms, each in its own accumulator cells, from ONE instruction (K/8 instructions per product instead of
// 3 K/8). The epilogue adds the two column halves per lane, stages the 128 rows in shared memory and adds row r + 64
// to row r there.
//
// Tiles. A CTA owns 64 rows of the unit. Operands pass through registers exactly once (the previous unit's BatchNorm +
// activation is applied there, then the split) and are written to shared memory in the two layouts of tc5.cuh with
// conflict-free 16-byte stores:
//   forward   z = f(X) W^T       : A = f(X) [64 x K] K-major, B = W [h x K] K-major                          N = 2 h
//   backward  g_in = g_z W       : A = g_z [64 x h] K-major,  B = W [h (inner) x K] MN-major                 N = 2 K
//             g_W  = g_z^T f(X)  : A = g_z [64 (inner) x h] MN-major, B = f(X) [64 (inner) x K] MN-major     M = 2 h, N = 2 K
// One thread issues the MMAs, one tcgen05.commit arrives on an mbarrier, and the epilogue reads TMEM with
// tcgen05.ld (32x32b; M = 128: D row i is TMEM lane i), stages the tile in shared memory (the operand buffers are dead
// by then) and leaves with coalesced 128-bit stores; BatchNorm partials are taken from the staged tile.
#pragma once
#include "tc5.cuh"

namespace cwn {

constexpr int T5R = 64;    // rows per CTA tile
constexpr int T5T = 256;   // threads per CTA

// number of k-range accumulators, given how many the TMEM budget allows and the inner extent; every accumulator
// receives at least one k-step
__host__ __device__ __forceinline__ int t5_blocks(int want, int inner) {
  const int ksteps = inner / 8;
  int nb = want < ksteps ? want : ksteps;
  if (nb < 1) nb = 1;
  const int per = (ksteps + nb - 1) / nb;
  return (ksteps + per - 1) / per;
}
__host__ __device__ __forceinline__ uint32_t t5_pow2_cols(int cols) {
  uint32_t c = 32;
  while ((int)c < cols) c <<= 1;
  return c;
}

// Issue D = A_cat * B_cat over `inner` (one instruction per k-step of 8), k-range block b into the accumulator at
// d_tmem + b * N. ONE thread. Lean by construction: a first version that rebuilt four descriptors and divided at every
// step spent ~120 cycles per MMA in integer arithmetic; descriptors now advance by adding to their start-address
// field (bits 0..13, units of 16 bytes).
__device__ __forceinline__ void t5_issue(uint32_t d_tmem, int N, int nb, uint32_t a, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step,
                                         uint32_t a_layout, uint32_t b, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step,
                                         uint32_t b_layout, int inner, uint32_t idesc) {
  const int ksteps = inner >> 3, per = (ksteps + nb - 1) / nb;
  uint64_t da = tc5::smem_desc(a, a_lbo, a_sbo, a_layout), db = tc5::smem_desc(b, b_lbo, b_sbo, b_layout);
  const uint64_t sa = a_step >> 4, sb = b_step >> 4;
  int ks = 0;
  for (int blk = 0; blk < nb; ++blk, d_tmem += (uint32_t)N) {
    const int end = (ks + per < ksteps) ? ks + per : ksteps;
    for (uint32_t acc = 0; ks < end; ++ks, acc = 1u, da += sa, db += sb) tc5::mma_tf32(d_tmem, da, db, idesc, acc);
  }
}

// v[0..16) = SUM over the nb accumulators (`stride` columns apart) of 16 consecutive columns of this thread's TMEM
// lane: two loads in flight per wait (register budget: the forward kernel runs two CTAs per SM), blocks added in order
__device__ __forceinline__ void t5_sum16(uint32_t taddr, int stride, int nb, float (&v)[16]) {
  float b1[16];
  tc5::tmem_ld16(taddr, v);
  if (nb > 1) tc5::tmem_ld16(taddr + (uint32_t)stride, b1);
  tc5::tmem_ld_wait();
  if (nb > 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b1[j];
  }
  for (int b = 2; b < nb; ++b) {
    tc5::tmem_ld16(taddr + (uint32_t)(b * stride), b1);
    tc5::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b1[j];
  }
}

// One lane's row of a concatenated product, columns [c0, c0 + 16) of the logical result: the (x*hi) half summed over
// its k-range blocks, then the (x*lo) half; written to the staging tile T[128][ldt] (row = TMEM lane).
__device__ __forceinline__ void t5_stage16(uint32_t tmem_lane, int n, int nb, int c0, float* trow) {
  float a[16], b[16];
  t5_sum16(tmem_lane + (uint32_t)c0, 2 * n, nb, a);
  t5_sum16(tmem_lane + (uint32_t)(n + c0), 2 * n, nb, b);
#pragma unroll
  for (int e = 0; e < 16; e += 4)
    *reinterpret_cast<float4*>(trow + c0 + e) = make_float4(a[e] + b[e], a[e + 1] + b[e + 1], a[e + 2] + b[e + 2], a[e + 3] + b[e + 3]);
}

__device__ __forceinline__ float4 t5_transform(const float4 v, const float4 mu, const float4 sc, const float4 be, int act_code,
                                               int A) {
  float4 o;
  if (A == CWN_ACT_ID) {
    o.x = (v.x - mu.x) * sc.x + be.x; o.y = (v.y - mu.y) * sc.y + be.y; o.z = (v.z - mu.z) * sc.z + be.z; o.w = (v.w - mu.w) * sc.w + be.w;
  } else if (A == CWN_ACT_RELU) {
    o.x = fmaxf((v.x - mu.x) * sc.x + be.x, 0.f); o.y = fmaxf((v.y - mu.y) * sc.y + be.y, 0.f);
    o.z = fmaxf((v.z - mu.z) * sc.z + be.z, 0.f); o.w = fmaxf((v.w - mu.w) * sc.w + be.w, 0.f);
  } else {
    o.x = act_apply_rt(act_code, (v.x - mu.x) * sc.x + be.x); o.y = act_apply_rt(act_code, (v.y - mu.y) * sc.y + be.y);
    o.z = act_apply_rt(act_code, (v.z - mu.z) * sc.z + be.z); o.w = act_apply_rt(act_code, (v.w - mu.w) * sc.w + be.w);
  }
  return o;
}

// One thread's 4-column chunk of a unit's input f_in(X) (virtual concat [X0 | X1] + the previous unit's BatchNorm /
// activation). The column is the same for every row a thread touches (256 threads, K/4 a power of two), so the block
// selection and the transform vectors are resolved ONCE; raw loads are separated from the transform so that a whole
// batch of 128-bit loads is in flight before the first is consumed.
template <class D>
struct T5InCol {
  const float* base;
  int64_t ld;
  float4 mu, sc, be;
  bool tf;
  int act;
  __device__ __forceinline__ T5InCol(const D& d, int c, int64_t row0, bool transform) {
    const bool first = c < d.k0;
    base = first ? d.x0 + row0 * d.ld_x0 + c : d.x1 + row0 * d.ld_x1 + (c - d.k0);
    ld = first ? d.ld_x0 : d.ld_x1;
    mu = make_float4(0.f, 0.f, 0.f, 0.f); sc = make_float4(1.f, 1.f, 1.f, 1.f); be = mu;
    tf = transform; act = d.in_act;
    const float* sp = first ? d.in_scale0 : d.in_scale1;
    if (transform && sp) {
      const int cc = first ? c : c - d.k0;
      mu = ldg_f4((first ? d.in_mean0 : d.in_mean1) + cc);
      sc = ldg_f4(sp + cc);
      be = ldg_f4((first ? d.in_beta0 : d.in_beta1) + cc);
    }
  }
  __device__ __forceinline__ float4 raw(int r, int rows) const {  // zero past the matrix
    return r < rows ? ldg_f4(base + (int64_t)r * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  template <int A_IN>
  __device__ __forceinline__ float4 apply(float4 v, int r, int rows) const {  // padding rows stay zero
    return (tf && r < rows) ? t5_transform(v, mu, sc, be, act, A_IN) : v;
  }
};

__device__ __forceinline__ int t5_log2(int v) { return 31 - __clz(v); }

// ------------------------------------------------------------------------------------------------ forward
struct T5FwdSmem {  // byte offsets into dynamic shared memory (host and device agree through this one function)
  uint32_t a, b, total;   // A_cat = [f(X)_hi ; f(X)_lo] tiled over 128 rows, B_cat = [W_hi ; W_lo] tiled over 2 h rows
  int nb;                 // k-range accumulators of 2 h columns each
  uint32_t tmem_cols;
  __host__ __device__ T5FwdSmem(int K, int h) {
    const uint32_t ab = tc5::Tiled::bytes(2 * T5R, K), bb = tc5::Tiled::bytes(2 * h, K);
    a = 0; b = ab;
    total = ab + bb;
    const uint32_t stage = 2u * T5R * (uint32_t)(h + 4) * 4u;  // 128 staged rows over the dead operand buffers
    if (total < stage) total = stage;
    int fit = 512 / (2 * h);            // accumulators that fit TMEM
    int want = K <= 64 ? 2 : 4;         // (K <= 64: 2 keeps h = 64 at 256 columns, i.e. two CTAs per SM)
    if (want > fit) want = fit;
    nb = t5_blocks(want, K);
    tmem_cols = t5_pow2_cols(nb * 2 * h);
  }
};

template <int A_IN>
__global__ void __launch_bounds__(T5T, 2) linear_fwd_tc5_kernel(const __grid_constant__ Group<cwn_linear_desc> g) {
  extern __shared__ __align__(1024) unsigned char smem5[];
  __shared__ cwn_linear_desc sd;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ float red[T5T];
  __shared__ float meanv[128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  if (tid == 0) { tc5::mbar_init(&bar, 1); tc5::mbar_fence_init(); }
  const cwn_linear_desc d = stage_desc(g, p, &sd);  // (contains the CTA barrier that publishes the mbarrier init)
  pdl_trigger();  // the next kernel of the stream may start its prologue
  pdl_wait();     // ... and this one goes no further until its predecessor has completed (common.cuh)
  CWN_PHASE(1);
  const int rt = blockIdx.x - g.start[p];
  const int K = d.k0 + d.k1, K4 = K >> 2, h = d.h;
  const T5FwdSmem L(K, h);
  const int nb = L.nb;
  const int64_t row0 = (int64_t)rt * T5R;
  const int rows = (int)((d.n_rows - row0 < T5R) ? d.n_rows - row0 : T5R);
  const int64_t n_live = live_rows(d);  // rows past it pad a fixed-capacity batch: computed, but outside the statistics
  const int rows_live = (int)((n_live - row0 < rows) ? (n_live - row0 > 0 ? n_live - row0 : 0) : rows);
  const bool transform = d.in_scale0 || d.in_scale1 || d.in_act != CWN_ACT_ID;
  const tc5::Tiled ta(2 * T5R), tb(2 * h);
  {  // operands: global -> registers (transform, split) -> shared memory. Thread -> column chunk c4 (fixed) and rows
     // rb, rb + rs, ...; per trip up to 4 rows of X and 4 rows of W are requested before any is consumed.
    const int lg = t5_log2(K4), c4 = tid & (K4 - 1), rb = tid >> lg, rs = T5T >> lg;
    const T5InCol<cwn_linear_desc> col(d, c4 * 4, row0, transform);
    const float* wcol = d.w + c4 * 4;
    const int r_end = h > T5R ? h : T5R;
    for (int r0 = rb; r0 < r_end; r0 += 4 * rs) {
      float4 va[4], vw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j * rs;
        va[j] = col.raw(r, rows);
        vw[j] = r < h ? ldg_f4(wcol + (int64_t)r * d.ld_w) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (warp == 0 && r0 == rb) tc5::tmem_alloc(&tmem_s, L.tmem_cols);  // (~330 cycles: under the loads just requested)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j * rs;
        float4 hi, lo;
        if (r < T5R) {
          tc5::split_tf32x4(col.template apply<A_IN>(va[j], r, rows), hi, lo);
          *reinterpret_cast<float4*>(smem5 + L.a + ta.off(r, c4)) = hi;
          *reinterpret_cast<float4*>(smem5 + L.a + ta.off(r + T5R, c4)) = lo;
        }
        if (r < h) {
          tc5::split_tf32x4(vw[j], hi, lo);
          *reinterpret_cast<float4*>(smem5 + L.b + tb.off(r, c4)) = hi;
          *reinterpret_cast<float4*>(smem5 + L.b + tb.off(r + h, c4)) = lo;
        }
      }
    }
  }
  tc5::fence_async_smem();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  CWN_PHASE(2);
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    const uint32_t s0 = tc5::smem_u32(smem5);
    t5_issue(tmem, 2 * h, nb, s0 + L.a, ta.s_c, ta.s_r, 2 * ta.s_c, 0, s0 + L.b, tb.s_c, tb.s_r, 2 * tb.s_c, 0, K,
             tc5::idesc_tf32(2 * T5R, 2 * h, 0, 0));
    tc5::mma_commit(&bar);
  }
  tc5::mbar_wait(&bar, 0);
  tc5::fence_after_sync();
  CWN_PHASE(3);
  // epilogue: TMEM -> staged rows T[128][h + 4] over the operand buffers (all MMAs have completed); row r of the unit
  // is T[r] (hi*hi + hi*lo) + T[r + 64] (lo*hi + lo*lo)
  float* T = reinterpret_cast<float*>(smem5);
  const int ldy = h + 4;
  {
    const int q = warp & 3, half = warp >> 2;     // TMEM lane quadrant; the two warps of a quadrant split the columns
    const int per_half = h >= 32 ? (h >> 1) : h;  // (h = 16: one warp per quadrant does it all)
    const int c_lo = half * per_half, c_hi = (c_lo + per_half < h) ? c_lo + per_half : h;
    const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
    for (int c0 = c_lo; c0 < c_hi; c0 += 16)      // (warp-uniform trip count: tcgen05.ld is warp-collective)
      t5_stage16(tl, h, nb, c0, T + (32 * q + lane) * ldy);
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, L.tmem_cols);
  CWN_PHASE(4);
  {  // combine the two row halves, add the bias, leave with coalesced stores; the tile stays staged for the statistics
    const int h4 = h >> 2, lgh = t5_log2(h4);
    for (int i = tid; i < rows * h4; i += T5T) {
      const int r = i >> lgh, c = (i & (h4 - 1)) * 4;
      float4 v = f4_add(*reinterpret_cast<const float4*>(T + r * ldy + c), *reinterpret_cast<const float4*>(T + (r + T5R) * ldy + c));
      if (d.bias) v = f4_add(v, ldg_f4(d.bias + c));
      *reinterpret_cast<float4*>(T + r * ldy + c) = v;
      *reinterpret_cast<float4*>(d.z + (row0 + r) * d.ld_z + c) = v;
    }
  }
  CWN_PHASE(5);
  if (!d.stats) {
    if (d.bn_mean && !d.bn_training && blockIdx.x == g.start[p])  // eval: statistics are the running ones
      bn_finalize_body(nullptr, 0, d.n_rows, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 0, d.bn_running_mean,
                       d.bn_running_var, nullptr, d.bn_mean, d.bn_scale, d.bn_rstd, nullptr);
    return;
  }
  __syncthreads();
  if (rows_live > 0) {  // per-column (mean, M2) of this tile's live rows: 256 / h row groups per column, fixed order
    const int parts = T5T / h, c = tid % h, part = tid / h;
    float s = 0.f;
    for (int r = part; r < rows_live; r += parts) s += T[r * ldy + c];
    red[tid] = s;
    __syncthreads();
    if (tid < h) {
      float tot = 0.f;
      for (int q = 0; q < parts; ++q) tot += red[q * h + tid];
      meanv[tid] = tot / (float)rows_live;
    }
    __syncthreads();
    const float mu = meanv[c];
    float m2 = 0.f;
    for (int r = part; r < rows_live; r += parts) {
      const float dv = T[r * ldy + c] - mu;
      m2 = fmaf(dv, dv, m2);
    }
    red[tid] = m2;
    __syncthreads();
    if (tid < h) {
      float tot = 0.f;
      for (int q = 0; q < parts; ++q) tot += red[q * h + tid];
      d.stats[((int64_t)rt * 2 + 0) * h + tid] = meanv[tid];
      d.stats[((int64_t)rt * 2 + 1) * h + tid] = tot;
    }
  }
  CWN_PHASE(6);
  if (d.bn_mean && d.counter) {
    const int total = g.start[p + 1] - g.start[p];
    const bool last_ = last_cta_of_problem(d.counter, total);
    CWN_PHASE(7);
    if (last_) {
      const int n_tiles = (int)((n_live + T5R - 1) / T5R);  // (tiles without a live row wrote no record)
      bn_finalize_body(d.stats, n_tiles, n_live, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 1, d.bn_running_mean,
                       d.bn_running_var, d.bn_num_batches_tracked, d.bn_mean, d.bn_scale, d.bn_rstd,
                       reinterpret_cast<float*>(smem5), (int)(L.total / 4), T5R);
      if (threadIdx.x == 0) *d.counter = 0;
      CWN_PHASE(8);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward (h = 64)
struct T5BwdSmem {
  // gm = [g_z_hi | g_z_lo] MN-major (inner = rows, M = 2 h); xm = [f(X)_hi | f(X)_lo] MN-major (inner = rows, N = 2 K);
  // gt = [g_z_hi ; g_z_lo] K-major tiled over 128 rows; wm = [W_hi | W_lo] MN-major (inner = h, N = 2 K)
  uint32_t gm, xm, gt, wm, vout, bsum, npart, total;
  uint32_t g_lbo, x_lbo, w_lbo;  // MN-major: stride between 32-column groups (sbo = 512 everywhere)
  int nb1, nb2;
  uint32_t col2, tmem_cols;
  __host__ __device__ T5BwdSmem(int K, int h) {
    g_lbo = x_lbo = 512u * (T5R / 4);
    w_lbo = 512u * (uint32_t)(h / 4);
    const uint32_t gmb = 2u * (uint32_t)(h / 32) * g_lbo, xmb = 2u * (uint32_t)(K / 32) * x_lbo, wmb = 2u * (uint32_t)(K / 32) * w_lbo;
    const uint32_t gtb = (tc5::Tiled::bytes(2 * T5R, h) + 1023u) & ~1023u;
    gm = 0; xm = gmb; gt = gmb + xmb; wm = gt + gtb;   // (the staging tile T[128][K + 4] lives over gm + xm + gt)
    vout = wm + wmb;                                   // [6][h] floats
    bsum = vout + 6u * (uint32_t)h * 4u;               // [256 / (h/4)][h] floats
    npart = bsum + (uint32_t)(T5T / (h / 4)) * (uint32_t)h * 4u;  // [2][256 / (K/4)][K] floats: fused upstream reduction
    total = npart + 2u * (uint32_t)T5T * 4u * 4u;
    nb1 = t5_blocks(K <= 64 ? 2 : 1, h);               // g_in: accumulators of 2 K columns
    nb2 = 1;                                           // g_W (only rtol 1e-4 is asked of parameter gradients)
    col2 = (uint32_t)(nb1 * 2 * K);
    tmem_cols = t5_pow2_cols((int)col2 + nb2 * 2 * K);
  }
};

template <int A_IN, int A_OUT>
__global__ void __launch_bounds__(T5T, 1) unit_bwd_tc5_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  extern __shared__ __align__(1024) unsigned char smem5[];
  __shared__ cwn_unit_bwd_desc sd;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  if (tid == 0) { tc5::mbar_init(&bar, 1); tc5::mbar_fence_init(); }
  const cwn_unit_bwd_desc d = stage_desc(g, p, &sd);
  pdl_trigger();
  pdl_wait();
  CWN_PHASE(1);
  const int j = blockIdx.x - g.start[p];
  const int K = d.k0 + d.k1, K4 = K >> 2, h = d.h, h4 = h >> 2;
  const bool want_gin = d.g_in0 || d.g_in1;
  const T5BwdSmem L(K, h);
  const tc5::Tiled tg(2 * T5R);
  const int n_tiles = (int)((d.n_rows + T5R - 1) / T5R);
  const int64_t n_live = live_rows(d);  // rows past it pad a fixed-capacity batch: their g_z is zero
  float* wpart = d.w_partials + (int64_t)j * h * K;
  float* bpart = d.b_partials + (int64_t)j * h;
  const bool transform = d.in_scale0 || d.in_scale1 || d.in_act != CWN_ACT_ID;
  float* vout = reinterpret_cast<float*>(smem5 + L.vout);
  float* bsum = reinterpret_cast<float*>(smem5 + L.bsum);
  for (int c = tid; c < h; c += T5T) {
    const bool live = d.has_bn;
    vout[c] = live ? __ldg(d.mean + c) : 0.f;
    vout[h + c] = live ? __ldg(d.scale + c) : 1.f;
    vout[2 * h + c] = live ? __ldg(d.rstd + c) : 0.f;
    vout[3 * h + c] = (live && d.beta) ? __ldg(d.beta + c) : 0.f;
    vout[4 * h + c] = live ? __ldcg(d.c1 + c) : 0.f;
    vout[5 * h + c] = live ? __ldcg(d.c2 + c) : 0.f;
  }
  const int lgk = t5_log2(K4), c4x = tid & (K4 - 1), rbx = tid >> lgk, rsx = T5T >> lgk;  // this thread's chunk of X / W rows
  const uint32_t lo_x = (uint32_t)(K / 32) * L.x_lbo, lo_w = (uint32_t)(K / 32) * L.w_lbo, lo_g = (uint32_t)(h / 32) * L.g_lbo;
  {  // W [h (inner) x K] -> MN-major swizzled, once per CTA
    const float* wcol = d.w + c4x * 4;
    for (int r0 = rbx; r0 < h; r0 += 4 * rsx) {
      float4 v[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + jj * rsx;
        v[jj] = r < h ? ldg_f4(wcol + (int64_t)r * d.ld_w) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (warp == 0 && r0 == rbx) tc5::tmem_alloc(&tmem_s, L.tmem_cols);  // (~330 cycles: under the loads just requested)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + jj * rsx;
        if (r >= h) continue;
        float4 hi, lo;
        tc5::split_tf32x4(v[jj], hi, lo);
        const uint32_t o = tc5::mn_off(r, c4x, L.w_lbo, 512u);
        *reinterpret_cast<float4*>(smem5 + L.wm + o) = hi;
        *reinterpret_cast<float4*>(smem5 + L.wm + lo_w + o) = lo;
      }
    }
  }
  __syncthreads();  // vout ready
  const uint32_t tl = ((uint32_t)(32 * (warp & 3)) << 16);  // this warp's TMEM lane quadrant
  const int R = 32 * (warp & 3) + lane, half = warp >> 2;   // this thread's accumulator row; the quadrant's two warps split the columns
  float* T = reinterpret_cast<float*>(smem5);               // staging [128][K + 4] over gm / xm / gt
  const int ldt = K + 4;
  uint32_t parity = 0;
  bool first = true;
  // ---- fused BatchNorm-backward reduction of the upstream units (cwn_unit_bwd_desc::next_*): this thread's column chunk
  //      c4x of the g_in tile belongs to input block nhf; the upstream unit's BatchNorm + activation are this unit's input
  //      transform, its z is this unit's raw input
  const bool any_next = d.next_red0 || d.next_red1;
  const int nkq = c4x * 4, nhf = nkq < d.k0 ? 0 : 1, nkc = nhf ? nkq - d.k0 : nkq;
  const bool do_next = nhf ? d.next_red1 != nullptr : d.next_red0 != nullptr;
  float4 n_mean = make_float4(0.f, 0.f, 0.f, 0.f), n_scale = make_float4(1.f, 1.f, 1.f, 1.f), n_beta = n_mean, n_rstd = n_mean;
  const float* n_x = nullptr;
  int64_t n_ldx = 0;
  if (do_next) {
    n_mean = ldg_f4((nhf ? d.in_mean1 : d.in_mean0) + nkc);
    n_scale = ldg_f4((nhf ? d.in_scale1 : d.in_scale0) + nkc);
    const float* bp = nhf ? d.in_beta1 : d.in_beta0;
    if (bp) n_beta = ldg_f4(bp + nkc);
    n_rstd = ldg_f4((nhf ? d.next_rstd1 : d.next_rstd0) + nkc);
    n_x = (nhf ? d.x1 : d.x0) + nkc;
    n_ldx = nhf ? d.ld_x1 : d.ld_x0;
  }
  float* npart = reinterpret_cast<float*>(smem5 + L.npart);
  for (int tile = j; tile < n_tiles; tile += d.n_ctas, first = false, parity ^= 1u) {
    const int64_t row0 = (int64_t)tile * T5R;
    const int rows_cap = (int)((d.n_rows - row0 < T5R) ? d.n_rows - row0 : T5R);
    const int rows = (int)((n_live - row0 < rows_cap) ? (n_live - row0 > 0 ? n_live - row0 : 0) : rows_cap);
    if (first) CWN_PHASE(2);
    {  // g_z = scale * (g_out act'(y) - c1 - zhat c2)  (plain g_out act'(z) without BatchNorm): registers -> both layouts;
       // f_in(X) -> MN-major. Thread -> (column chunk q of z / g_out, rows rbg + j rsg) and (chunk c4x of X, rows rbx + j rsx);
       // the loads of a trip (z, g_out and X rows) are all requested before the first is consumed.
      const int lgh = t5_log2(h4), q = tid & (h4 - 1), c = q * 4, rbg = tid >> lgh, rsg = T5T >> lgh;
      const float4 mu = *reinterpret_cast<const float4*>(vout + c), sc = *reinterpret_cast<const float4*>(vout + h + c);
      const float4 rs = *reinterpret_cast<const float4*>(vout + 2 * h + c), be = *reinterpret_cast<const float4*>(vout + 3 * h + c);
      const float4 c1 = *reinterpret_cast<const float4*>(vout + 4 * h + c), c2 = *reinterpret_cast<const float4*>(vout + 5 * h + c);
      const T5InCol<cwn_unit_bwd_desc> col(d, c4x * 4, row0, transform);
      const float* zcol = d.z + row0 * d.ld_z + c;
      const float* gcol = d.g_out + row0 * d.ld_g + c;
      float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
      constexpr int NB = 4;
      for (int t0 = 0; t0 * rsg + rbg < T5R || t0 * rsx + rbx < T5R; t0 += NB) {
        float4 zv[NB], gv[NB], xv[NB];
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          const int r = rbg + (t0 + jj) * rsg, rx = rbx + (t0 + jj) * rsx;
          zv[jj] = gv[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) {
            zv[jj] = ldg_f4(zcol + (int64_t)r * d.ld_z);
            gv[jj] = ldg_f4(gcol + (int64_t)r * d.ld_g);
          }
          xv[jj] = col.raw(rx, rows);
        }
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          const int r = rbg + (t0 + jj) * rsg, rx = rbx + (t0 + jj) * rsx;
          float4 hi, lo;
          if (r < T5R) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
              const float zz[4] = {zv[jj].x, zv[jj].y, zv[jj].z, zv[jj].w}, gg[4] = {gv[jj].x, gv[jj].y, gv[jj].z, gv[jj].w};
              const float m_[4] = {mu.x, mu.y, mu.z, mu.w}, s_[4] = {sc.x, sc.y, sc.z, sc.w}, r_[4] = {rs.x, rs.y, rs.z, rs.w};
              const float b_[4] = {be.x, be.y, be.z, be.w}, c1_[4] = {c1.x, c1.y, c1.z, c1.w}, c2_[4] = {c2.x, c2.y, c2.z, c2.w};
              float ov[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float zc = zz[e] - m_[e];
                const float gy = gg[e] * act_grad<A_OUT>(d.act, zc * s_[e] + b_[e]);
                ov[e] = d.has_bn ? s_[e] * (gy - c1_[e] - zc * r_[e] * c2_[e]) : gy;
              }
              o = make_float4(ov[0], ov[1], ov[2], ov[3]);
              colsum = f4_add(colsum, o);
            }
            tc5::split_tf32x4(o, hi, lo);
            const uint32_t om = tc5::mn_off(r, q, L.g_lbo, 512u);
            *reinterpret_cast<float4*>(smem5 + L.gm + om) = hi;
            *reinterpret_cast<float4*>(smem5 + L.gm + lo_g + om) = lo;
            *reinterpret_cast<float4*>(smem5 + L.gt + tg.off(r, q)) = hi;
            *reinterpret_cast<float4*>(smem5 + L.gt + tg.off(r + T5R, q)) = lo;
          }
          if (rx < T5R) {
            tc5::split_tf32x4(col.template apply<A_IN>(xv[jj], rx, rows), hi, lo);
            const uint32_t o = tc5::mn_off(rx, c4x, L.x_lbo, 512u);
            *reinterpret_cast<float4*>(smem5 + L.xm + o) = hi;
            *reinterpret_cast<float4*>(smem5 + L.xm + lo_x + o) = lo;
          }
        }
      }
      *reinterpret_cast<float4*>(bsum + rbg * h + c) = colsum;
    }
    tc5::fence_async_smem();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    if (first) CWN_PHASE(3);
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
      const uint32_t s0 = tc5::smem_u32(smem5);
      // g_W [2 h x 2 K] = [g_z_hi | g_z_lo]^T (A MN-major, inner = rows) * [f(X)_hi | f(X)_lo] (B MN-major)
      t5_issue(tmem + L.col2, 2 * K, L.nb2, s0 + L.gm, L.g_lbo, 512u, 1024u, 1, s0 + L.xm, L.x_lbo, 512u, 1024u, 1, T5R,
               tc5::idesc_tf32(2 * h, 2 * K, 1, 1));
      if (want_gin)  // g_in [128 x 2 K] = [g_z_hi ; g_z_lo] (K-major) * [W_hi | W_lo] (MN-major), inner = h
        t5_issue(tmem, 2 * K, L.nb1, s0 + L.gt, tg.s_c, tg.s_r, 2 * tg.s_c, 0, s0 + L.wm, L.w_lbo, 512u, 1024u, 1, h,
                 tc5::idesc_tf32(2 * T5R, 2 * K, 0, 1));
      tc5::mma_commit(&bar);
    }
    if (tid < h) {  // bias-gradient partial = column sums of g_z (row groups in order), while the tensor core works
      float s_ = 0.f;
      for (int m = 0; m < T5T / h4; ++m) s_ += bsum[m * h + tid];  // (thread group m summed rows m, m + 256 / h4, ...)
      bpart[tid] = first ? s_ : bpart[tid] + s_;
    }
    tc5::mbar_wait(&bar, parity);
    tc5::fence_after_sync();
    if (first) CWN_PHASE(4);
    const int k_lo = half * (K >> 1), k_hi = k_lo + (K >> 1);
    // ---- epilogue A: weight-gradient partial. T[R] = the lane's (x*hi + x*lo); g_W[c] = T[c] + T[c + 64]
    for (int k0 = k_lo; k0 < k_hi; k0 += 16) t5_stage16(tmem + tl + L.col2, K, L.nb2, k0, T + R * ldt);
    tc5::fence_before_sync();
    __syncthreads();
    for (int i = tid; i < h * K4; i += T5T) {
      const int c = i >> lgk, k = (i & (K4 - 1)) * 4;
      float4 v = f4_add(*reinterpret_cast<const float4*>(T + c * ldt + k), *reinterpret_cast<const float4*>(T + (c + h) * ldt + k));
      float4* dst = reinterpret_cast<float4*>(wpart + (int64_t)c * K + k);
      if (!first) v = f4_add(*dst, v);
      *dst = v;
    }
    // ---- epilogue B: input gradient, the same way; coalesced (accumulating) stores
    if (want_gin) {
      __syncthreads();  // T is free again
      tc5::fence_after_sync();
      for (int k0 = k_lo; k0 < k_hi; k0 += 16) t5_stage16(tmem + tl, K, L.nb1, k0, T + R * ldt);
      tc5::fence_before_sync();
      __syncthreads();
      float4 ns1 = make_float4(0.f, 0.f, 0.f, 0.f), ns2 = ns1;
      constexpr int NX = 4;  // rows of a thread per trip: the upstream z values of a trip are requested together
      for (int i0 = tid; i0 < rows_cap * K4; i0 += NX * T5T) {  // (padding rows leave as the zeros their g_z produced)
        float4 xv[NX];
#pragma unroll
        for (int jj = 0; jj < NX; ++jj) {
          const int i = i0 + jj * T5T;
          xv[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (do_next && i < rows_cap * K4) xv[jj] = ldg_f4(n_x + (row0 + (i >> lgk)) * n_ldx);
        }
#pragma unroll
        for (int jj = 0; jj < NX; ++jj) {
          const int i = i0 + jj * T5T;
          if (i >= rows_cap * K4) break;
          const int rr = i >> lgk, kq = (i & (K4 - 1)) * 4;
          float* base = (kq < d.k0) ? (d.g_in0 ? d.g_in0 + (row0 + rr) * d.ld_gi0 + kq : nullptr)
                                    : (d.g_in1 ? d.g_in1 + (row0 + rr) * d.ld_gi1 + (kq - d.k0) : nullptr);
          if (!base) continue;
          float4 o = f4_add(*reinterpret_cast<const float4*>(T + rr * ldt + kq), *reinterpret_cast<const float4*>(T + (rr + T5R) * ldt + kq));
          if (d.accumulate_in) o = f4_add(*reinterpret_cast<const float4*>(base), o);
          *reinterpret_cast<float4*>(base) = o;
          if (do_next) {  // the upstream unit's column sums  SUM g act'(y),  SUM g act'(y) zhat  over this thread's rows (in order)
            const float4 x_ = xv[jj];
            float zc, gy;
            zc = x_.x - n_mean.x; gy = o.x * act_grad<A_IN>(d.in_act, zc * n_scale.x + n_beta.x); ns1.x += gy; ns2.x = fmaf(gy, zc * n_rstd.x, ns2.x);
            zc = x_.y - n_mean.y; gy = o.y * act_grad<A_IN>(d.in_act, zc * n_scale.y + n_beta.y); ns1.y += gy; ns2.y = fmaf(gy, zc * n_rstd.y, ns2.y);
            zc = x_.z - n_mean.z; gy = o.z * act_grad<A_IN>(d.in_act, zc * n_scale.z + n_beta.z); ns1.z += gy; ns2.z = fmaf(gy, zc * n_rstd.z, ns2.z);
            zc = x_.w - n_mean.w; gy = o.w * act_grad<A_IN>(d.in_act, zc * n_scale.w + n_beta.w); ns1.w += gy; ns2.w = fmaf(gy, zc * n_rstd.w, ns2.w);
          }
        }
      }
      if (any_next) {  // row groups merged in order -> this tile's record of the upstream unit's red_partials
        *reinterpret_cast<float4*>(npart + rbx * K + nkq) = ns1;
        *reinterpret_cast<float4*>(npart + T5T * 4 + rbx * K + nkq) = ns2;
        __syncthreads();
        if (tid < K) {
          const int hf = tid < d.k0 ? 0 : 1, kc = hf ? tid - d.k0 : tid, kh = hf ? d.k1 : d.k0;
          float* red = hf ? d.next_red1 : d.next_red0;
          if (red) {
            float a = 0.f, b = 0.f;
            for (int m = 0; m < rsx; ++m) {
              a += npart[m * K + tid];
              b += npart[T5T * 4 + m * K + tid];
            }
            red[((int64_t)tile * 2 + 0) * kh + kc] = a;
            red[((int64_t)tile * 2 + 1) * kh + kc] = b;
          }
        }
      }
    }
    tc5::fence_before_sync();
    __syncthreads();  // staging and TMEM reads are done before the next tile overwrites them
    tc5::fence_after_sync();
    if (first) CWN_PHASE(5);
  }
  if (warp == 0) tc5::tmem_dealloc(tmem_s, L.tmem_cols);
  if (any_next && d.next_counter) {  // the last CTA of the problem finalises the upstream units (c1, c2, g_gamma, g_beta)
    if (last_cta_of_problem(d.next_counter, g.start[p + 1] - g.start[p])) {
      float* stage = reinterpret_cast<float*>(smem5);  // everything staged there is dead (>= 32 KB)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float* red = hf ? d.next_red1 : d.next_red0;
        if (!red) continue;
        cwn_unit_bwd_desc u = {};
        u.n_rows = d.n_rows; u.n_rows_live = d.n_rows_live; u.h = hf ? d.k1 : d.k0; u.red_partials = red;
        u.c1 = hf ? d.next_c1_1 : d.next_c1_0; u.c2 = hf ? d.next_c2_1 : d.next_c2_0;
        u.g_gamma = hf ? d.next_g_gamma1 : d.next_g_gamma0; u.g_beta = hf ? d.next_g_beta1 : d.next_g_beta0;
        u.accumulate_affine = hf ? d.next_accumulate_affine1 : d.next_accumulate_affine0;
        unit_bwd_finalize_body(u, stage, T5R);
        __syncthreads();
      }
      if (tid == 0) *d.next_counter = 0;
    }
  }
  CWN_PHASE(8);
}

// ------------------------------------------------------------------------------------------------ eligibility
inline bool t5_enabled() {
  static const bool on = [] { const char* v = getenv("CWN_B200_DENSE_TC5"); return !(v && v[0] == '0'); }();
  return on;
}

inline bool t5_fwd_ok(const cwn_linear_desc& d, size_t& smem) {
  const int K = d.k0 + d.k1, h = d.h;
  const bool shape = (h == 16 || h == 32 || h == 64 || h == 128) && (K == 8 || K == 16 || K == 32 || K == 64 || K == 128) &&
                     d.k0 % 4 == 0 && d.k1 % 4 == 0;
  if (!shape) return false;
  auto vec_ok = [](const float* a, const float* b, const float* c) { return aligned16(a) && aligned16(b) && aligned16(c); };
  const bool lay = aligned16(d.x0) && d.ld_x0 % 4 == 0 && (d.k1 == 0 || (aligned16(d.x1) && d.ld_x1 % 4 == 0)) && aligned16(d.w) &&
                   d.ld_w % 4 == 0 && aligned16(d.z) && d.ld_z % 4 == 0 && (!d.bias || aligned16(d.bias)) &&
                   (!d.in_scale0 || vec_ok(d.in_mean0, d.in_scale0, d.in_beta0)) &&
                   (!d.in_scale1 || vec_ok(d.in_mean1, d.in_scale1, d.in_beta1)) && (!d.stats || aligned16(d.stats));
  if (!lay) return false;
  const T5FwdSmem L(K, h);
  if (L.total > smem) smem = L.total;
  return L.total <= 200u * 1024u;
}

inline bool t5_bwd_ok(const cwn_unit_bwd_desc& d, size_t& smem) {
  const int K = d.k0 + d.k1, h = d.h;
  const bool shape = h == 64 && (K == 32 || K == 64 || K == 128) && d.k0 % 4 == 0 && d.k1 % 4 == 0;
  if (!shape) return false;
  auto vec_ok = [](const float* a, const float* b, const float* c) { return aligned16(a) && aligned16(b) && aligned16(c); };
  const bool lay = aligned16(d.x0) && d.ld_x0 % 4 == 0 && (d.k1 == 0 || (aligned16(d.x1) && d.ld_x1 % 4 == 0)) && aligned16(d.w) &&
                   d.ld_w % 4 == 0 && aligned16(d.z) && d.ld_z % 4 == 0 && aligned16(d.g_out) && d.ld_g % 4 == 0 &&
                   (!d.in_scale0 || vec_ok(d.in_mean0, d.in_scale0, d.in_beta0)) &&
                   (!d.in_scale1 || vec_ok(d.in_mean1, d.in_scale1, d.in_beta1)) &&
                   (!d.g_in0 || (aligned16(d.g_in0) && d.ld_gi0 % 4 == 0)) && (!d.g_in1 || (aligned16(d.g_in1) && d.ld_gi1 % 4 == 0)) &&
                   aligned16(d.w_partials) && (!d.has_bn || (aligned16(d.mean) && aligned16(d.scale) && aligned16(d.rstd)));
  if (!lay) return false;
  const T5BwdSmem L(K, h);
  if (L.total > smem) smem = L.total;
  return L.total <= 220u * 1024u;
}

}  // namespace cwn
