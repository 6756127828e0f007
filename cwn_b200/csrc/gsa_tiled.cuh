// Tile-staged variants of the CSR passes for the HBM-bound regime (>= 32 768 destination rows).
//
// ncu / timing of the row-per-group kernels at 1M rows showed them bound by a DEPENDENT chain per row, not by DRAM:
// rowptr -> message indices -> gathers (2-3 trips) -> residual -> store, 4-5 exposed memory latencies for ~640
// compulsory bytes. Here a persistent CTA walks tiles of TR consecutive rows and the plan of a tile (its rowptr slice
// and the contiguous slices of its one or two payload arrays) is brought into shared memory by TMA bulk copies
// (cp.async.bulk, completion on an mbarrier) TWO / ONE tiles ahead of use:
//     iteration j:   issue rowptr(j+2)   |   issue indices(j+1) (needs rowptr(j+1), which landed an iteration ago)
//                    |   compute tile j with rowptr(j), indices(j) already resident
// so the only exposed latency of a row is the feature gather itself, issued U messages (x operands) at a time.
// Bulk copies need 16-byte aligned global addresses and sizes: a slice [m0, m1) is copied as its aligned hull
// [m0 & ~3, m1 & ~3) and the <= 3 trailing entries are moved by three threads with plain loads (registers during the
// compute phase, shared memory before the end-of-tile barrier). A tile with more than kTileCap messages is not staged:
// its rows read the plan from global memory (same code path, one predicate).
//
// FEATURE WINDOWS. Batches of complexes are block-diagonal: the messages of TR consecutive destination rows come from
// a narrow, contiguous range of source rows (the cells of the same few complexes). Once the indices of the next tile
// are resident, the CTA reduces them to [lo, hi] per payload array (redux.sync + two shared atomics per warp); if
// the windows fit the feature buffers, warp 0 brings rows lo..hi of the operand matrices into shared memory with bulk
// copies (one per operand when rows are contiguous, else one per row) and the tile's gathers become ld.shared.v4.
// DRAM is then read in large asynchronous bursts issued a whole tile ahead instead of one dependent 256-byte gather at
// a time per lane group (which left ~10 KB in flight per SM, a quarter of what the HBM latency-bandwidth product
// needs). A tile whose windows do not fit (random adjacency) gathers straight from global memory as before.
//
// Accumulation order is unchanged: one lane group owns a row and adds its messages in plan order, so results are
// bit-identical to the row-per-group kernels and to a sequential CPU scatter_add_.
#pragma once
#include "common.cuh"

namespace cwn {

constexpr int kTileCap = 1024;  // messages staged per tile (per payload array)
#ifndef CWN_TILED_MIN_BLOCKS
#define CWN_TILED_MIN_BLOCKS 3  // A/B switch: resident CTAs per SM the register allocation must allow
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared (this CTA), bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Shared-memory plan of the tiles in flight + the bookkeeping of the pipeline.
template <int TR, int NARR>
struct TileStage {
  static_assert(TR % 4 == 0, "tiles must start on a 16-byte boundary of rowptr");
  struct Smem {
    alignas(16) int32_t rp[3][TR + 4];
    alignas(16) int32_t ix[2][NARR][kTileCap + 4];
    alignas(8) uint64_t rp_bar[3];
    alignas(8) uint64_t ix_bar[2];
    alignas(8) uint64_t ft_bar;
    int32_t ft_lo[NARR];    // first row of the feature window of the CURRENT tile
    int32_t ft_ok;          // the current tile's features are in shared memory
  };
  Smem& s;
  const int32_t* rowptr;
  const int32_t* arr[NARR];
  // operand matrices behind the payload arrays (feature staging): base, pitch and row width in bytes, buffer + capacity
  const char* mat[NARR];
  uint32_t pitch[NARR], rowbytes, ft_cap[NARR];
  unsigned char* ft_buf[NARR];
  int64_t n_rows, n_tiles;
  // tail entries travelling through registers during the compute phase
  int32_t rp_tail_v, ix_tail_v;
  int rp_tail_at, ix_tail_at;  // shared-memory slot (-1: nothing)
  int ix_tail_buf, ix_tail_arr, rp_tail_buf;

  __device__ __forceinline__ TileStage(Smem& smem, const int32_t* rowptr_, int64_t n_rows_)
      : s(smem), rowptr(rowptr_), n_rows(n_rows_), n_tiles((n_rows_ + TR - 1) / TR), rp_tail_at(-1), ix_tail_at(-1) {}

  __device__ __forceinline__ int64_t tile_of(int j) const { return (int64_t)blockIdx.x + (int64_t)j * gridDim.x; }
  __device__ __forceinline__ int rows_in(int64_t tile) const {
    const int64_t left = n_rows - tile * TR;
    return (int)(left < TR ? left : TR);
  }

  __device__ __forceinline__ void init() {
    if (threadIdx.x == 0) {
      for (int b = 0; b < 3; ++b) mbar_init(&s.rp_bar[b], 1);
      for (int b = 0; b < 2; ++b) mbar_init(&s.ix_bar[b], 1);
      mbar_init(&s.ft_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  // rowptr slice of tile j: bulk part now (thread 0), tail through a register (threads 32..34)
  __device__ __forceinline__ void issue_rowptr(int j) {
    const int64_t tile = tile_of(j);
    const int cnt = rows_in(tile) + 1, bulk = cnt & ~3, buf = j % 3;
    const int32_t* src = rowptr + tile * TR;
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&s.rp_bar[buf], (uint32_t)bulk * 4u);
      if (bulk) bulk_g2s(&s.rp[buf][0], src, (uint32_t)bulk * 4u, &s.rp_bar[buf]);
    }
    const int t = (int)threadIdx.x - 32;
    if (t >= 0 && t < cnt - bulk) {
      rp_tail_v = __ldg(src + bulk + t);
      rp_tail_at = bulk + t;
      rp_tail_buf = buf;
    }
  }
  __device__ __forceinline__ void wait_rowptr(int j) { mbar_wait(&s.rp_bar[j % 3], (uint32_t)(j / 3) & 1u); }

  // message range of tile j (its rowptr must be resident) and whether it fits the staging buffers
  __device__ __forceinline__ void range(int j, int& m0, int& m1, int& a0, bool& staged) const {
    const int nrow = rows_in(tile_of(j));
    m0 = s.rp[j % 3][0];
    m1 = s.rp[j % 3][nrow];
    a0 = m0 & ~3;
    staged = (m1 - a0) <= kTileCap;
  }

  __device__ __forceinline__ void issue_indices(int j) {
    int m0, m1, a0;
    bool staged;
    range(j, m0, m1, a0, staged);
    const int buf = j % 2;
    const int bulk_end = m1 & ~3;
    const uint32_t bytes = staged ? (uint32_t)(bulk_end - a0) * 4u : 0u;
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&s.ix_bar[buf], bytes * NARR);
      if (bytes)
#pragma unroll
        for (int a = 0; a < NARR; ++a) bulk_g2s(&s.ix[buf][a][0], arr[a] + a0, bytes, &s.ix_bar[buf]);
    }
    if (staged) {
      const int t = (int)threadIdx.x - 64;  // threads 64.. : 4 slots per array (3 used at most)
      const int a = t >> 2, k = t & 3;
      if (t >= 0 && a < NARR && bulk_end + k < m1) {
        const int32_t* p = (NARR > 1 && a == 1) ? arr[NARR - 1] : arr[0];  // no dynamic indexing: keeps `arr` in registers
        ix_tail_v = __ldg(p + bulk_end + k);
        ix_tail_at = bulk_end + k - a0;
        ix_tail_buf = buf;
        ix_tail_arr = a;
      }
    }
  }
  __device__ __forceinline__ void wait_indices(int j) { mbar_wait(&s.ix_bar[j % 2], (uint32_t)(j / 2) & 1u); }

  // write the tails picked up by issue_* to shared memory (call before the barrier that ends the iteration)
  __device__ __forceinline__ void commit_tails() {
    if (rp_tail_at >= 0) { s.rp[rp_tail_buf][rp_tail_at] = rp_tail_v; rp_tail_at = -1; }
    if (ix_tail_at >= 0) { s.ix[ix_tail_buf][ix_tail_arr][ix_tail_at] = ix_tail_v; ix_tail_at = -1; }
  }

  // Feature windows of tile j, by WARP 0 alone (the feature buffers must be free): wait for the tile's indices,
  // reduce them to [lo, hi] per payload array (a tile has a few hundred messages: ~10 shared loads per lane and
  // one redux.sync pair), decide, issue the copies. No CTA barrier.
  __device__ __forceinline__ void stage_features(int j) {
    if (threadIdx.x >= 32) return;
    wait_indices(j);
    int m0, m1, a0;
    bool staged;
    range(j, m0, m1, a0, staged);
    bool ok = staged && m1 > m0 && ft_cap[0] > 0;
    uint32_t total = 0;
    int lo[NARR], cnt[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) {
      int l = INT32_MAX, h = INT32_MIN;
      if (ok) {
        const int32_t* ix = &s.ix[j % 2][a][0] - a0;
        for (int m = m0 + (int)threadIdx.x; m < m1; m += 32) {
          const int v = ix[m];
          l = min(l, v);
          h = max(h, v);
        }
      }
      l = __reduce_min_sync(0xffffffffu, l);
      h = __reduce_max_sync(0xffffffffu, h);
      lo[a] = l;
      cnt[a] = ok ? h - l + 1 : 0;
      ok = ok && cnt[a] > 0 && (uint64_t)cnt[a] * rowbytes <= ft_cap[a];
      total += (uint32_t)cnt[a] * rowbytes;
    }
    if (threadIdx.x == 0) {
#pragma unroll
      for (int a = 0; a < NARR; ++a) s.ft_lo[a] = lo[a];
      s.ft_ok = ok ? 1 : 0;
      mbar_arrive_expect_tx(&s.ft_bar, ok ? total : 0u);
    }
    __syncwarp();
    if (ok) {
#pragma unroll
      for (int a = 0; a < NARR; ++a) {
        const char* src = mat[a] + (uint64_t)(uint32_t)lo[a] * pitch[a];
        if (pitch[a] == rowbytes) {  // contiguous rows: one copy
          if (threadIdx.x == 0) bulk_g2s(ft_buf[a], src, (uint32_t)cnt[a] * rowbytes, &s.ft_bar);
        } else {
          for (int rr = (int)threadIdx.x; rr < cnt[a]; rr += 32)
            bulk_g2s(ft_buf[a] + (size_t)rr * rowbytes, src + (uint64_t)rr * pitch[a], rowbytes, &s.ft_bar);
        }
      }
    }
  }
  __device__ __forceinline__ void wait_features(int j) { mbar_wait(&s.ft_bar, (uint32_t)j & 1u); }

  // Drives the pipeline; `body(r0, nrow, rp, staged, ix_base, a0, ft_ok)` computes one tile:
  //   rp[k] = rowptr[r0 + k] (shared), message m of payload a is ix_base[a*(kTileCap+4) + m - a0] when staged,
  //   arr[a][m] otherwise; when ft_ok, row q of operand a sits at ft_buf[a] + (q - s.ft_lo[a]) * rowbytes.
  // Only warps 0 and 2 ever poll an mbarrier (a polling loop executes instructions: 256 threads spinning on the
  // feature barrier cost more issue slots than the tile's arithmetic — seen in ncu); everyone else learns through
  // the CTA barrier that follows.
  template <class Body>
  __device__ __forceinline__ void run(Body body) {
    if (tile_of(0) >= n_tiles) return;
    const int warp = threadIdx.x >> 5;
    init();
    issue_rowptr(0);
    commit_tails();
    if (tile_of(1) < n_tiles) issue_rowptr(1);
    commit_tails();
    __syncthreads();
    if (warp == 0 || warp == 2) {
      wait_rowptr(0);
      issue_indices(0);
      commit_tails();
    }
    __syncthreads();
    stage_features(0);
    for (int j = 0; tile_of(j) < n_tiles; ++j) {
      if (tile_of(j + 2) < n_tiles) issue_rowptr(j + 2);
      const bool has_next = tile_of(j + 1) < n_tiles;
      if (has_next && (warp == 0 || warp == 2)) {
        wait_rowptr(j + 1);
        issue_indices(j + 1);
      }
      if (warp == 0) {
        wait_rowptr(j);
        wait_indices(j);
        wait_features(j);
      }
      __syncthreads();
      int m0, m1, a0;
      bool staged;
      range(j, m0, m1, a0, staged);
      const int64_t tile = tile_of(j);
      body(tile * TR, rows_in(tile), &s.rp[j % 3][0], staged, &s.ix[j % 2][0][0], a0, s.ft_ok != 0);
      commit_tails();
      __syncthreads();  // tile j is done with its plan slices and the feature buffers
      if (has_next) stage_features(j + 1);
    }
  }
};

// dynamic shared memory: the feature buffers of the operands, [cap0 | cap1] bytes
extern __shared__ __align__(128) unsigned char cwn_tile_features[];

template <int LPR> struct TileRows { static constexpr int GPB = kThreads / LPR, value = (4 * GPB < 64) ? 64 : ((4 * GPB > 128) ? 128 : 4 * GPB); };

// out[r] = (1+eps) * x_res[r] + REDUCE_i x_src[idx[i]]      (tile-staged; idx != NULL, F == W * LPR * VPL)
template <typename V, int LPR, int VPL, int REDUCE>
__global__ void __launch_bounds__(kThreads, (VPL == 1) ? CWN_TILED_MIN_BLOCKS : 2)
csr_gather_reduce_tiled_kernel(const float* __restrict__ x_src, int64_t ld_src, const int32_t* __restrict__ rowptr,
                               const int32_t* __restrict__ idx, int64_t n_rows, int FV,
                               const float* __restrict__ x_res, int64_t ld_res, const float* __restrict__ eps,
                               float* __restrict__ out, int64_t ld_out, uint32_t cap0) {
  using O = VecOps<V>;
  constexpr int GPB = kThreads / LPR, TR = TileRows<LPR>::value;
  constexpr int U = (VPL == 1) ? 8 : 4;
  using Stage = TileStage<TR, 1>;
  __shared__ typename Stage::Smem smem;
  Stage st(smem, rowptr, n_rows);
  const uint32_t pitch_s = (uint32_t)ld_src * 4u, pitch_r = (uint32_t)ld_res * 4u, pitch_o = (uint32_t)ld_out * 4u;
  st.arr[0] = idx;
  st.mat[0] = reinterpret_cast<const char*>(x_src);
  st.pitch[0] = pitch_s;
  st.rowbytes = (uint32_t)FV * (uint32_t)sizeof(V);
  st.ft_cap[0] = cap0;
  st.ft_buf[0] = cwn_tile_features;
  const int lane = threadIdx.x % LPR, sub = threadIdx.x / LPR;
  const float scale = x_res ? __fadd_rn(1.f, eps ? __ldg(eps) : 0.f) : 0.f;
  const bool live0 = lane < FV, live1 = VPL == 2 && lane + LPR < FV;
  const char* xs = opaque(reinterpret_cast<const char*>(x_src) + (size_t)lane * sizeof(V));
  const char* xr = opaque(reinterpret_cast<const char*>(x_res) + (size_t)lane * sizeof(V));
  char* xo = reinterpret_cast<char*>(out) + (size_t)lane * sizeof(V);
  const uint32_t fs = smem_u32(cwn_tile_features) + (uint32_t)lane * (uint32_t)sizeof(V);
  st.run([&](int64_t r0, int nrow, const int32_t* rp, bool staged, const int32_t* ix_base, int a0, bool ft_ok) {
    if (!live0) return;
    const int32_t* ix0 = ix_base - a0;
    const SharedRows<V, LPR> lds{fs - (uint32_t)smem.ft_lo[0] * st.rowbytes, st.rowbytes};
    const GlobalRows<V, LPR> ldx{xs, pitch_s};
    for (int rr = sub; rr < nrow; rr += GPB) {
      const int beg = rp[rr], end = rp[rr + 1];
      const uint32_t r = (uint32_t)(r0 + rr);
      V acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
      if (ft_ok) gather_row<V, VPL, REDUCE, U>(acc, beg, end, lds, [&](int m) { return ix0[m]; }, live1);
      else gather_row<V, VPL, REDUCE, 4>(acc, beg, end, ldx, [&](int m) { return staged ? ix0[m] : __ldg(idx + m); }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        V a = acc[k];
        if (REDUCE == CWN_REDUCE_MEAN) {
          const float cnt = (float)max(end - beg, 1);
          a = O::map2(a, a, [cnt](float x, float) { return __fdiv_rn(x, cnt); });
        }
        if (x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, r, pitch_r) + k * LPR * (int)sizeof(V)));
        O::store(row_at(xo, r, pitch_o) + k * LPR * (int)sizeof(V), a);
      }
    }
  });
}

// out[r] = (1+eps) * x_res[r] + SUM_i act(P[src[i]] + Q[cob[i]])      (tile-staged)
template <typename V, int LPR, int VPL, int ACT>
__global__ void __launch_bounds__(kThreads, (VPL == 1) ? CWN_TILED_MIN_BLOCKS : 2)
csr_cob_fwd_tiled_kernel(const float* __restrict__ P, int64_t ld_p, const float* __restrict__ Q, int64_t ld_q,
                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                         const int32_t* __restrict__ cob, int64_t n_rows, int FV, const float* __restrict__ x_res,
                         int64_t ld_res, const float* __restrict__ eps, float* __restrict__ out, int64_t ld_out,
                         uint32_t cap0, uint32_t cap1) {
  using O = VecOps<V>;
  constexpr int GPB = kThreads / LPR, TR = TileRows<LPR>::value;
  constexpr int U = (VPL == 1) ? 4 : 2;
  using Stage = TileStage<TR, 2>;
  __shared__ typename Stage::Smem smem;
  Stage st(smem, rowptr, n_rows);
  const uint32_t pitch_p = (uint32_t)ld_p * 4u, pitch_q = (uint32_t)ld_q * 4u, pitch_r = (uint32_t)ld_res * 4u,
                 pitch_o = (uint32_t)ld_out * 4u;
  st.arr[0] = src;
  st.arr[1] = cob;
  st.mat[0] = reinterpret_cast<const char*>(P);
  st.mat[1] = reinterpret_cast<const char*>(Q);
  st.pitch[0] = pitch_p;
  st.pitch[1] = pitch_q;
  st.rowbytes = (uint32_t)FV * (uint32_t)sizeof(V);
  st.ft_cap[0] = cap0;
  st.ft_cap[1] = cap1;
  st.ft_buf[0] = cwn_tile_features;
  st.ft_buf[1] = cwn_tile_features + cap0;
  const int lane = threadIdx.x % LPR, sub = threadIdx.x / LPR;
  const float scale = x_res ? __fadd_rn(1.f, eps ? __ldg(eps) : 0.f) : 0.f;
  const bool live0 = lane < FV, live1 = VPL == 2 && lane + LPR < FV;
  const char* pl = opaque(reinterpret_cast<const char*>(P) + (size_t)lane * sizeof(V));
  const char* ql = opaque(reinterpret_cast<const char*>(Q) + (size_t)lane * sizeof(V));
  const char* xr = opaque(reinterpret_cast<const char*>(x_res) + (size_t)lane * sizeof(V));
  char* xo = reinterpret_cast<char*>(out) + (size_t)lane * sizeof(V);
  const uint32_t fs0 = smem_u32(cwn_tile_features) + (uint32_t)lane * (uint32_t)sizeof(V), fs1 = fs0 + cap0;
  st.run([&](int64_t r0, int nrow, const int32_t* rp, bool staged, const int32_t* ix_base, int a0, bool ft_ok) {
    if (!live0) return;
    const int32_t *ix0 = ix_base - a0, *ix1 = ix0 + (kTileCap + 4);
    const SharedRows<V, LPR> lsp{fs0 - (uint32_t)smem.ft_lo[0] * st.rowbytes, st.rowbytes};
    const SharedRows<V, LPR> lsq{fs1 - (uint32_t)smem.ft_lo[1] * st.rowbytes, st.rowbytes};
    const GlobalRows<V, LPR> lgp{pl, pitch_p}, lgq{ql, pitch_q};
    for (int rr = sub; rr < nrow; rr += GPB) {
      const int beg = rp[rr], end = rp[rr + 1];
      const uint32_t r = (uint32_t)(r0 + rr);
      V acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
      if (ft_ok)
        cob_fwd_row<V, VPL, ACT, U>(acc, beg, end, lsp, lsq, [&](int m) { return ix0[m]; }, [&](int m) { return ix1[m]; }, live1);
      else
        cob_fwd_row<V, VPL, ACT, 2>(acc, beg, end, lgp, lgq, [&](int m) { return staged ? ix0[m] : __ldg(src + m); },
                                    [&](int m) { return staged ? ix1[m] : __ldg(cob + m); }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        V a = acc[k];
        if (x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, r, pitch_r) + k * LPR * (int)sizeof(V)));
        O::store(row_at(xo, r, pitch_o) + k * LPR * (int)sizeof(V), a);
      }
    }
  });
}

// gA[r] = SUM_i G[dst[i]] * act'(A[r] + B[oth[i]])      (tile-staged)
template <typename V, int LPR, int VPL, int ACT>
__global__ void __launch_bounds__(kThreads, (VPL == 1) ? CWN_TILED_MIN_BLOCKS : 2)
csr_cob_bwd_tiled_kernel(const float* __restrict__ G, int64_t ld_g, const float* __restrict__ A, int64_t ld_a,
                         const float* __restrict__ B, int64_t ld_b, const int32_t* __restrict__ rowptr,
                         const int32_t* __restrict__ dst, const int32_t* __restrict__ oth, int64_t n_rows, int FV,
                         float* __restrict__ gA, int64_t ld_ga, uint32_t cap0, uint32_t cap1) {
  using O = VecOps<V>;
  constexpr int GPB = kThreads / LPR, TR = TileRows<LPR>::value;
  constexpr int U = (VPL == 1) ? 4 : 2;
  using Stage = TileStage<TR, 2>;
  __shared__ typename Stage::Smem smem;
  Stage st(smem, rowptr, n_rows);
  const uint32_t pitch_g = (uint32_t)ld_g * 4u, pitch_a = (uint32_t)ld_a * 4u, pitch_b = (uint32_t)ld_b * 4u,
                 pitch_o = (uint32_t)ld_ga * 4u;
  st.arr[0] = dst;
  st.arr[1] = oth;
  st.mat[0] = reinterpret_cast<const char*>(G);
  st.mat[1] = reinterpret_cast<const char*>(B);
  st.pitch[0] = pitch_g;
  st.pitch[1] = pitch_b;
  st.rowbytes = (uint32_t)FV * (uint32_t)sizeof(V);
  st.ft_cap[0] = cap0;
  st.ft_cap[1] = cap1;
  st.ft_buf[0] = cwn_tile_features;
  st.ft_buf[1] = cwn_tile_features + cap0;
  const int lane = threadIdx.x % LPR, sub = threadIdx.x / LPR;
  const bool live0 = lane < FV, live1 = VPL == 2 && lane + LPR < FV;
  const char* gl = opaque(reinterpret_cast<const char*>(G) + (size_t)lane * sizeof(V));
  const char* al = opaque(reinterpret_cast<const char*>(A) + (size_t)lane * sizeof(V));
  const char* bl = opaque(reinterpret_cast<const char*>(B) + (size_t)lane * sizeof(V));
  char* xo = reinterpret_cast<char*>(gA) + (size_t)lane * sizeof(V);
  const uint32_t fs0 = smem_u32(cwn_tile_features) + (uint32_t)lane * (uint32_t)sizeof(V), fs1 = fs0 + cap0;
  st.run([&](int64_t r0, int nrow, const int32_t* rp, bool staged, const int32_t* ix_base, int a0, bool ft_ok) {
    if (!live0) return;
    const int32_t *ix0 = ix_base - a0, *ix1 = ix0 + (kTileCap + 4);
    const SharedRows<V, LPR> lsg{fs0 - (uint32_t)smem.ft_lo[0] * st.rowbytes, st.rowbytes};
    const SharedRows<V, LPR> lsb{fs1 - (uint32_t)smem.ft_lo[1] * st.rowbytes, st.rowbytes};
    const GlobalRows<V, LPR> lgg{gl, pitch_g}, lgb{bl, pitch_b};
    for (int rr = sub; rr < nrow; rr += GPB) {
      const int beg = rp[rr], end = rp[rr + 1];
      const uint32_t r = (uint32_t)(r0 + rr);
      V acc[VPL], a[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        acc[k] = O::zero();
        a[k] = (beg < end && (k == 0 || live1)) ? O::load(row_at(al, r, pitch_a) + k * LPR * (int)sizeof(V)) : O::zero();
      }
      if (ft_ok)
        cob_bwd_row<V, VPL, ACT, U>(acc, a, beg, end, lsg, lsb, [&](int m) { return ix0[m]; }, [&](int m) { return ix1[m]; }, live1);
      else
        cob_bwd_row<V, VPL, ACT, 2>(acc, a, beg, end, lgg, lgb, [&](int m) { return staged ? ix0[m] : __ldg(dst + m); },
                                    [&](int m) { return staged ? ix1[m] : __ldg(oth + m); }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        O::store(row_at(xo, r, pitch_o) + k * LPR * (int)sizeof(V), acc[k]);
      }
    }
  });
}

}  // namespace cwn
