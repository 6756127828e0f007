// Warp-specialised, TMA-fed CSR passes for the HBM-bound regime (>= 32 768 destination rows per launch).
//
// Why: ncu on the register-level kernels of gsa.cu at 1M rows (profiles/README.md) shows no wasted DRAM traffic but only
// 30-40 % of the DRAM throughput: every warp both waits for its gathers (long-scoreboard, 10 warps per issue) and does
// the arithmetic, and the bytes a lane group keeps in flight are bounded by its registers. The first tile-staged kernels
// (gsa_tiled.cuh) moved the gathers to shared memory but kept every warp on the mbarriers / CTA barriers and scanned the
// window of a tile inside the kernel. Here the two roles are split and the scan is gone:
//
//   * the PLAN carries, per tile of TR consecutive rows, its message range and the contiguous range of operand rows its
//     messages touch (`cwn_csr_tile_windows`, once per plan: batches of complexes are block-diagonal, so the window of a
//     tile is a little more than TR rows);
//   * ONE producer warp per CTA (persistent, one CTA per SM) walks the CTA's tiles S stages ahead. For every tile it
//     issues TMA bulk copies (cp.async.bulk global -> shared, completion on the stage's `full` mbarrier) of the rowptr
//     slice, the payload slices, the operand windows and the tile's own rows of the residual / row operand. Tile metadata
//     for 32 tiles at a time sits in the producer's registers (one coalesced load per 32 tiles, double-buffered), so
//     nothing on the producer's path waits on DRAM except the `empty` barrier of a stage;
//   * EIGHT consumer warps wait on `full`, run the rows of the tile entirely out of shared memory (ld.shared for the
//     plan and the features), store the result rows straight to global memory and arrive on `empty` (one lane per warp).
//
// The DRAM stream is therefore S-1 stages (~100-150 KB per SM) of asynchronous, fully coalesced bursts instead of
// dependent 256-byte gathers. Accumulation order is unchanged — one lane group owns a row and adds its messages in plan
// order with IEEE round-to-nearest adds — so results are bit-identical to the row kernels and to a sequential CPU
// scatter_add_. A tile that cannot be staged (window larger than the buffer, the ragged last tile, unaligned tail of the
// payload arrays) runs the same row loop through generic pointers to global memory.
//
// Replaces (reference): Tensor.index_select + torch_scatter.scatter, mp/cell_mp.py:195-198 + :423-479, and the
// per-message Linear of mp/layers.py:290-293 in split-weight form, for batches large enough to be HBM-bound.
#include "common.cuh"

namespace cwn {
namespace ws {

#ifndef CWN_WS_CONSUMER_WARPS
// A/B switch (build_variant). Measured, edge-upper F = 64 at 1M rows, one vector per lane: 8 consumer warps 108 us,
// 16 -> 84 us, 20 -> 85 us, 24 -> 98 us. Registers are allocated per SM sub-partition (4 x 16 384): with the producer that
// makes 17 warps, 5 on one sub-partition, i.e. at most 96 registers per thread (what ptxas derives from
// __launch_bounds__(544, 1); 112 or 120 via __maxnreg__ fail to launch) — enough for one vector per lane.
#define CWN_WS_CONSUMER_WARPS 16
#endif
constexpr int kConsumerWarps = CWN_WS_CONSUMER_WARPS;
constexpr int kBlock = 32 * (kConsumerWarps + 1);
constexpr int kMaxStages = 8;
#ifndef CWN_WS_U2
#define CWN_WS_U2 2
#endif
constexpr uint32_t kSmemBudget = 220 * 1024;   // dynamic shared memory of a CTA (227 KB less barriers / control / slack)

struct Params {
  const int32_t* rowptr;
  const int32_t* arr[2];   // payload columns of the plan
  const int32_t* windows;  // [n_tiles][8] from cwn_csr_tile_windows
  const char* mat[2];      // operand matrices behind the payload columns
  uint32_t pitch[2];       // bytes
  const char* rowop;       // x_res / A: rows aligned with the destination rows (nullable)
  uint32_t pitch_ro;
  char* out;
  uint32_t pitch_out;
  const float* eps;
  int64_t n_rows, E_al;    // E_al = E & ~3: payload entries a 16-byte bulk copy may touch
  int32_t n_tiles, FV, tile_rows, stages;
  uint32_t rowbytes, cap_ix, cap_ft[2];
  uint32_t off_ix[2], off_ft[2], off_ro, stage_bytes;
};

struct StageCtl {
  int32_t m0, m1, a0, lo0, lo1, rows, flags, pad;
};
constexpr int kRp = 1, kIx = 2, kFt0 = 4, kFt1 = 8, kRo = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// TMA 1-D bulk copy global -> shared (this CTA); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int32_t lds32(uint32_t a) {
  int32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// ---- arithmetic: the same roundings as VecOps<float4> of gsa.cu (IEEE rn adds, packed two per instruction; products
// rounded before they are added)
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 add_scaled4(float4 a, float s, float4 b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(__fmul_rn(s, b.x), __fmul_rn(s, b.y)));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(__fmul_rn(s, b.z), __fmul_rn(s, b.w)));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <class Fn>
__device__ __forceinline__ float4 map1(float4 a, Fn f) { return make_float4(f(a.x), f(a.y), f(a.z), f(a.w)); }
template <class Fn>
__device__ __forceinline__ float4 map2(float4 a, float4 b, Fn f) {
  return make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
}

// ---- passes: what a message contributes and how a row is finished
template <int REDUCE>
struct Gather {  // out[r] = (1+eps) x_res[r] + REDUCE_i x_src[idx[i]]
  static constexpr int NARR = 1, kU = 4;  // messages in flight per lane group (shared-memory latency only)
  static constexpr bool kRowFirst = false;
  static __device__ __forceinline__ float4 fold(float4 acc, float4 v0, float4, float4) { return add4(acc, v0); }
  static __device__ __forceinline__ float4 finish(float4 acc, int cnt, bool has_ro, float scale, float4 ro) {
    if (REDUCE == CWN_REDUCE_MEAN) {
      const float c = (float)max(cnt, 1);
      acc = map1(acc, [c](float x) { return __fdiv_rn(x, c); });
    }
    return has_ro ? add_scaled4(acc, scale, ro) : acc;
  }
};
template <int ACT>
struct CobFwd {  // out[r] = (1+eps) x_res[r] + SUM_i act(P[src[i]] + Q[cob[i]])
  static constexpr int NARR = 2, kU = CWN_WS_U2;  // (registers: 16 consumer warps leave 120 per thread)
  static constexpr bool kRowFirst = false;
  static __device__ __forceinline__ float4 fold(float4 acc, float4 v0, float4 v1, float4) {
    float4 pre = add4(v0, v1);
    pre = map1(pre, [](float x) { return act_fwd<ACT>(x); });
    return add4(acc, pre);
  }
  static __device__ __forceinline__ float4 finish(float4 acc, int, bool has_ro, float scale, float4 ro) {
    return has_ro ? add_scaled4(acc, scale, ro) : acc;
  }
};
template <int ACT>
struct CobBwd {  // gA[r] = SUM_i G[dst[i]] * act'(A[r] + B[oth[i]])
  static constexpr int NARR = 2, kU = CWN_WS_U2;
  static constexpr bool kRowFirst = true;
  static __device__ __forceinline__ float4 fold(float4 acc, float4 g, float4 b, float4 a) {
    const float4 pre = add4(a, b);
    if (ACT == CWN_ACT_RELU)  // g * {0,1} as a select (what torch's threshold_backward does)
      return add4(acc, map2(g, pre, [](float g_, float z) { return z > 0.f ? g_ : 0.f; }));
    const float4 d = map1(pre, [](float z) { return act_bwd<ACT>(z); });
    return add4(acc, map2(g, d, [](float g_, float d_) { return __fmul_rn(g_, d_); }));
  }
  static __device__ __forceinline__ float4 finish(float4 acc, int, bool, float, float4) { return acc; }
};

// ---- how a consumer reads a staged tile: shared memory (everything staged) or generic pointers (any mix)
template <int NARR>
struct SmemView {
  uint32_t rp, ix[NARR], ft[NARR], ro;  // shared-space byte addresses, lane offset and window origin folded in
  uint32_t rowbytes;
  __device__ __forceinline__ int row_begin(int i) const { return lds32(rp + 4u * (uint32_t)i); }
  __device__ __forceinline__ int index(int a, int m) const { return lds32(ix[a] + 4u * (uint32_t)m); }
  __device__ __forceinline__ float4 feature(int a, int row, uint32_t koff) const {
    return lds128(ft[a] + (uint32_t)row * rowbytes + koff);
  }
  __device__ __forceinline__ float4 rowop(int i, uint32_t koff) const { return lds128(ro + (uint32_t)i * rowbytes + koff); }
};
template <int NARR>
struct GenericView {
  const int32_t* rp;        // indexed by the row's position in the tile
  const int32_t* ix[NARR];  // indexed by the absolute message position
  const char* ft[NARR];     // lane base, indexed by the absolute operand row
  uint32_t ft_pitch[NARR];
  const char* ro;           // lane base, indexed by the row's position in the tile
  uint32_t ro_pitch;
  __device__ __forceinline__ int row_begin(int i) const { return rp[i]; }
  __device__ __forceinline__ int index(int a, int m) const { return ix[a][m]; }
  __device__ __forceinline__ float4 feature(int a, int row, uint32_t koff) const {
    return *reinterpret_cast<const float4*>(ft[a] + (int64_t)row * ft_pitch[a] + koff);
  }
  __device__ __forceinline__ float4 rowop(int i, uint32_t koff) const {
    return *reinterpret_cast<const float4*>(ro + (int64_t)i * ro_pitch + koff);
  }
};

// A lane owns kVPL vectors of a row, LPR * 16 bytes apart. Two vectors per lane (F = 64: 8 lanes per row, four rows per
// warp instruction) share the plan reads, address arithmetic, predicates and per-row bookkeeping between twice the bytes,
// but need 128 registers, i.e. 15 consumer warps instead of 16: measured equal (86.4 vs 85.5 us on the edge-upper pass,
// 138 vs 141 us cob bwd, 124 vs 118 us cob fwd; profiles/README.md), so the default stays one vector per lane and 16
// warps. `live1`: the lane's second vector is inside the row.
#ifndef CWN_WS_VPL
#define CWN_WS_VPL 1  // A/B switch (build_variant): vectors per lane
#endif
constexpr int kVPL = CWN_WS_VPL;

template <class Pass, int LPR, class View>
__device__ __forceinline__ void rows_of_tile(const View& v, int first, int rows, int m1, bool has_ro, float scale,
                                             char* out_lane, int64_t r0, uint32_t pitch_out, bool live1) {
  constexpr int G = kConsumerWarps * 32 / LPR;
  constexpr int NARR = Pass::NARR, kU = Pass::kU;
  constexpr uint32_t kStep = LPR * 16;
  for (int i = first; i < rows; i += G) {
    const int beg = v.row_begin(i);
    const int end = (i == rows - 1) ? m1 : v.row_begin(i + 1);
    float4 a_row[kVPL], acc[kVPL];
#pragma unroll
    for (int k = 0; k < kVPL; ++k) {
      acc[k] = zero4();
      a_row[k] = (Pass::kRowFirst && (k == 0 || live1)) ? v.rowop(i, k * kStep) : zero4();
    }
    for (int m = beg; m < end; m += kU) {
      int j0[kU], j1[kU];
      float4 v0[kU][kVPL], v1[kU][kVPL];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (m + u < end) {
          j0[u] = v.index(0, m + u);
          if (NARR == 2) j1[u] = v.index(NARR - 1, m + u);
        }
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (m + u < end) {
#pragma unroll
          for (int k = 0; k < kVPL; ++k) {
            v0[u][k] = v1[u][k] = zero4();
            if (k == 0 || live1) {
              v0[u][k] = v.feature(0, j0[u], k * kStep);
              if (NARR == 2) v1[u][k] = v.feature(NARR - 1, j1[u], k * kStep);
            }
          }
        }
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (m + u < end) {
#pragma unroll
          for (int k = 0; k < kVPL; ++k) acc[k] = Pass::fold(acc[k], v0[u][k], v1[u][k], a_row[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kVPL; ++k) {
      if (k == 1 && !live1) break;
      float4 ro = zero4();
      if (!Pass::kRowFirst && has_ro) ro = v.rowop(i, k * kStep);
      const float4 res = Pass::finish(acc[k], end - beg, has_ro, scale, ro);
      *reinterpret_cast<float4*>(out_lane + (r0 + i) * (int64_t)pitch_out + k * kStep) = res;
    }
  }
}

template <class Pass, int LPR>
__global__ void __launch_bounds__(kBlock, 1) csr_ws_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ StageCtl ctl[kMaxStages];
  constexpr int NARR = Pass::NARR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages, TR = p.tile_rows;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- producer
    const int4* win = reinterpret_cast<const int4*>(p.windows);
    auto load_meta = [&](int kb, int4& a, int4& b) {  // metadata of this CTA's tiles kb .. kb+31, one per lane
      const int64_t t = (int64_t)blockIdx.x + (int64_t)(kb + lane) * gridDim.x;
      a = b = make_int4(0, 0, 0, 0);
      if (t < p.n_tiles) {
        a = __ldg(win + 2 * t);
        b = __ldg(win + 2 * t + 1);
      }
    };
    int4 ca, cb, na, nb;
    load_meta(0, ca, cb);
    load_meta(32, na, nb);
    int s = 0;
    uint32_t ph = 0;  // parity of the `empty` phase that must have completed before a stage is filled again
    for (int k = 0;; ++k) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      if (tile >= p.n_tiles) break;
      if (k && (k & 31) == 0) {
        ca = na;
        cb = nb;
        load_meta(k + 32, na, nb);
      }
      const int sl = k & 31;
      const int m0 = __shfl_sync(0xffffffffu, ca.x, sl), m1 = __shfl_sync(0xffffffffu, ca.y, sl);
      int lo[2], cnt[2];
      lo[0] = __shfl_sync(0xffffffffu, ca.z, sl);
      cnt[0] = __shfl_sync(0xffffffffu, ca.w, sl);
      lo[1] = __shfl_sync(0xffffffffu, cb.x, sl);
      cnt[1] = __shfl_sync(0xffffffffu, cb.y, sl);
      if (k >= S) mbar_wait(&empty_bar[s], ph);

      const int64_t r0 = tile * TR;
      const int rows = (int)((p.n_rows - r0 < TR) ? p.n_rows - r0 : TR);
      const int a0 = m0 & ~3, hull = (m1 > m0) ? ((m1 + 3) & ~3) - a0 : 0;
      unsigned char* st = dyn + (size_t)s * p.stage_bytes;
      int flags = 0;
      uint32_t bytes = 0;
      if (rows == TR) {
        flags |= kRp;
        bytes += (uint32_t)TR * 4u;
      }
      if (hull <= (int)p.cap_ix && (int64_t)a0 + hull <= p.E_al) {
        flags |= kIx;
        bytes += (uint32_t)hull * 4u * NARR;
      }
#pragma unroll
      for (int a = 0; a < NARR; ++a)
        if ((uint64_t)cnt[a] * p.rowbytes <= p.cap_ft[a]) {
          flags |= (kFt0 << a);
          bytes += (uint32_t)cnt[a] * p.rowbytes;
        }
      if (p.rowop) {
        flags |= kRo;
        bytes += (uint32_t)rows * p.rowbytes;
      }
      if (lane == 0) {
        StageCtl c;
        c.m0 = m0; c.m1 = m1; c.a0 = a0; c.lo0 = lo[0]; c.lo1 = lo[1]; c.rows = rows; c.flags = flags; c.pad = 0;
        ctl[s] = c;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
      }
      __syncwarp();
      uint64_t* bar = &full_bar[s];
      if ((flags & kRp) && lane == 0) bulk_g2s(st, p.rowptr + r0, (uint32_t)TR * 4u, bar);
      if ((flags & kIx) && hull > 0 && lane < NARR)
        bulk_g2s(st + p.off_ix[lane], p.arr[lane] + a0, (uint32_t)hull * 4u, bar);
#pragma unroll
      for (int a = 0; a < NARR; ++a) {
        if (!(flags & (kFt0 << a)) || cnt[a] == 0) continue;
        const char* src = p.mat[a] + (size_t)lo[a] * p.pitch[a];
        if (p.pitch[a] == p.rowbytes) {
          if (lane == 2 + a) bulk_g2s(st + p.off_ft[a], src, (uint32_t)cnt[a] * p.rowbytes, bar);
        } else {
          for (int i = lane; i < cnt[a]; i += 32)
            bulk_g2s(st + p.off_ft[a] + (size_t)i * p.rowbytes, src + (size_t)i * p.pitch[a], p.rowbytes, bar);
        }
      }
      if (flags & kRo) {
        const char* src = p.rowop + (size_t)r0 * p.pitch_ro;
        if (p.pitch_ro == p.rowbytes) {
          if (lane == 4) bulk_g2s(st + p.off_ro, src, (uint32_t)rows * p.rowbytes, bar);
        } else {
          for (int i = lane; i < rows; i += 32)
            bulk_g2s(st + p.off_ro + (size_t)i * p.rowbytes, src + (size_t)i * p.pitch_ro, p.rowbytes, bar);
        }
      }
      if (++s == S) {
        s = 0;
        if (k >= S) ph ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------- consumers
    const int ct = (warp - 1) * 32 + lane;
    const int grp = ct / LPR, gl = ct % LPR;
    const bool live = gl < p.FV;
    const bool live1 = gl + LPR < p.FV;
    const bool has_ro = p.rowop != nullptr;
    const float scale = (has_ro && !Pass::kRowFirst) ? __fadd_rn(1.f, p.eps ? __ldg(p.eps) : 0.f) : 0.f;
    char* out_lane = p.out + (size_t)gl * 16;
    const uint32_t dyn_u32 = smem_u32(dyn);
    int s = 0;
    uint32_t ph = 0;
    for (int k = 0;; ++k) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      if (tile >= p.n_tiles) break;
      mbar_wait(&full_bar[s], ph);
      const StageCtl c = ctl[s];
      const int64_t r0 = tile * TR;
      constexpr int need = kRp | kIx | kFt0 | (NARR == 2 ? kFt1 : 0);
      const bool fast = (c.flags & need) == need && (!has_ro || (c.flags & kRo));
      // rows are dealt round-robin ACROSS tiles (row j of the CTA's row sequence goes to group j % G), so that when G
      // does not divide the tile height the same groups do not get the extra row every time (warps only meet at the
      // `empty` barriers, S stages apart: imbalance inside one tile is harmless, a persistent one is not)
      constexpr int G = kConsumerWarps * 32 / LPR;
      const int first = (grp + G - (int)(((int64_t)k * TR) % G)) % G;
      if (live) {
        if (fast) {
          SmemView<NARR> v;
          const uint32_t st = dyn_u32 + (uint32_t)s * p.stage_bytes;
          v.rowbytes = p.rowbytes;
          v.rp = st;
          v.ro = st + p.off_ro + (uint32_t)gl * 16u;
#pragma unroll
          for (int a = 0; a < NARR; ++a) {
            v.ix[a] = st + p.off_ix[a] - 4u * (uint32_t)c.a0;
            v.ft[a] = st + p.off_ft[a] + (uint32_t)gl * 16u - (uint32_t)(a == 0 ? c.lo0 : c.lo1) * p.rowbytes;
          }
          rows_of_tile<Pass, LPR>(v, first, c.rows, c.m1, has_ro, scale, out_lane, r0, p.pitch_out, live1);
        } else {
          GenericView<NARR> v;
          const unsigned char* st = dyn + (size_t)s * p.stage_bytes;
          v.rp = (c.flags & kRp) ? reinterpret_cast<const int32_t*>(st) : p.rowptr + r0;
#pragma unroll
          for (int a = 0; a < NARR; ++a) {
            v.ix[a] = (c.flags & kIx) ? reinterpret_cast<const int32_t*>(st + p.off_ix[a]) - c.a0 : p.arr[a];
            const int lo_a = (a == 0) ? c.lo0 : c.lo1;
            const bool staged = (c.flags & (kFt0 << a)) != 0;
            v.ft_pitch[a] = staged ? p.rowbytes : p.pitch[a];
            v.ft[a] = (staged ? reinterpret_cast<const char*>(st + p.off_ft[a]) - (int64_t)lo_a * p.rowbytes : p.mat[a]) +
                      (size_t)gl * 16;
          }
          if (c.flags & kRo) {
            v.ro = reinterpret_cast<const char*>(st + p.off_ro) + (size_t)gl * 16;
            v.ro_pitch = p.rowbytes;
          } else {
            v.ro = p.rowop ? p.rowop + (size_t)r0 * p.pitch_ro + (size_t)gl * 16 : nullptr;
            v.ro_pitch = p.pitch_ro;
          }
          rows_of_tile<Pass, LPR>(v, first, c.rows, c.m1, has_ro, scale, out_lane, r0, p.pitch_out, live1);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == S) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
}

// one warp per tile: message range and operand-row windows; running maxima for the host's buffer sizing
__global__ void tile_windows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ pay0,
                                    const int32_t* __restrict__ pay1, int64_t n_rows, int tile_rows, int64_t n_tiles,
                                    int32_t* __restrict__ windows) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= n_tiles) return;
  const int64_t r0 = tile * tile_rows;
  const int64_t r1 = (r0 + tile_rows < n_rows) ? r0 + tile_rows : n_rows;
  const int m0 = __ldg(rowptr + r0), m1 = __ldg(rowptr + r1);
  int l0 = INT32_MAX, h0 = INT32_MIN, l1 = INT32_MAX, h1 = INT32_MIN;
  for (int m = m0 + lane; m < m1; m += 32) {
    if (pay0) {
      const int v = __ldg(pay0 + m);
      l0 = min(l0, v);
      h0 = max(h0, v);
    }
    if (pay1) {
      const int v = __ldg(pay1 + m);
      l1 = min(l1, v);
      h1 = max(h1, v);
    }
  }
  l0 = __reduce_min_sync(0xffffffffu, l0);
  h0 = __reduce_max_sync(0xffffffffu, h0);
  l1 = __reduce_min_sync(0xffffffffu, l1);
  h1 = __reduce_max_sync(0xffffffffu, h1);
  if (lane == 0) {
    const int c0 = (pay0 && m1 > m0) ? h0 - l0 + 1 : 0, c1 = (pay1 && m1 > m0) ? h1 - l1 + 1 : 0;
    int4* w = reinterpret_cast<int4*>(windows) + 2 * tile;
    w[0] = make_int4(m0, m1, c0 ? l0 : 0, c0);
    w[1] = make_int4(c1 ? l1 : 0, c1, 0, 0);
    int32_t* stats = windows + 8 * n_tiles;
    atomicMax(stats + 0, c0);
    atomicMax(stats + 1, c1);
    atomicMax(stats + 2, (m1 > m0) ? ((m1 + 3) & ~3) - (m0 & ~3) : 0);
  }
}

static uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// shared-memory layout of a stage; returns the number of stages that fit (0: the configuration cannot run)
static int configure(Params& p, int F, int tile_rows, int cap_rows0, int cap_rows1, int cap_msgs, int narr, bool rowop) {
  p.rowbytes = (uint32_t)F * 4u;
  p.tile_rows = tile_rows;
  p.cap_ix = up((uint32_t)cap_msgs, 4);
  uint32_t off = up((uint32_t)tile_rows * 4u, 16);
  for (int a = 0; a < 2; ++a) {
    p.off_ix[a] = off;
    if (a < narr) off += p.cap_ix * 4u;
  }
  off = up(off, 128);
  const int cap_rows[2] = {cap_rows0, cap_rows1};
  for (int a = 0; a < 2; ++a) {
    p.off_ft[a] = off;
    p.cap_ft[a] = (a < narr) ? (uint32_t)cap_rows[a] * p.rowbytes : 0u;
    off += up(p.cap_ft[a], 128);
  }
  p.off_ro = off;
  if (rowop) off += up((uint32_t)tile_rows * p.rowbytes, 128);
  p.stage_bytes = up(off, 128);
  int stages = (int)(kSmemBudget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  return stages;
}

static int lanes_per_row(int FV) {  // two vectors per lane (kVPL)
  int lpr = 2;
  while (lpr * kVPL < FV) lpr <<= 1;
  return lpr;  // (FV <= 32: at most 32 with one vector per lane, 16 with two)
}

template <class Pass, int LPR>
static int launch(const Params& p, cudaStream_t st) {
  auto kern = csr_ws_kernel<Pass, LPR>;
  const size_t dynamic = (size_t)p.stages * p.stage_bytes;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    if (e != cudaSuccess) return cuda_status(e, "cwn ws: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  kern<<<grid, kBlock, dynamic, st>>>(p);
  return CWN_OK;
}

#define CWN_WS_BY_LPR(PASS)                                   \
  switch (lanes_per_row(p.FV)) {                              \
    case 2: rc = launch<PASS, 2>(p, st); break;               \
    case 4: rc = launch<PASS, 4>(p, st); break;               \
    case 8: rc = launch<PASS, 8>(p, st); break;               \
    case 16: rc = launch<PASS, 16>(p, st); break;             \
    default: rc = launch<PASS, 32>(p, st); break;             \
  }

#define CWN_WS_BY_ACT(PASS)                                                 \
  switch (act) {                                                            \
    case CWN_ACT_ID: CWN_WS_BY_LPR(PASS<CWN_ACT_ID>) break;                 \
    case CWN_ACT_RELU: CWN_WS_BY_LPR(PASS<CWN_ACT_RELU>) break;             \
    case CWN_ACT_ELU: CWN_WS_BY_LPR(PASS<CWN_ACT_ELU>) break;               \
    case CWN_ACT_SIGMOID: CWN_WS_BY_LPR(PASS<CWN_ACT_SIGMOID>) break;       \
    default: CWN_WS_BY_LPR(PASS<CWN_ACT_TANH>) break;                       \
  }

static int check_common(const char* who, int64_t n_rows, int32_t F, int32_t tile_rows, const int32_t* rowptr,
                        const int32_t* windows, int64_t E) {
  if (n_rows <= 0 || n_rows > INT32_MAX || E < 0 || E > INT32_MAX) return fail(CWN_E_SHAPE, who);
  if (F <= 0 || F % 4 != 0 || F > 128) return fail(CWN_E_SHAPE, "cwn ws: F must be a multiple of 4, <= 128");
  if (tile_rows < 4 || tile_rows % 4 != 0) return fail(CWN_E_SHAPE, "cwn ws: tile_rows must be a multiple of 4");
  if (!rowptr || !windows) return fail(CWN_E_NULL, who);
  if (!aligned16(rowptr) || !aligned16(windows)) return fail(CWN_E_ALIGN, "cwn ws: rowptr / windows must be 16-byte aligned");
  return CWN_OK;
}
static int check_mat(const float* m, int64_t ld, int F, const char* name) {
  if (!m) return fail(CWN_E_NULL, name);
  if (ld < F || ld * 4 > (int64_t)UINT32_MAX) return fail(CWN_E_SHAPE, name);
  if (!aligned16(m) || ld % 4 != 0) return fail(CWN_E_ALIGN, name);
  return CWN_OK;
}

}  // namespace ws
}  // namespace cwn

using namespace cwn;
using namespace cwn::ws;

extern "C" int cwn_csr_tile_windows(const int32_t* rowptr, const int32_t* pay0, const int32_t* pay1, int64_t n_rows,
                                    int32_t tile_rows, int32_t* windows, cwn_stream_t stream) {
  if (n_rows <= 0 || n_rows > INT32_MAX || tile_rows <= 0) return fail(CWN_E_SHAPE, "cwn_csr_tile_windows: bad n_rows / tile_rows");
  if (!rowptr || !windows) return fail(CWN_E_NULL, "cwn_csr_tile_windows");
  if (!aligned16(windows)) return fail(CWN_E_ALIGN, "cwn_csr_tile_windows: windows must be 16-byte aligned");
  const int64_t n_tiles = (n_rows + tile_rows - 1) / tile_rows;
  const int threads = 256;
  const int64_t blocks = (n_tiles * 32 + threads - 1) / threads;
  tile_windows_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(rowptr, pay0, pay1, n_rows, tile_rows,
                                                                               n_tiles, windows);
  return launched("cwn_csr_tile_windows");
}

extern "C" int cwn_csr_ws_consumer_threads(void) { return 32 * kConsumerWarps; }
extern "C" int cwn_csr_ws_lanes_per_row(int32_t F) { return (F > 0 && F % 4 == 0) ? lanes_per_row(F / 4) : 0; }

extern "C" int cwn_csr_ws_stages(int32_t F, int32_t tile_rows, int32_t cap_rows0, int32_t cap_rows1, int32_t cap_msgs,
                                 int32_t n_arrays, int32_t has_row_operand) {
  if (F <= 0 || F % 4 != 0 || F > 128 || tile_rows < 4 || tile_rows % 4 != 0 || n_arrays < 1 || n_arrays > 2 ||
      cap_rows0 < 0 || cap_rows1 < 0 || cap_msgs < 0)
    return 0;
  Params p{};
  return configure(p, F, tile_rows, cap_rows0, cap_rows1, cap_msgs, n_arrays, has_row_operand != 0);
}

extern "C" int cwn_csr_gather_reduce_ws_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr,
                                            const int32_t* idx, int64_t E, const int32_t* windows, int32_t tile_rows,
                                            int32_t cap_rows, int32_t cap_msgs, int64_t n_rows, int32_t F,
                                            const float* x_res, int64_t ld_res, const float* eps, float* out,
                                            int64_t ld_out, int32_t reduce, cwn_stream_t stream) {
  int rc;
  if ((rc = check_common("cwn_csr_gather_reduce_ws_f32", n_rows, F, tile_rows, rowptr, windows, E))) return rc;
  if (reduce != CWN_REDUCE_ADD && reduce != CWN_REDUCE_MEAN) return fail(CWN_E_ENUM, "cwn ws: reduce must be add or mean");
  if (x_res && reduce != CWN_REDUCE_ADD) return fail(CWN_E_ENUM, "x_res requires CWN_REDUCE_ADD");
  if (!idx || !aligned16(idx)) return fail(CWN_E_ALIGN, "cwn ws: idx must be non-null and 16-byte aligned");
  if ((rc = check_mat(x_src, ld_src, F, "x_src")) || (rc = check_mat(out, ld_out, F, "out"))) return rc;
  if (x_res && (rc = check_mat(x_res, ld_res, F, "x_res"))) return rc;
  Params p{};
  if (configure(p, F, tile_rows, cap_rows, 0, cap_msgs, 1, x_res != nullptr) < 2)
    return fail(CWN_E_SHAPE, "cwn ws: fewer than two stages fit in shared memory");
  p.rowptr = rowptr; p.arr[0] = idx; p.arr[1] = nullptr; p.windows = windows;
  p.mat[0] = reinterpret_cast<const char*>(x_src); p.pitch[0] = (uint32_t)ld_src * 4u;
  p.rowop = reinterpret_cast<const char*>(x_res); p.pitch_ro = (uint32_t)ld_res * 4u;
  p.out = reinterpret_cast<char*>(out); p.pitch_out = (uint32_t)ld_out * 4u;
  p.eps = eps; p.n_rows = n_rows; p.E_al = E & ~(int64_t)3;
  p.n_tiles = (int32_t)((n_rows + tile_rows - 1) / tile_rows); p.FV = F / 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (reduce == CWN_REDUCE_ADD) { CWN_WS_BY_LPR(Gather<CWN_REDUCE_ADD>) }
  else { CWN_WS_BY_LPR(Gather<CWN_REDUCE_MEAN>) }
  if (rc) return rc;
  return launched("cwn_csr_gather_reduce_ws_f32");
}

extern "C" int cwn_csr_cob_fwd_ws_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q, const int32_t* rowptr,
                                      const int32_t* src, const int32_t* cob, int64_t E, const int32_t* windows,
                                      int32_t tile_rows, int32_t cap_rows0, int32_t cap_rows1, int32_t cap_msgs,
                                      int64_t n_rows, int32_t F, int32_t act, const float* x_res, int64_t ld_res,
                                      const float* eps, float* out, int64_t ld_out, cwn_stream_t stream) {
  int rc;
  if ((rc = check_common("cwn_csr_cob_fwd_ws_f32", n_rows, F, tile_rows, rowptr, windows, E))) return rc;
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (!src || !cob || !aligned16(src) || !aligned16(cob)) return fail(CWN_E_ALIGN, "cwn ws: src / cob must be non-null and 16-byte aligned");
  if ((rc = check_mat(P, ld_p, F, "P")) || (rc = check_mat(Q, ld_q, F, "Q")) || (rc = check_mat(out, ld_out, F, "out"))) return rc;
  if (x_res && (rc = check_mat(x_res, ld_res, F, "x_res"))) return rc;
  Params p{};
  if (configure(p, F, tile_rows, cap_rows0, cap_rows1, cap_msgs, 2, x_res != nullptr) < 2)
    return fail(CWN_E_SHAPE, "cwn ws: fewer than two stages fit in shared memory");
  p.rowptr = rowptr; p.arr[0] = src; p.arr[1] = cob; p.windows = windows;
  p.mat[0] = reinterpret_cast<const char*>(P); p.pitch[0] = (uint32_t)ld_p * 4u;
  p.mat[1] = reinterpret_cast<const char*>(Q); p.pitch[1] = (uint32_t)ld_q * 4u;
  p.rowop = reinterpret_cast<const char*>(x_res); p.pitch_ro = (uint32_t)ld_res * 4u;
  p.out = reinterpret_cast<char*>(out); p.pitch_out = (uint32_t)ld_out * 4u;
  p.eps = eps; p.n_rows = n_rows; p.E_al = E & ~(int64_t)3;
  p.n_tiles = (int32_t)((n_rows + tile_rows - 1) / tile_rows); p.FV = F / 4;
  cudaStream_t st = (cudaStream_t)stream;
  CWN_WS_BY_ACT(CobFwd)
  if (rc) return rc;
  return launched("cwn_csr_cob_fwd_ws_f32");
}

extern "C" int cwn_csr_cob_bwd_ws_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B,
                                      int64_t ld_b, const int32_t* rowptr, const int32_t* dst, const int32_t* oth,
                                      int64_t E, const int32_t* windows, int32_t tile_rows, int32_t cap_rows0,
                                      int32_t cap_rows1, int32_t cap_msgs, int64_t n_rows, int32_t F, int32_t act,
                                      float* gA, int64_t ld_ga, cwn_stream_t stream) {
  int rc;
  if ((rc = check_common("cwn_csr_cob_bwd_ws_f32", n_rows, F, tile_rows, rowptr, windows, E))) return rc;
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (!dst || !oth || !aligned16(dst) || !aligned16(oth)) return fail(CWN_E_ALIGN, "cwn ws: dst / oth must be non-null and 16-byte aligned");
  if ((rc = check_mat(G, ld_g, F, "G")) || (rc = check_mat(A, ld_a, F, "A")) || (rc = check_mat(B, ld_b, F, "B")) ||
      (rc = check_mat(gA, ld_ga, F, "gA")))
    return rc;
  Params p{};
  if (configure(p, F, tile_rows, cap_rows0, cap_rows1, cap_msgs, 2, true) < 2)
    return fail(CWN_E_SHAPE, "cwn ws: fewer than two stages fit in shared memory");
  p.rowptr = rowptr; p.arr[0] = dst; p.arr[1] = oth; p.windows = windows;
  p.mat[0] = reinterpret_cast<const char*>(G); p.pitch[0] = (uint32_t)ld_g * 4u;
  p.mat[1] = reinterpret_cast<const char*>(B); p.pitch[1] = (uint32_t)ld_b * 4u;
  p.rowop = reinterpret_cast<const char*>(A); p.pitch_ro = (uint32_t)ld_a * 4u;
  p.out = reinterpret_cast<char*>(gA); p.pitch_out = (uint32_t)ld_ga * 4u;
  p.eps = nullptr; p.n_rows = n_rows; p.E_al = E & ~(int64_t)3;
  p.n_tiles = (int32_t)((n_rows + tile_rows - 1) / tile_rows); p.FV = F / 4;
  cudaStream_t st = (cudaStream_t)stream;
  CWN_WS_BY_ACT(CobBwd)
  if (rc) return rc;
  return launched("cwn_csr_cob_bwd_ws_f32");
}
