// Fused gather -> message -> reduce kernels over CSR plans (one destination row per thread group, no atomics).
//
// Layout: a feature row of F fp32 is covered by LPR lanes (power of two <= 32) that each own VPL vectors of
// W floats (W = 4: 128-bit loads when F, every ld and every base pointer allow it; W = 1 otherwise, e.g. the
// F = 1 fixtures of the reference tests). A 256-thread CTA therefore owns 256/LPR destination rows; the grid is
// a multiple of the 148 SMs and strides over the rows. Inside a row the messages are visited in plan order
// (ascending original message id), four gathers in flight at a time, and accumulated sequentially, which makes
// the result independent of the launch geometry and bit-identical to a sequential CPU scatter_add_.
//
// Replaces (reference): Tensor.index_select + torch_scatter.scatter, mp/cell_mp.py:195-198 + :423-479;
// the per-message MLP of mp/layers.py:210-211,290-293 (in its split-weight form, see include/cwn_b200.h).
#include <stdlib.h>
#include "common.cuh"

namespace cwn {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;  // gathers in flight per row

template <typename V> struct VecOps;
template <> struct VecOps<float4> {
  static constexpr int W = 4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float4 load(const float* p) { return ldg_f4(p); }
  static __device__ __forceinline__ float4 load(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void store(char* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  template <class Fn> static __device__ __forceinline__ float4 map2(float4 a, float4 b, Fn f) {
    return make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
  }
  template <class Fn> static __device__ __forceinline__ float4 map3(float4 a, float4 b, float4 c, Fn f) {
    return make_float4(f(a.x, b.x, c.x), f(a.y, b.y, c.y), f(a.z, b.z, c.z), f(a.w, b.w, c.w));
  }
  // IEEE round-to-nearest adds, two lanes per instruction (sm_100 FADD2): same bits as four __fadd_rn, half the
  // issue slots — these kernels are issue-bound long before they are FP32-pipe bound
  static __device__ __forceinline__ float4 add(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  // a + s * b with the product rounded before the sum (two roundings, like the reference's separate ops). The
  // products stay scalar __fmul_rn: ptxas contracts __fmul2_rn + __fadd2_rn into one FFMA2 (seen in the SASS), which
  // would round once.
  static __device__ __forceinline__ float4 add_scaled(float4 a, float s, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(__fmul_rn(s, b.x), __fmul_rn(s, b.y)));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(__fmul_rn(s, b.z), __fmul_rn(s, b.w)));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
};
template <> struct VecOps<float> {
  static constexpr int W = 1;
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float load(const char* p) { return __ldg(reinterpret_cast<const float*>(p)); }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void store(char* p, float v) { *reinterpret_cast<float*>(p) = v; }
  template <class Fn> static __device__ __forceinline__ float map2(float a, float b, Fn f) { return f(a, b); }
  template <class Fn> static __device__ __forceinline__ float map3(float a, float b, float c, Fn f) {
    return f(a, b, c);
  }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float add_scaled(float a, float s, float b) { return __fadd_rn(a, __fmul_rn(s, b)); }
};

struct AddRn { __device__ __forceinline__ float operator()(float a, float b) const { return __fadd_rn(a, b); } };
struct MaxOp { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };

// Address of row `row` seen from a lane's base pointer: ONE IMAD.WIDE.U32 (32-bit row x 32-bit pitch + 64-bit base).
// ncu on the first version of these kernels: 60 % issue-slot utilisation at 24 % of DRAM throughput, ~95 SASS
// instructions per two-message trip of which 24 were FP and 8 loads — the rest was 64-bit index arithmetic
// (int64 row x int64 ld), -1 sentinels carried as 64-bit values, and per-vector column predicates. Rows and pitches
// are validated to fit 32 bits by the entry points.
// (lane bases are passed through `opaque()` so that the compiler keeps base + lane offset in one 64-bit register instead
// of re-adding the kernel parameter after every multiply)
template <class T> __device__ __forceinline__ T* opaque(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ const char* row_at(const char* lane_base, uint32_t row, uint32_t pitch_bytes) {
  uint64_t a;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(row), "r"(pitch_bytes), "l"(lane_base));
  return reinterpret_cast<const char*>(a);
}
__device__ __forceinline__ char* row_at(char* lane_base, uint32_t row, uint32_t pitch_bytes) {
  return lane_base + (uint64_t)row * pitch_bytes;
}

// ---------------------------------------------------------------------------------------------- row bodies
// One destination row, one column block; messages [beg, end) visited in plan order, U gathers (x operands) in flight.
// `idx_at(m)` yields the payload of message m (global plan, or its shared-memory copy in the tile-staged kernels);
// `ld(row, k)` yields vector k of a lane's share of feature row `row` (global memory, or the feature window a
// tile-staged kernel holds in shared memory). `live1`: the second vector of a lane (VPL == 2) is inside the row.

template <typename V, int LPR>
struct GlobalRows {  // rows of a matrix in global memory, read through the read-only path
  const char* base;  // lane base: matrix + lane's column offset
  uint32_t pitch;
  __device__ __forceinline__ V operator()(uint32_t row, int k) const {
    return VecOps<V>::load(row_at(base, row, pitch) + k * LPR * (int)sizeof(V));
  }
};

template <typename V, int LPR>
struct SharedRows;  // rows of a feature window staged in shared memory (128-bit path only)
template <int LPR>
struct SharedRows<float4, LPR> {
  uint32_t base;  // shared-space address of the lane's share of row 0 of the MATRIX (window base - lo * rowbytes)
  uint32_t rowbytes;
  __device__ __forceinline__ float4 operator()(uint32_t row, int k) const {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(base + row * rowbytes + (uint32_t)(k * LPR * 16)));
    return v;
  }
};

template <typename V, int VPL, int REDUCE, int U, class Ld, class Idx>
__device__ __forceinline__ void gather_row(V (&acc)[VPL], int beg, int end, Ld ld, Idx idx_at, bool live1) {
  using O = VecOps<V>;
  for (int i = beg; i < end; i += U) {
    V v[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
        const uint32_t s_ = (uint32_t)idx_at(i + u);
        v[u][0] = ld(s_, 0);
        if (VPL == 2 && live1) v[u][VPL - 1] = ld(s_, 1);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (REDUCE == CWN_REDUCE_MAX) acc[k] = (i + u == beg) ? v[u][k] : O::map2(acc[k], v[u][k], MaxOp());
          else acc[k] = O::add(acc[k], v[u][k]);
        }
      }
    }
  }
}

template <typename V, int VPL, int ACT, int U, class LdP, class LdQ, class IdxS, class IdxQ>
__device__ __forceinline__ void cob_fwd_row(V (&acc)[VPL], int beg, int end, LdP ldp, LdQ ldq, IdxS src_at,
                                            IdxQ cob_at, bool live1) {
  using O = VecOps<V>;
  for (int i = beg; i < end; i += U) {
    V vp[U][VPL], vq[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
        const uint32_t s_ = (uint32_t)src_at(i + u), q_ = (uint32_t)cob_at(i + u);
        vp[u][0] = ldp(s_, 0);
        vq[u][0] = ldq(q_, 0);
        if (VPL == 2 && live1) {
          vp[u][VPL - 1] = ldp(s_, 1);
          vq[u][VPL - 1] = ldq(q_, 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          V pre = O::add(vp[u][k], vq[u][k]);
          pre = O::map2(pre, pre, [](float x, float) { return act_fwd<ACT>(x); });
          acc[k] = O::add(acc[k], pre);
        }
      }
    }
  }
}

template <typename V, int VPL, int ACT, int U, class LdG, class LdB, class IdxT, class IdxO>
__device__ __forceinline__ void cob_bwd_row(V (&acc)[VPL], const V (&a)[VPL], int beg, int end, LdG ldg_, LdB ldb,
                                            IdxT dst_at, IdxO oth_at, bool live1) {
  using O = VecOps<V>;
  for (int i = beg; i < end; i += U) {
    V vg[U][VPL], vb[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
        const uint32_t t_ = (uint32_t)dst_at(i + u), o_ = (uint32_t)oth_at(i + u);
        vg[u][0] = ldg_(t_, 0);
        vb[u][0] = ldb(o_, 0);
        if (VPL == 2 && live1) {
          vg[u][VPL - 1] = ldg_(t_, 1);
          vb[u][VPL - 1] = ldb(o_, 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          V pre = O::add(a[k], vb[u][k]);
          if (ACT == CWN_ACT_RELU) {  // g * {0,1} as a select (what torch's threshold_backward does)
            acc[k] = O::add(acc[k], O::map2(vg[u][k], pre, [](float g_, float z) { return z > 0.f ? g_ : 0.f; }));
          } else {
            V d = O::map2(pre, pre, [](float z, float) { return act_bwd<ACT>(z); });
            acc[k] = O::add(acc[k], O::map2(vg[u][k], d, [](float g_, float d_) { return __fmul_rn(g_, d_); }));
          }
        }
      }
    }
  }
}

// out[r] = (1+eps) * x_res[r] + REDUCE_i x_src[idx ? idx[i] : i]
template <typename V, int LPR, int VPL, int REDUCE>
__global__ void __launch_bounds__(kThreads)
csr_gather_reduce_kernel(const float* __restrict__ x_src, int64_t ld_src, const int32_t* __restrict__ rowptr,
                         const int32_t* __restrict__ idx, int64_t n_rows, int FV /* F / W */,
                         const float* __restrict__ x_res, int64_t ld_res, const float* __restrict__ eps,
                         float* __restrict__ out, int64_t ld_out, const float* __restrict__ x_res2 = nullptr,
                         int64_t ld_res2 = 0, const float* __restrict__ eps2 = nullptr) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using O = VecOps<V>;
  constexpr int RPB = kThreads / LPR;
  const int lane = threadIdx.x % LPR;
  const int sub = threadIdx.x / LPR;
  const float scale = x_res ? __fadd_rn(1.f, eps ? __ldg(eps) : 0.f) : 0.f;
  const float scale2 = x_res2 ? __fadd_rn(1.f, eps2 ? __ldg(eps2) : 0.f) : 0.f;
  const uint32_t pitch_s = (uint32_t)ld_src * 4u, pitch_r = (uint32_t)ld_res * 4u, pitch_o = (uint32_t)ld_out * 4u;
  const uint32_t pitch_r2 = (uint32_t)ld_res2 * 4u;
  for (int c0 = 0; c0 < FV; c0 += LPR * VPL) {  // one trip unless F > 4*32*VPL
    const int c = c0 + lane;
    if (c >= FV) return;  // (no warp-level primitive below: idle lanes of a narrow row may leave)
    const bool live1 = VPL == 2 && c + LPR < FV;
    const char* xs = opaque(reinterpret_cast<const char*>(x_src) + (size_t)c * sizeof(V));
    const char* xr = opaque(reinterpret_cast<const char*>(x_res) + (size_t)c * sizeof(V));
    const char* xr2 = opaque(reinterpret_cast<const char*>(x_res2) + (size_t)c * sizeof(V));
    char* xo = reinterpret_cast<char*>(out) + (size_t)c * sizeof(V);
    for (int64_t r = (int64_t)blockIdx.x * RPB + sub; r < n_rows; r += (int64_t)gridDim.x * RPB) {
      const int beg = rowptr ? __ldg(rowptr + r) : 0, end = rowptr ? __ldg(rowptr + r + 1) : 0;
      V acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
      const GlobalRows<V, LPR> ldx{xs, pitch_s};
      constexpr int U = (VPL == 1) ? kUnroll : kUnroll / 2;  // same number of 128-bit gathers in flight per lane
      if (idx) gather_row<V, VPL, REDUCE, U>(acc, beg, end, ldx, [&](int m) { return __ldg(idx + m); }, live1);
      else gather_row<V, VPL, REDUCE, U>(acc, beg, end, ldx, [](int m) { return m; }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        V a = acc[k];
        if (REDUCE == CWN_REDUCE_MEAN) {
          const float cnt = (float)max(end - beg, 1);
          a = O::map2(a, a, [cnt](float x, float) { return __fdiv_rn(x, cnt); });
        }
        // GIN residual added after the aggregation, as mp/layers.py:191-192 does
        if (x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, (uint32_t)r, pitch_r) + k * LPR * (int)sizeof(V)));
        if (x_res2) a = O::add_scaled(a, scale2, O::load(row_at(xr2, (uint32_t)r, pitch_r2) + k * LPR * (int)sizeof(V)));
        O::store(row_at(xo, (uint32_t)r, pitch_o) + k * LPR * (int)sizeof(V), a);
      }
    }
  }
}

// Same contract as csr_gather_reduce_kernel, restructured for the HBM-bound regime (ncu on the first version: 60 %
// issue-slot utilisation, 27 % DRAM throughput — the per-row scalar loads of rowptr/idx and the row-at-a-time gather
// window were the cost). A group of LPR lanes now owns LPR CONSECUTIVE rows: the row pointers and the message indices
// of the whole chunk are fetched with coalesced loads (one element per lane) and broadcast by warp shuffles, and the
// 128-bit gathers are issued four at a time ACROSS row boundaries. Messages are still accumulated strictly in plan
// order, one row after the other, so results stay bit-identical to the sequential definition.
template <typename V, int LPR, int REDUCE>
__global__ void __launch_bounds__(kThreads)
csr_gather_reduce_chunked_kernel(const float* __restrict__ x_src, int64_t ld_src, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ idx, int64_t n_rows, int FV,
                                 const float* __restrict__ x_res, int64_t ld_res, const float* __restrict__ eps,
                                 float* __restrict__ out, int64_t ld_out) {
  using O = VecOps<V>;
  constexpr int GPB = kThreads / LPR;  // groups per CTA
  const int lane = threadIdx.x % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x % 32) / LPR * LPR));
  const float scale = x_res ? __fadd_rn(1.f, eps ? __ldg(eps) : 0.f) : 0.f;
  const bool col_live = lane < FV;
  const int64_t n_chunks = (n_rows + LPR - 1) / LPR;
  for (int64_t chunk = (int64_t)blockIdx.x * GPB + threadIdx.x / LPR; chunk < n_chunks; chunk += (int64_t)gridDim.x * GPB) {
    const int64_t r0 = chunk * LPR;
    const int nrow = (int)((n_rows - r0 < LPR) ? n_rows - r0 : LPR);
    const int my_beg = (lane < nrow) ? __ldg(rowptr + r0 + lane) : 0;
    const int my_end = (lane < nrow) ? __ldg(rowptr + r0 + lane + 1) : 0;
    const int m0 = __shfl_sync(gmask, my_beg, 0, LPR);
    const int m1 = __shfl_sync(gmask, my_end, nrow - 1, LPR);
    int cur = 0;
    int cur_end = __shfl_sync(gmask, my_end, 0, LPR);
    int cur_beg = m0;
    V acc = O::zero();

    auto flush = [&]() {  // finish row `cur`
      V a = acc;
      if (REDUCE == CWN_REDUCE_MEAN) {
        const float cnt = (float)max(cur_end - cur_beg, 1);
        a = O::map2(a, a, [cnt](float v, float) { return __fdiv_rn(v, cnt); });
      }
      if (col_live) {
        const int64_t r = r0 + cur;
        if (x_res) {
          V xr = O::load(x_res + r * ld_res + (int64_t)lane * O::W);
          a = O::map2(a, xr, [scale](float s_, float x) { return __fadd_rn(s_, __fmul_rn(scale, x)); });
        }
        O::store(out + r * ld_out + (int64_t)lane * O::W, a);
      }
      acc = O::zero();
      ++cur;
      cur_beg = cur_end;
      cur_end = __shfl_sync(gmask, my_end, cur < LPR ? cur : LPR - 1, LPR);
    };

    for (int base = m0; base < m1; base += LPR) {
      const int cnt = (m1 - base < LPR) ? m1 - base : LPR;
      const int my_idx = (lane < cnt) ? (idx ? __ldg(idx + base + lane) : base + lane) : 0;
      for (int j0 = 0; j0 < cnt; j0 += kUnroll) {
        V v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int s = __shfl_sync(gmask, my_idx, (j0 + u < LPR) ? j0 + u : LPR - 1, LPR);
          if (j0 + u < cnt && col_live) v[u] = O::load(x_src + (int64_t)s * ld_src + (int64_t)lane * O::W);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int m = base + j0 + u;
          if (j0 + u >= cnt) break;
          while (m >= cur_end) flush();  // also writes the (zero / residual-only) rows that have no message
          if (col_live) {
            if (REDUCE == CWN_REDUCE_MAX) acc = (m == cur_beg) ? v[u] : O::map2(acc, v[u], MaxOp());
            else acc = O::map2(acc, v[u], AddRn());
          }
        }
      }
    }
    while (cur < nrow) flush();
  }
}

// out[r] = (1+eps) * x_res[r] + SUM_i act(P[src[i]] + Q[cob[i]])
template <typename V, int LPR, int VPL, int ACT>
__global__ void __launch_bounds__(kThreads)
csr_cob_fwd_kernel(const float* __restrict__ P, int64_t ld_p, const float* __restrict__ Q, int64_t ld_q,
                   const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                   const int32_t* __restrict__ cob, int64_t n_rows, int FV, const float* __restrict__ x_res,
                   int64_t ld_res, const float* __restrict__ eps, float* __restrict__ out, int64_t ld_out) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using O = VecOps<V>;
  constexpr int RPB = kThreads / LPR;
  constexpr int U = (VPL == 1) ? 4 : 2;  // two operands per message: 2*U*VPL gathers in flight
  const int lane = threadIdx.x % LPR;
  const int sub = threadIdx.x / LPR;
  const float scale = x_res ? __fadd_rn(1.f, eps ? __ldg(eps) : 0.f) : 0.f;
  const uint32_t pitch_p = (uint32_t)ld_p * 4u, pitch_q = (uint32_t)ld_q * 4u, pitch_r = (uint32_t)ld_res * 4u,
                 pitch_o = (uint32_t)ld_out * 4u;
  for (int c0 = 0; c0 < FV; c0 += LPR * VPL) {
    const int c = c0 + lane;
    if (c >= FV) return;
    const bool live1 = VPL == 2 && c + LPR < FV;
    const char* pl = opaque(reinterpret_cast<const char*>(P) + (size_t)c * sizeof(V));
    const char* ql = opaque(reinterpret_cast<const char*>(Q) + (size_t)c * sizeof(V));
    const char* xr = opaque(reinterpret_cast<const char*>(x_res) + (size_t)c * sizeof(V));
    char* xo = reinterpret_cast<char*>(out) + (size_t)c * sizeof(V);
    for (int64_t r = (int64_t)blockIdx.x * RPB + sub; r < n_rows; r += (int64_t)gridDim.x * RPB) {
      const int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
      V acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
      cob_fwd_row<V, VPL, ACT, U>(acc, beg, end, GlobalRows<V, LPR>{pl, pitch_p}, GlobalRows<V, LPR>{ql, pitch_q},
                                  [&](int m) { return __ldg(src + m); }, [&](int m) { return __ldg(cob + m); }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        V a = acc[k];
        if (x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, (uint32_t)r, pitch_r) + k * LPR * (int)sizeof(V)));
        O::store(row_at(xo, (uint32_t)r, pitch_o) + k * LPR * (int)sizeof(V), a);
      }
    }
  }
}

// gA[r] = SUM_i G[dst[i]] * act'(A[r] + B[oth[i]])
template <typename V, int LPR, int VPL, int ACT>
__global__ void __launch_bounds__(kThreads)
csr_cob_bwd_kernel(const float* __restrict__ G, int64_t ld_g, const float* __restrict__ A, int64_t ld_a,
                   const float* __restrict__ B, int64_t ld_b, const int32_t* __restrict__ rowptr,
                   const int32_t* __restrict__ dst, const int32_t* __restrict__ oth, int64_t n_rows, int FV,
                   float* __restrict__ gA, int64_t ld_ga) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using O = VecOps<V>;
  constexpr int RPB = kThreads / LPR;
  constexpr int U = (VPL == 1) ? 4 : 2;
  const int lane = threadIdx.x % LPR;
  const int sub = threadIdx.x / LPR;
  const uint32_t pitch_g = (uint32_t)ld_g * 4u, pitch_a = (uint32_t)ld_a * 4u, pitch_b = (uint32_t)ld_b * 4u,
                 pitch_o = (uint32_t)ld_ga * 4u;
  for (int c0 = 0; c0 < FV; c0 += LPR * VPL) {
    const int c = c0 + lane;
    if (c >= FV) return;
    const bool live1 = VPL == 2 && c + LPR < FV;
    const char* gl = opaque(reinterpret_cast<const char*>(G) + (size_t)c * sizeof(V));
    const char* al = opaque(reinterpret_cast<const char*>(A) + (size_t)c * sizeof(V));
    const char* bl = opaque(reinterpret_cast<const char*>(B) + (size_t)c * sizeof(V));
    char* xo = reinterpret_cast<char*>(gA) + (size_t)c * sizeof(V);
    for (int64_t r = (int64_t)blockIdx.x * RPB + sub; r < n_rows; r += (int64_t)gridDim.x * RPB) {
      const int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
      V acc[VPL], a[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        acc[k] = O::zero();
        a[k] = (beg < end && (k == 0 || live1)) ? O::load(row_at(al, (uint32_t)r, pitch_a) + k * LPR * (int)sizeof(V))
                                                : O::zero();
      }
      cob_bwd_row<V, VPL, ACT, U>(acc, a, beg, end, GlobalRows<V, LPR>{gl, pitch_g}, GlobalRows<V, LPR>{bl, pitch_b},
                                  [&](int m) { return __ldg(dst + m); }, [&](int m) { return __ldg(oth + m); }, live1);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k == 1 && !live1) break;
        O::store(row_at(xo, (uint32_t)r, pitch_o) + k * LPR * (int)sizeof(V), acc[k]);
      }
    }
  }
}

// out[e] = scale * x[idx[e]]  (int64 API indices; one message row per thread group)
template <typename V, int LPR>
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float* __restrict__ x, int64_t ld_x, const int64_t* __restrict__ idx, int64_t E, int FV,
                   float scale, float* __restrict__ out, int64_t ld_out) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using O = VecOps<V>;
  constexpr int RPB = kThreads / LPR;
  const int lane = threadIdx.x % LPR;
  const int sub = threadIdx.x / LPR;
  for (int64_t e = (int64_t)blockIdx.x * RPB + sub; e < E; e += (int64_t)gridDim.x * RPB) {
    const int64_t s = __ldg(idx + e);
    for (int c = lane; c < FV; c += LPR) {
      V v = O::load(x + s * ld_x + (int64_t)c * O::W);
      if (scale != 1.f) v = O::map2(v, v, [scale](float a, float) { return __fmul_rn(scale, a); });
      O::store(out + e * ld_out + (int64_t)c * O::W, v);
    }
  }
}

}  // namespace cwn
#include "gsa_tiled.cuh"
#include "gsa_pipelined.cuh"
namespace cwn {

// ---- max aggregation with argument tracking (backward support). torch_scatter's CPU scatter_max keeps the FIRST
// maximum in message order (strict '>'), which is what a sequential walk in plan order does; the backward routes the
// gradient of (row, feature) to that one message. One thread per (row, feature), coalesced across features.
__global__ void csr_gather_max_arg_kernel(const float* __restrict__ x_src, int64_t ld_src,
                                          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ idx,
                                          const int32_t* __restrict__ perm, int64_t n_rows, int F,
                                          float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ arg) {
  const int64_t total = n_rows * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / F;
    const int f = (int)(e - r * F);
    const int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
    float best = 0.f;  // rows without messages stay zero, as torch_scatter fills them
    int32_t who = -1;
    for (int i = beg; i < end; ++i) {
      const float v = __ldg(x_src + (int64_t)__ldg(idx + i) * ld_src + f);
      if (i == beg || v > best) {
        best = v;
        who = perm ? __ldg(perm + i) : i;
      }
    }
    out[r * ld_out + f] = best;
    arg[e] = who;
  }
}

// gX[s, f] = SUM_{i in row s of the by-source plan} (arg[dst[i], f] == message id of i) ? G[dst[i], f] : 0
__global__ void csr_max_bwd_kernel(const float* __restrict__ G, int64_t ld_g, const int32_t* __restrict__ arg,
                                   const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dst,
                                   const int32_t* __restrict__ perm, int64_t n_rows, int F, float* __restrict__ gX,
                                   int64_t ld_gx) {
  const int64_t total = n_rows * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s_ = e / F;
    const int f = (int)(e - s_ * F);
    const int beg = __ldg(rowptr + s_), end = __ldg(rowptr + s_ + 1);
    float acc = 0.f;
    for (int i = beg; i < end; ++i) {
      const int64_t t = __ldg(dst + i);
      const int32_t id = perm ? __ldg(perm + i) : i;
      if (__ldg(arg + t * F + f) == id) acc = __fadd_rn(acc, __ldg(G + t * ld_g + f));
    }
    gX[s_ * ld_gx + f] = acc;
  }
}

__global__ void check_index_range_kernel(const int64_t* __restrict__ idx, int64_t E, int64_t n, int32_t* flags) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx[e];
    if (v < 0 || v >= n) atomicOr(flags, 2);
  }
}

// ------------------------------------------------------------------------------------------------ dispatch
struct Geometry {
  bool vec;  // 128-bit path
  int fv;    // F / W
  int lpr;   // lanes per row
  int vpl;   // vectors per lane
};

static Geometry geometry(int F, bool can_vec) {
  Geometry g;
  g.vec = can_vec && (F % 4 == 0);
  g.fv = g.vec ? F / 4 : F;
  int lpr = 1;
  while (lpr < g.fv && lpr < 32) lpr <<= 1;
  g.lpr = lpr;
  g.vpl = (g.fv > 32) ? 2 : 1;
  return g;
}

static int grid_for(int64_t rows, int lpr) {
  const int64_t rpb = kThreads / lpr;
  int64_t need = (rows + rpb - 1) / rpb;
  const int64_t cap = (int64_t)kNumSMs * 16;  // 8 resident CTAs/SM x 2 waves, multiple of the SM count
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

#define CWN_DISPATCH_LPR(LPRV, ...)                    \
  switch (LPRV) {                                      \
    case 1: { constexpr int LPR = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int LPR = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int LPR = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int LPR = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int LPR = 16; __VA_ARGS__; } break; \
    default: { constexpr int LPR = 32; __VA_ARGS__; } break; \
  }

#define CWN_DISPATCH_ACT(ACTV, ...)                                            \
  switch (ACTV) {                                                              \
    case CWN_ACT_ID: { constexpr int ACT = CWN_ACT_ID; __VA_ARGS__; } break;       \
    case CWN_ACT_RELU: { constexpr int ACT = CWN_ACT_RELU; __VA_ARGS__; } break;   \
    case CWN_ACT_ELU: { constexpr int ACT = CWN_ACT_ELU; __VA_ARGS__; } break;     \
    case CWN_ACT_SIGMOID: { constexpr int ACT = CWN_ACT_SIGMOID; __VA_ARGS__; } break; \
    default: { constexpr int ACT = CWN_ACT_TANH; __VA_ARGS__; } break;             \
  }

static int check_matrix(const float* p, int64_t ld, int F, const char* name) {
  if (!p) return fail(CWN_E_NULL, name);
  if (ld < F) return fail(CWN_E_SHAPE, "leading dimension smaller than F");
  if (!aligned4(p)) return fail(CWN_E_ALIGN, name);
  return CWN_OK;
}

static bool vec_ok(const float* p, int64_t ld) { return p == nullptr || (aligned16(p) && ld % 4 == 0); }

static bool getenv_flag(const char* name) {  // A/B switches for profiling; read once
  const char* v = getenv(name);
  return v && v[0] == '1';
}


// Launch geometry of a tile-staged kernel: dynamic shared memory = the feature-window buffers (`n_ops` operands of
// F floats per row); persistent grid = resident CTAs per SM x 148, capped by the number of tiles. The budget aims at
// 3 CTAs per SM (24 warps: the gathers hit shared memory, the DRAM stream is driven by TMA), 2 when rows are wider
// than 256 bytes. (The occupancy query is a host-side table lookup, legal during stream capture.)
struct TiledLaunch {
  int grid;
  uint32_t cap;    // bytes per operand buffer
  size_t dynamic;  // total dynamic shared memory
};
template <class K>
static TiledLaunch tiled_launch(K kernel, int64_t n_rows, int lpr, int F, int n_ops) {
  TiledLaunch t;
  const size_t budget = (size_t)F * 4 <= 256 ? 54 * 1024 : 92 * 1024;
  static const bool no_features = getenv_flag("CWN_B200_TILED_NO_FEATURES");  // A/B: stage the plan slices only
  t.cap = no_features ? 0u : (uint32_t)((budget / n_ops) & ~(size_t)127);
  t.dynamic = (size_t)t.cap * n_ops;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.dynamic);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, t.dynamic) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int gpb = kThreads / lpr;
  const int64_t tr = (4 * gpb < 64) ? 64 : ((4 * gpb > 128) ? 128 : 4 * gpb);  // TileRows<LPR>
  int64_t tiles = (n_rows + tr - 1) / tr;
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  if (tiles > cap) tiles = cap;
  t.grid = (int)(tiles < 1 ? 1 : tiles);
  return t;
}

// Which kernel family serves the HBM-bound regime (>= kTiledMinRows rows, vector path): 'r' plain rows / chunked,
// 'p' software-pipelined rows, 't' tile-staged (TMA bulk copies of the plan and of the feature windows).
// Defaults = the fastest measured at 1M rows per dimension, block-diagonal ZINC-like layout (profiles/README.md):
//   identity pass  F <= 128: chunked 0.57-0.67 of the HBM peak | pipelined 0.53-0.60 | tile-staged 0.49-0.52
//                  F  = 256: pipelined 1.03 (L2 reuse of shared sources) | rows 0.78 | tile-staged 0.64
//   coboundary fwd F = 64:   rows 0.52 | pipelined (U=2) 0.50 | tile-staged 0.35-0.44
//   coboundary bwd F = 64:   pipelined (U=2) 0.66 | rows 0.61 | tile-staged 0.35-0.38
// A/B switches for profiling: CWN_B200_LARGE_GATHER, CWN_B200_LARGE_COB = r | p | t.
static char large_mode_env(const char* env) {
  const char* v = getenv(env);
  return (v && (v[0] == 't' || v[0] == 'p' || v[0] == 'r')) ? v[0] : 0;
}
// (the environment is read ONCE per switch: these sit on the launch path of every adjacency pass)
static char large_mode(const char* env, char dflt) {
  static const char gather = large_mode_env("CWN_B200_LARGE_GATHER"), cob = large_mode_env("CWN_B200_LARGE_COB");
  const char v = env[15] == 'G' ? gather : cob;  // "CWN_B200_LARGE_G..." / "CWN_B200_LARGE_C..."
  return v ? v : dflt;
}

template <class Pass>
static void launch_pipelined(const typename Pass::Params& prm, const int32_t* rowptr, int64_t n_rows, cudaStream_t st) {
  auto kern = csr_pipelined_kernel<Pass>;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t rpb = kThreads / Pass::LPR;
  int64_t need = (n_rows + rpb - 1) / rpb;
  const int64_t cap = (int64_t)kNumSMs * per_sm;  // persistent: every resident group walks a strided row sequence
  if (need > cap) need = cap;
  kern<<<(int)need, kThreads, 0, st>>>(prm, rowptr, n_rows);
}

static bool tiled_enabled() {
  static const bool off = getenv_flag("CWN_B200_NO_TILED");
  return !off;
}
constexpr int64_t kTiledMinRows = 32768;
}  // namespace cwn

using namespace cwn;

extern "C" int cwn_csr_gather_reduce_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr,
                                         const int32_t* idx, int64_t n_rows, int32_t F, const float* x_res,
                                         int64_t ld_res, const float* eps, float* out, int64_t ld_out,
                                         int32_t reduce, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_gather_reduce_f32: bad n_rows/F");
  if (reduce < CWN_REDUCE_ADD || reduce > CWN_REDUCE_MAX) return fail(CWN_E_ENUM, "unknown reduce");
  if (x_res && reduce != CWN_REDUCE_ADD) return fail(CWN_E_ENUM, "x_res requires CWN_REDUCE_ADD");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr) return fail(CWN_E_NULL, "rowptr");
  int rc;
  if ((rc = check_matrix(out, ld_out, F, "out"))) return rc;
  // x_src may be NULL only if there are no messages at all; the kernel never dereferences it then
  if (x_src && (rc = check_matrix(x_src, ld_src, F, "x_src"))) return rc;
  if (x_res && (rc = check_matrix(x_res, ld_res, F, "x_res"))) return rc;
  const Geometry g = geometry(F, vec_ok(x_src, ld_src) && vec_ok(x_res, ld_res) && vec_ok(out, ld_out));
  const int grid = grid_for(n_rows, g.lpr);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(VT, VPLV, RED)                                                                                   \
  launch_pdl((csr_gather_reduce_kernel<VT, LPR, VPLV, RED>), grid, kThreads, 0, st, x_src, ld_src, rowptr, idx, n_rows, \
             g.fv, x_res, ld_res, eps, out, ld_out, (const float*)nullptr, (int64_t)0, (const float*)nullptr)
#define BY_REDUCE(VT, VPLV)                                              \
  if (reduce == CWN_REDUCE_ADD) LAUNCH(VT, VPLV, CWN_REDUCE_ADD);        \
  else if (reduce == CWN_REDUCE_MEAN) LAUNCH(VT, VPLV, CWN_REDUCE_MEAN); \
  else LAUNCH(VT, VPLV, CWN_REDUCE_MAX)
#define LAUNCH_CHUNKED(VT, RED)                                                                                        \
  csr_gather_reduce_chunked_kernel<VT, LPR, RED><<<grid_c, kThreads, 0, st>>>(x_src, ld_src, rowptr, idx, n_rows, g.fv, \
                                                                              x_res, ld_res, eps, out, ld_out)
#define BY_REDUCE_CHUNKED(VT)                                              \
  if (reduce == CWN_REDUCE_ADD) LAUNCH_CHUNKED(VT, CWN_REDUCE_ADD);        \
  else if (reduce == CWN_REDUCE_MEAN) LAUNCH_CHUNKED(VT, CWN_REDUCE_MEAN); \
  else LAUNCH_CHUNKED(VT, CWN_REDUCE_MAX)
  // One group of LPR lanes per LPR consecutive rows. Only worth it when there are enough rows to fill the machine
  // with chunks (the HBM-bound regime); small batches keep one row per group for latency.
  const int64_t chunks = (n_rows + g.lpr - 1) / g.lpr;
  const int grid_c = grid_for(chunks, g.lpr);
  static const bool gather_v1 = getenv_flag("CWN_B200_GATHER_V1"), no_tiled_gather = getenv_flag("CWN_B200_NO_TILED_GATHER");
  const bool chunked = g.vpl == 1 && g.lpr >= 4 && n_rows >= 32768 && !gather_v1;
  const bool tiled = g.vec && g.lpr >= 4 && idx && n_rows >= kTiledMinRows && aligned16(rowptr) && aligned16(idx) &&
                     g.fv <= g.lpr * g.vpl && tiled_enabled() && !no_tiled_gather;
#define LAUNCH_TILED(VPLV, RED)                                                                                   \
  {                                                                                                               \
    auto kern = csr_gather_reduce_tiled_kernel<float4, LPR, VPLV, RED>;                                           \
    const TiledLaunch tl = tiled_launch(kern, n_rows, LPR, F, 1);                                                 \
    kern<<<tl.grid, kThreads, tl.dynamic, st>>>(x_src, ld_src, rowptr, idx, n_rows, g.fv, x_res, ld_res, eps, out, \
                                                ld_out, tl.cap);                                                  \
  }
#define BY_REDUCE_TILED(VPLV)                                              \
  if (reduce == CWN_REDUCE_ADD) LAUNCH_TILED(VPLV, CWN_REDUCE_ADD)         \
  else if (reduce == CWN_REDUCE_MEAN) LAUNCH_TILED(VPLV, CWN_REDUCE_MEAN)  \
  else LAUNCH_TILED(VPLV, CWN_REDUCE_MAX)
  const char mode = large_mode("CWN_B200_LARGE_GATHER", g.vpl == 2 ? 'p' : 'r');
  const bool large_ok = g.vec && g.lpr >= 4 && idx && n_rows >= kTiledMinRows && g.fv <= g.lpr * g.vpl;
#define LAUNCH_PIPE(VPLV, RED)                                                                                       \
  {                                                                                                                  \
    using Pass = GatherPass<float4, LPR, VPLV, RED>;                                                                 \
    Pass::Params prm{x_src, idx, x_res, eps, out, (uint32_t)ld_src * 4u, (uint32_t)ld_res * 4u, (uint32_t)ld_out * 4u, g.fv}; \
    launch_pipelined<Pass>(prm, rowptr, n_rows, st);                                                                 \
  }
#define BY_REDUCE_PIPE(VPLV)                                             \
  if (reduce == CWN_REDUCE_ADD) LAUNCH_PIPE(VPLV, CWN_REDUCE_ADD)        \
  else if (reduce == CWN_REDUCE_MEAN) LAUNCH_PIPE(VPLV, CWN_REDUCE_MEAN) \
  else LAUNCH_PIPE(VPLV, CWN_REDUCE_MAX)
  if (large_ok && mode == 'p') {
    if (g.vpl == 2) { constexpr int LPR = 32; BY_REDUCE_PIPE(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; BY_REDUCE_PIPE(1) } break;
      case 8: { constexpr int LPR = 8; BY_REDUCE_PIPE(1) } break;
      case 16: { constexpr int LPR = 16; BY_REDUCE_PIPE(1) } break;
      default: { constexpr int LPR = 32; BY_REDUCE_PIPE(1) } break;
    }
  } else if (tiled && mode == 't') {
    if (g.vpl == 2) { constexpr int LPR = 32; BY_REDUCE_TILED(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; BY_REDUCE_TILED(1) } break;
      case 8: { constexpr int LPR = 8; BY_REDUCE_TILED(1) } break;
      case 16: { constexpr int LPR = 16; BY_REDUCE_TILED(1) } break;
      default: { constexpr int LPR = 32; BY_REDUCE_TILED(1) } break;
    }
  } else if (g.vec) {
    if (chunked) {
      switch (g.lpr) {
        case 4: { constexpr int LPR = 4; BY_REDUCE_CHUNKED(float4); } break;
        case 8: { constexpr int LPR = 8; BY_REDUCE_CHUNKED(float4); } break;
        case 16: { constexpr int LPR = 16; BY_REDUCE_CHUNKED(float4); } break;
        default: { constexpr int LPR = 32; BY_REDUCE_CHUNKED(float4); } break;
      }
    } else if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, BY_REDUCE(float4, 1)) } else { constexpr int LPR = 32; BY_REDUCE(float4, 2); }
  } else {
    if (chunked) {
      switch (g.lpr) {
        case 4: { constexpr int LPR = 4; BY_REDUCE_CHUNKED(float); } break;
        case 8: { constexpr int LPR = 8; BY_REDUCE_CHUNKED(float); } break;
        case 16: { constexpr int LPR = 16; BY_REDUCE_CHUNKED(float); } break;
        default: { constexpr int LPR = 32; BY_REDUCE_CHUNKED(float); } break;
      }
    } else if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, BY_REDUCE(float, 1)) } else { constexpr int LPR = 32; BY_REDUCE(float, 2); }
  }
#undef BY_REDUCE_PIPE
#undef LAUNCH_PIPE
#undef BY_REDUCE_TILED
#undef LAUNCH_TILED
#undef BY_REDUCE_CHUNKED
#undef LAUNCH_CHUNKED
#undef BY_REDUCE
#undef LAUNCH
  return launched("cwn_csr_gather_reduce_f32");
}

extern "C" int cwn_csr_gather_reduce2_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr,
                                          const int32_t* idx, int64_t n_rows, int32_t F, const float* x_res,
                                          int64_t ld_res, const float* eps, const float* x_res2, int64_t ld_res2,
                                          const float* eps2, float* out, int64_t ld_out, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_gather_reduce2_f32: bad n_rows/F");
  if (n_rows == 0) return CWN_OK;
  int rc;
  if ((rc = check_matrix(out, ld_out, F, "out"))) return rc;
  if (x_src && (rc = check_matrix(x_src, ld_src, F, "x_src"))) return rc;
  if (x_res && (rc = check_matrix(x_res, ld_res, F, "x_res"))) return rc;
  if (x_res2 && (rc = check_matrix(x_res2, ld_res2, F, "x_res2"))) return rc;
  if (rowptr && !x_src) rowptr = nullptr;  // no source matrix: there can be no messages
  const Geometry g = geometry(F, vec_ok(x_src, ld_src) && vec_ok(x_res, ld_res) && vec_ok(x_res2, ld_res2) &&
                                     vec_ok(out, ld_out));
  const int grid = grid_for(n_rows, g.lpr);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(VT, VPLV)                                                                                               \
  launch_pdl((csr_gather_reduce_kernel<VT, LPR, VPLV, CWN_REDUCE_ADD>), grid, kThreads, 0, st,                          \
             x_src, ld_src, rowptr, idx, n_rows, g.fv, x_res, ld_res, eps, out, ld_out, x_res2, ld_res2, eps2)
  if (g.vec) {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float4, 1)) } else { constexpr int LPR = 32; LAUNCH(float4, 2); }
  } else {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float, 1)) } else { constexpr int LPR = 32; LAUNCH(float, 2); }
  }
#undef LAUNCH
  return launched("cwn_csr_gather_reduce2_f32");
}

extern "C" int cwn_csr_gather_max_arg_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                                          const int32_t* perm, int64_t n_rows, int32_t F, float* out, int64_t ld_out,
                                          int32_t* arg, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_gather_max_arg_f32: bad n_rows/F");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr || !arg) return fail(CWN_E_NULL, "rowptr/arg");
  int rc;
  if ((rc = check_matrix(out, ld_out, F, "out"))) return rc;
  if (x_src && (rc = check_matrix(x_src, ld_src, F, "x_src"))) return rc;
  int64_t need = (n_rows * F + 255) / 256;
  if (need > (int64_t)kNumSMs * 16) need = (int64_t)kNumSMs * 16;
  csr_gather_max_arg_kernel<<<(int)need, 256, 0, (cudaStream_t)stream>>>(x_src, ld_src, rowptr, idx, perm, n_rows, F, out,
                                                                         ld_out, arg);
  return launched("cwn_csr_gather_max_arg_f32");
}

extern "C" int cwn_csr_max_bwd_f32(const float* G, int64_t ld_g, const int32_t* arg, const int32_t* rowptr,
                                   const int32_t* dst, const int32_t* perm, int64_t n_rows, int32_t F, float* gX,
                                   int64_t ld_gx, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_max_bwd_f32: bad n_rows/F");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr) return fail(CWN_E_NULL, "rowptr");
  int rc;
  if ((rc = check_matrix(gX, ld_gx, F, "gX"))) return rc;
  if (G && (rc = check_matrix(G, ld_g, F, "G"))) return rc;
  int64_t need = (n_rows * F + 255) / 256;
  if (need > (int64_t)kNumSMs * 16) need = (int64_t)kNumSMs * 16;
  csr_max_bwd_kernel<<<(int)need, 256, 0, (cudaStream_t)stream>>>(G, ld_g, arg, rowptr, dst, perm, n_rows, F, gX, ld_gx);
  return launched("cwn_csr_max_bwd_f32");
}

extern "C" int cwn_gather_rows_f32(const float* x, int64_t ld_x, const int64_t* idx, int64_t E, int32_t F,
                                   float scale, float* out, int64_t ld_out, cwn_stream_t stream) {
  if (E < 0 || F <= 0) return fail(CWN_E_SHAPE, "cwn_gather_rows_f32: bad E/F");
  if (E == 0) return CWN_OK;
  if (!idx) return fail(CWN_E_NULL, "idx");
  int rc;
  if ((rc = check_matrix(x, ld_x, F, "x"))) return rc;
  if ((rc = check_matrix(out, ld_out, F, "out"))) return rc;
  const Geometry g = geometry(F, vec_ok(x, ld_x) && vec_ok(out, ld_out));
  const int grid = grid_for(E, g.lpr);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.vec) {
    CWN_DISPATCH_LPR(g.lpr, launch_pdl((gather_rows_kernel<float4, LPR>), grid, kThreads, 0, st, x, ld_x, idx, E, g.fv, scale, out, ld_out))
  } else {
    CWN_DISPATCH_LPR(g.lpr, launch_pdl((gather_rows_kernel<float, LPR>), grid, kThreads, 0, st, x, ld_x, idx, E, g.fv, scale, out, ld_out))
  }
  return launched("cwn_gather_rows_f32");
}

extern "C" int cwn_csr_cob_fwd_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q,
                                   const int32_t* rowptr, const int32_t* src, const int32_t* cob, int64_t n_rows,
                                   int32_t F, int32_t act, const float* x_res, int64_t ld_res, const float* eps,
                                   float* out, int64_t ld_out, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_cob_fwd_f32: bad n_rows/F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr || !src || !cob) return fail(CWN_E_NULL, "plan");
  int rc;
  if ((rc = check_matrix(P, ld_p, F, "P"))) return rc;
  if ((rc = check_matrix(Q, ld_q, F, "Q"))) return rc;
  if ((rc = check_matrix(out, ld_out, F, "out"))) return rc;
  if (x_res && (rc = check_matrix(x_res, ld_res, F, "x_res"))) return rc;
  const Geometry g = geometry(F, vec_ok(P, ld_p) && vec_ok(Q, ld_q) && vec_ok(x_res, ld_res) && vec_ok(out, ld_out));
  const int grid = grid_for(n_rows, g.lpr);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(VT, VPLV)                                                                                           \
  CWN_DISPATCH_ACT(act, launch_pdl((csr_cob_fwd_kernel<VT, LPR, VPLV, ACT>), grid, kThreads, 0, st,                  \
                            P, ld_p, Q, ld_q, rowptr, src, cob, n_rows, g.fv, x_res, ld_res, eps, out, ld_out))
  // (a chunked variant like csr_gather_reduce_chunked_kernel was measured SLOWER for the two-operand passes —
  //  0.38 vs 0.40 of the HBM peak forward, 0.41 vs 0.53 backward at 1M rows, 80 registers — and was removed)
  const bool tiled = g.vec && g.lpr >= 4 && n_rows >= kTiledMinRows && aligned16(rowptr) && aligned16(src) &&
                     aligned16(cob) && g.fv <= g.lpr * g.vpl && tiled_enabled();
#define LAUNCH_TILED(VPLV)                                                                                          \
  CWN_DISPATCH_ACT(act, {                                                                                           \
    auto kern = csr_cob_fwd_tiled_kernel<float4, LPR, VPLV, ACT>;                                                   \
    const TiledLaunch tl = tiled_launch(kern, n_rows, LPR, F, 2);                                                   \
    kern<<<tl.grid, kThreads, tl.dynamic, st>>>(P, ld_p, Q, ld_q, rowptr, src, cob, n_rows, g.fv, x_res, ld_res, eps, \
                                                out, ld_out, tl.cap, tl.cap);                                       \
  })
  const char mode = large_mode("CWN_B200_LARGE_COB", 'r');
  const bool large_ok = g.vec && g.lpr >= 4 && n_rows >= kTiledMinRows && g.fv <= g.lpr * g.vpl;
#define LAUNCH_PIPE(VPLV)                                                                                            \
  CWN_DISPATCH_ACT(act, {                                                                                            \
    using Pass = CobFwdPass<float4, LPR, VPLV, ACT>;                                                                 \
    Pass::Params prm{P, Q, src, cob, x_res, eps, out, (uint32_t)ld_p * 4u, (uint32_t)ld_q * 4u, (uint32_t)ld_res * 4u, \
                     (uint32_t)ld_out * 4u, g.fv};                                                                   \
    launch_pipelined<Pass>(prm, rowptr, n_rows, st);                                                                 \
  })
  if (large_ok && mode == 'p') {
    if (g.vpl == 2) { constexpr int LPR = 32; LAUNCH_PIPE(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; LAUNCH_PIPE(1) } break;
      case 8: { constexpr int LPR = 8; LAUNCH_PIPE(1) } break;
      case 16: { constexpr int LPR = 16; LAUNCH_PIPE(1) } break;
      default: { constexpr int LPR = 32; LAUNCH_PIPE(1) } break;
    }
  } else if (tiled && mode == 't') {
    if (g.vpl == 2) { constexpr int LPR = 32; LAUNCH_TILED(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; LAUNCH_TILED(1) } break;
      case 8: { constexpr int LPR = 8; LAUNCH_TILED(1) } break;
      case 16: { constexpr int LPR = 16; LAUNCH_TILED(1) } break;
      default: { constexpr int LPR = 32; LAUNCH_TILED(1) } break;
    }
  } else if (g.vec) {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float4, 1)) } else { constexpr int LPR = 32; LAUNCH(float4, 2) }
  } else {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float, 1)) } else { constexpr int LPR = 32; LAUNCH(float, 2) }
  }
#undef LAUNCH_PIPE
#undef LAUNCH_TILED
#undef LAUNCH
  return launched("cwn_csr_cob_fwd_f32");
}

extern "C" int cwn_csr_cob_bwd_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B,
                                   int64_t ld_b, const int32_t* rowptr, const int32_t* dst, const int32_t* oth,
                                   int64_t n_rows, int32_t F, int32_t act, float* gA, int64_t ld_ga,
                                   cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_cob_bwd_f32: bad n_rows/F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr || !dst || !oth) return fail(CWN_E_NULL, "plan");
  int rc;
  if ((rc = check_matrix(A, ld_a, F, "A"))) return rc;
  if ((rc = check_matrix(gA, ld_ga, F, "gA"))) return rc;
  // G / B may be NULL only when there are no messages
  if (G && (rc = check_matrix(G, ld_g, F, "G"))) return rc;
  if (B && (rc = check_matrix(B, ld_b, F, "B"))) return rc;
  const Geometry g = geometry(F, vec_ok(G, ld_g) && vec_ok(A, ld_a) && vec_ok(B, ld_b) && vec_ok(gA, ld_ga));
  const int grid = grid_for(n_rows, g.lpr);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(VT, VPLV)                                                                  \
  CWN_DISPATCH_ACT(act, launch_pdl((csr_cob_bwd_kernel<VT, LPR, VPLV, ACT>), grid, kThreads, 0, st, \
                            G, ld_g, A, ld_a, B, ld_b, rowptr, dst, oth, n_rows, g.fv, gA, ld_ga))
  const bool tiled = g.vec && g.lpr >= 4 && G && B && n_rows >= kTiledMinRows && aligned16(rowptr) && aligned16(dst) &&
                     aligned16(oth) && g.fv <= g.lpr * g.vpl && tiled_enabled();
#define LAUNCH_TILED(VPLV)                                                                                       \
  CWN_DISPATCH_ACT(act, {                                                                                        \
    auto kern = csr_cob_bwd_tiled_kernel<float4, LPR, VPLV, ACT>;                                                \
    const TiledLaunch tl = tiled_launch(kern, n_rows, LPR, F, 2);                                                \
    kern<<<tl.grid, kThreads, tl.dynamic, st>>>(G, ld_g, A, ld_a, B, ld_b, rowptr, dst, oth, n_rows, g.fv, gA,   \
                                                ld_ga, tl.cap, tl.cap);                                          \
  })
  const char mode = large_mode("CWN_B200_LARGE_COB", g.vpl == 1 ? 'p' : 'r');
  const bool large_ok = g.vec && g.lpr >= 4 && G && B && n_rows >= kTiledMinRows && g.fv <= g.lpr * g.vpl;
#define LAUNCH_PIPE(VPLV)                                                                                          \
  CWN_DISPATCH_ACT(act, {                                                                                          \
    using Pass = CobBwdPass<float4, LPR, VPLV, ACT>;                                                               \
    Pass::Params prm{G, A, B, dst, oth, gA, (uint32_t)ld_g * 4u, (uint32_t)ld_a * 4u, (uint32_t)ld_b * 4u,         \
                     (uint32_t)ld_ga * 4u, g.fv};                                                                  \
    launch_pipelined<Pass>(prm, rowptr, n_rows, st);                                                               \
  })
  if (large_ok && mode == 'p') {
    if (g.vpl == 2) { constexpr int LPR = 32; LAUNCH_PIPE(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; LAUNCH_PIPE(1) } break;
      case 8: { constexpr int LPR = 8; LAUNCH_PIPE(1) } break;
      case 16: { constexpr int LPR = 16; LAUNCH_PIPE(1) } break;
      default: { constexpr int LPR = 32; LAUNCH_PIPE(1) } break;
    }
  } else if (tiled && mode == 't') {
    if (g.vpl == 2) { constexpr int LPR = 32; LAUNCH_TILED(2) }
    else switch (g.lpr) {
      case 4: { constexpr int LPR = 4; LAUNCH_TILED(1) } break;
      case 8: { constexpr int LPR = 8; LAUNCH_TILED(1) } break;
      case 16: { constexpr int LPR = 16; LAUNCH_TILED(1) } break;
      default: { constexpr int LPR = 32; LAUNCH_TILED(1) } break;
    }
  } else if (g.vec) {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float4, 1)) } else { constexpr int LPR = 32; LAUNCH(float4, 2) }
  } else {
    if (g.vpl == 1) { CWN_DISPATCH_LPR(g.lpr, LAUNCH(float, 1)) } else { constexpr int LPR = 32; LAUNCH(float, 2) }
  }
#undef LAUNCH_PIPE
#undef LAUNCH_TILED
#undef LAUNCH
  return launched("cwn_csr_cob_bwd_f32");
}

extern "C" int cwn_check_index_range(const int64_t* idx, int64_t E, int64_t n, int32_t* flags,
                                     cwn_stream_t stream) {
  if (E < 0) return fail(CWN_E_SHAPE, "cwn_check_index_range: bad E");
  if (E == 0) return CWN_OK;
  if (!idx || !flags) return fail(CWN_E_NULL, "idx/flags");
  int64_t need = (E + 255) / 256;
  if (need > kNumSMs * 8) need = kNumSMs * 8;
  check_index_range_kernel<<<(int)need, 256, 0, (cudaStream_t)stream>>>(idx, E, n, flags);
  return launched("cwn_check_index_range");
}
