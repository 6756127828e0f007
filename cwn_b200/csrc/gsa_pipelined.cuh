// Software-pipelined row-per-group CSR passes (the HBM-bound regime).
//
// After the instruction diet of the row kernels (gsa.cu) the limiter at 1M rows is the DEPENDENT chain of a row:
// rowptr -> message indices -> gathers -> (indices of the second trip -> gathers) -> store, ~4 exposed memory
// latencies for ~640 compulsory bytes with 32 resident warps per SM. Here the plan reads run ahead of the feature
// gathers, in registers:
//   * the row pointers of the group's NEXT row are requested when the current row starts;
//   * while the gathers of trip t are in flight, the payload indices of "whatever comes next" are requested — trip
//     t+1 of this row, or the first trip of the next row — and only then are the gathers consumed.
// The exposed chain of a row shrinks to its gather trips. Accumulation order is untouched (plan order inside a row,
// one lane group per row): results stay bit-identical to the sequential definition.
//
// A pass is described by a policy struct (what a message gathers, how it is folded into the accumulator, how the
// row is finished); the driver below owns the pipeline.
#pragma once
#include "common.cuh"

namespace cwn {

#ifndef CWN_PIPE_MIN_BLOCKS
#define CWN_PIPE_MIN_BLOCKS 1  // A/B switch: resident CTAs per SM the register allocation must allow
#endif

template <class Pass>
__global__ void __launch_bounds__(kThreads, CWN_PIPE_MIN_BLOCKS)
csr_pipelined_kernel(typename Pass::Params prm, const int32_t* __restrict__ rowptr, int64_t n_rows) {
  constexpr int U = Pass::U, NI = Pass::NIDX, LPR = Pass::LPR, RPB = kThreads / LPR;
  const int lane = threadIdx.x % LPR, sub = threadIdx.x / LPR;
  Pass pass;
  if (!pass.init(prm, lane)) return;  // lane beyond the row (no warp-level primitive below)
  const int64_t stride = (int64_t)gridDim.x * RPB;
  int64_t r = (int64_t)blockIdx.x * RPB + sub;
  if (r >= n_rows) return;
  int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
  int cur[U][NI];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (beg + u < end) pass.load_idx(prm, beg + u, cur[u]);
  while (true) {
    const int64_t rn = r + stride;
    const bool has_next = rn < n_rows;
    int nbeg = 0, nend = 0;
    if (has_next) {
      nbeg = __ldg(rowptr + rn);
      nend = __ldg(rowptr + rn + 1);
    }
    pass.begin_row(prm, (uint32_t)r, beg < end);
    if (beg < end) {
      for (int i = beg;;) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + u < end) pass.gather(prm, u, cur[u]);
        // the plan entries of what comes next: the following trip of this row, else the first trip of the next row
        const int ni = i + U;
        const bool more = ni < end;
        const int m = more ? ni : nbeg, lim = more ? end : nend;
        int nxt[U][NI];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (m + u < lim) pass.load_idx(prm, m + u, nxt[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + u < end) pass.consume(u, i + u == beg);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int a = 0; a < NI; ++a) cur[u][a] = nxt[u][a];
        if (!more) break;
        i = ni;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (nbeg + u < nend) pass.load_idx(prm, nbeg + u, cur[u]);
    }
    pass.end_row(prm, (uint32_t)r, beg, end);
    if (!has_next) break;
    r = rn;
    beg = nbeg;
    end = nend;
  }
}

// ------------------------------------------------------------------------------------------------ policies
template <typename V, int LPR_, int VPL, int REDUCE>
struct GatherPass {  // out[r] = (1+eps) * x_res[r] + REDUCE_i x_src[idx[i]]
  static constexpr int U = (VPL == 1) ? 4 : 2, NIDX = 1, LPR = LPR_;
  static constexpr int kStep = LPR * (int)sizeof(V);
  using O = VecOps<V>;
  struct Params {
    const float* x_src; const int32_t* idx; const float* x_res; const float* eps; float* out;
    uint32_t pitch_s, pitch_r, pitch_o; int FV;
  };
  const char *xs, *xr; char* xo;
  float scale; bool live1;
  V acc[VPL], v[U][VPL];
  __device__ __forceinline__ bool init(const Params& p, int lane) {
    if (lane >= p.FV) return false;
    live1 = VPL == 2 && lane + LPR < p.FV;
    xs = opaque(reinterpret_cast<const char*>(p.x_src) + (size_t)lane * sizeof(V));
    xr = opaque(reinterpret_cast<const char*>(p.x_res) + (size_t)lane * sizeof(V));
    xo = reinterpret_cast<char*>(p.out) + (size_t)lane * sizeof(V);
    scale = p.x_res ? __fadd_rn(1.f, p.eps ? __ldg(p.eps) : 0.f) : 0.f;
    return true;
  }
  __device__ __forceinline__ void load_idx(const Params& p, int m, int (&o)[1]) const { o[0] = __ldg(p.idx + m); }
  __device__ __forceinline__ void begin_row(const Params&, uint32_t, bool) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
  }
  __device__ __forceinline__ void gather(const Params& p, int u, const int (&ix)[1]) {
    const char* a = row_at(xs, (uint32_t)ix[0], p.pitch_s);
    v[u][0] = O::load(a);
    if (VPL == 2 && live1) v[u][VPL - 1] = O::load(a + kStep);
  }
  __device__ __forceinline__ void consume(int u, bool first) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (REDUCE == CWN_REDUCE_MAX) acc[k] = first ? v[u][k] : O::map2(acc[k], v[u][k], MaxOp());
      else acc[k] = O::add(acc[k], v[u][k]);
    }
  }
  __device__ __forceinline__ void end_row(const Params& p, uint32_t r, int beg, int end) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (k == 1 && !live1) break;
      V a = acc[k];
      if (REDUCE == CWN_REDUCE_MEAN) {
        const float cnt = (float)max(end - beg, 1);
        a = O::map2(a, a, [cnt](float x, float) { return __fdiv_rn(x, cnt); });
      }
      if (p.x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, r, p.pitch_r) + k * kStep));
      O::store(row_at(xo, r, p.pitch_o) + k * kStep, a);
    }
  }
};

template <typename V, int LPR_, int VPL, int ACT>
struct CobFwdPass {  // out[r] = (1+eps) * x_res[r] + SUM_i act(P[src[i]] + Q[cob[i]])
#ifdef CWN_COB_U
  static constexpr int U = CWN_COB_U;
#else
  static constexpr int U = 2;  // measured: 2 beats 3 and 4 (registers -> resident warps)
#endif
  static constexpr int NIDX = 2, LPR = LPR_;
  static constexpr int kStep = LPR * (int)sizeof(V);
  using O = VecOps<V>;
  struct Params {
    const float *P, *Q; const int32_t *src, *cob; const float* x_res; const float* eps; float* out;
    uint32_t pitch_p, pitch_q, pitch_r, pitch_o; int FV;
  };
  const char *pl, *ql, *xr; char* xo;
  float scale; bool live1;
  V acc[VPL], vp[U][VPL], vq[U][VPL];
  __device__ __forceinline__ bool init(const Params& p, int lane) {
    if (lane >= p.FV) return false;
    live1 = VPL == 2 && lane + LPR < p.FV;
    pl = opaque(reinterpret_cast<const char*>(p.P) + (size_t)lane * sizeof(V));
    ql = opaque(reinterpret_cast<const char*>(p.Q) + (size_t)lane * sizeof(V));
    xr = opaque(reinterpret_cast<const char*>(p.x_res) + (size_t)lane * sizeof(V));
    xo = reinterpret_cast<char*>(p.out) + (size_t)lane * sizeof(V);
    scale = p.x_res ? __fadd_rn(1.f, p.eps ? __ldg(p.eps) : 0.f) : 0.f;
    return true;
  }
  __device__ __forceinline__ void load_idx(const Params& p, int m, int (&o)[2]) const {
    o[0] = __ldg(p.src + m);
    o[1] = __ldg(p.cob + m);
  }
  __device__ __forceinline__ void begin_row(const Params&, uint32_t, bool) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = O::zero();
  }
  __device__ __forceinline__ void gather(const Params& p, int u, const int (&ix)[2]) {
    const char* a = row_at(pl, (uint32_t)ix[0], p.pitch_p);
    const char* b = row_at(ql, (uint32_t)ix[1], p.pitch_q);
    vp[u][0] = O::load(a);
    vq[u][0] = O::load(b);
    if (VPL == 2 && live1) {
      vp[u][VPL - 1] = O::load(a + kStep);
      vq[u][VPL - 1] = O::load(b + kStep);
    }
  }
  __device__ __forceinline__ void consume(int u, bool) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      V pre = O::add(vp[u][k], vq[u][k]);
      pre = O::map2(pre, pre, [](float x, float) { return act_fwd<ACT>(x); });
      acc[k] = O::add(acc[k], pre);
    }
  }
  __device__ __forceinline__ void end_row(const Params& p, uint32_t r, int, int) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (k == 1 && !live1) break;
      V a = acc[k];
      if (p.x_res) a = O::add_scaled(a, scale, O::load(row_at(xr, r, p.pitch_r) + k * kStep));
      O::store(row_at(xo, r, p.pitch_o) + k * kStep, a);
    }
  }
};

template <typename V, int LPR_, int VPL, int ACT>
struct CobBwdPass {  // gA[r] = SUM_i G[dst[i]] * act'(A[r] + B[oth[i]])
#ifdef CWN_COB_U
  static constexpr int U = CWN_COB_U;
#else
  static constexpr int U = 2;  // measured: 2 beats 3 and 4 (registers -> resident warps)
#endif
  static constexpr int NIDX = 2, LPR = LPR_;
  static constexpr int kStep = LPR * (int)sizeof(V);
  using O = VecOps<V>;
  struct Params {
    const float *G, *A, *B; const int32_t *dst, *oth; float* gA;
    uint32_t pitch_g, pitch_a, pitch_b, pitch_o; int FV;
  };
  const char *gl, *al, *bl; char* xo;
  bool live1;
  V acc[VPL], a[VPL], vg[U][VPL], vb[U][VPL];
  __device__ __forceinline__ bool init(const Params& p, int lane) {
    if (lane >= p.FV) return false;
    live1 = VPL == 2 && lane + LPR < p.FV;
    gl = opaque(reinterpret_cast<const char*>(p.G) + (size_t)lane * sizeof(V));
    al = opaque(reinterpret_cast<const char*>(p.A) + (size_t)lane * sizeof(V));
    bl = opaque(reinterpret_cast<const char*>(p.B) + (size_t)lane * sizeof(V));
    xo = reinterpret_cast<char*>(p.gA) + (size_t)lane * sizeof(V);
    return true;
  }
  __device__ __forceinline__ void load_idx(const Params& p, int m, int (&o)[2]) const {
    o[0] = __ldg(p.dst + m);
    o[1] = __ldg(p.oth + m);
  }
  __device__ __forceinline__ void begin_row(const Params& p, uint32_t r, bool has_messages) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      acc[k] = O::zero();
      a[k] = (has_messages && (k == 0 || live1)) ? O::load(row_at(al, r, p.pitch_a) + k * kStep) : O::zero();
    }
  }
  __device__ __forceinline__ void gather(const Params& p, int u, const int (&ix)[2]) {
    const char* g = row_at(gl, (uint32_t)ix[0], p.pitch_g);
    const char* b = row_at(bl, (uint32_t)ix[1], p.pitch_b);
    vg[u][0] = O::load(g);
    vb[u][0] = O::load(b);
    if (VPL == 2 && live1) {
      vg[u][VPL - 1] = O::load(g + kStep);
      vb[u][VPL - 1] = O::load(b + kStep);
    }
  }
  __device__ __forceinline__ void consume(int u, bool) {
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
  __device__ __forceinline__ void end_row(const Params& p, uint32_t r, int, int) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (k == 1 && !live1) break;
      O::store(row_at(xo, r, p.pitch_o) + k * kStep, acc[k]);
    }
  }
};

}  // namespace cwn
