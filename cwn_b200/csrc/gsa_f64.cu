// fp64 instantiations of the CSR passes: the reference switches the default dtype to float64 for
// the strongly-regular-graph isomorphism experiments (exp/run_exp.py:41-43), where 1e-? differences between embeddings
// decide the metric. Same contracts as the fp32 entry points of gsa.cu (include/cwn_b200.h), same accumulation order
// (messages of a row in plan order, one thread per output element, no atomics) => bit-identical to a sequential CPU
// scatter_add_ in float64. Those datasets are a few thousand small graphs: the kernels are written for clarity (one
// thread per (row, column), coalesced along the feature dimension), not tuned like the fp32 family.
#include "common.cuh"

namespace cwn {
namespace f64 {

constexpr int kThreads = 256;

__device__ __forceinline__ double act_fwd_d(int act, double v) {
  switch (act) {
    case CWN_ACT_RELU: return v > 0.0 ? v : 0.0;
    case CWN_ACT_ELU: return v > 0.0 ? v : expm1(v);
    case CWN_ACT_SIGMOID: return 1.0 / (1.0 + exp(-v));
    case CWN_ACT_TANH: return tanh(v);
    default: return v;
  }
}
__device__ __forceinline__ double act_bwd_d(int act, double v) {
  switch (act) {
    case CWN_ACT_RELU: return v > 0.0 ? 1.0 : 0.0;
    case CWN_ACT_ELU: return v > 0.0 ? 1.0 : exp(v);
    case CWN_ACT_SIGMOID: { const double s = 1.0 / (1.0 + exp(-v)); return s * (1.0 - s); }
    case CWN_ACT_TANH: { const double t = tanh(v); return 1.0 - t * t; }
    default: return 1.0;
  }
}

// out[r] = (1+eps) x_res[r] + REDUCE_i x_src[idx ? idx[i] : i]
__global__ void __launch_bounds__(kThreads)
gather_reduce_kernel(const double* __restrict__ x_src, int64_t ld_src, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ idx, int64_t n_rows, int F, const double* __restrict__ x_res,
                     int64_t ld_res, const double* __restrict__ eps, double* __restrict__ out, int64_t ld_out, int reduce) {
  const double scale = x_res ? __dadd_rn(1.0, eps ? *eps : 0.0) : 0.0;
  const int64_t total = n_rows * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    const int beg = rowptr[r], end = rowptr[r + 1];
    double acc = 0.0;
    for (int m = beg; m < end; ++m) {
      const double v = x_src[(int64_t)(idx ? idx[m] : m) * ld_src + c];
      if (reduce == CWN_REDUCE_MAX) acc = (m == beg) ? v : fmax(acc, v);
      else acc = __dadd_rn(acc, v);
    }
    if (reduce == CWN_REDUCE_MEAN) acc = __ddiv_rn(acc, (double)max(end - beg, 1));
    if (x_res) acc = __dadd_rn(acc, __dmul_rn(scale, x_res[r * ld_res + c]));  // two roundings, like the reference's ops
    out[r * ld_out + c] = acc;
  }
}

__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const double* __restrict__ x, int64_t ld_x, const int64_t* __restrict__ idx, int64_t E, int F,
                   double scale, double* __restrict__ out, int64_t ld_out) {
  const int64_t total = E * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / F;
    const int c = (int)(i - e * F);
    const double v = x[idx[e] * ld_x + c];
    out[e * ld_out + c] = scale == 1.0 ? v : __dmul_rn(scale, v);
  }
}

// out[r] = (1+eps) x_res[r] + SUM_i act(P[src[i]] + Q[cob[i]])
__global__ void __launch_bounds__(kThreads)
cob_fwd_kernel(const double* __restrict__ P, int64_t ld_p, const double* __restrict__ Q, int64_t ld_q,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const int32_t* __restrict__ cob,
               int64_t n_rows, int F, int act, const double* __restrict__ x_res, int64_t ld_res,
               const double* __restrict__ eps, double* __restrict__ out, int64_t ld_out) {
  const double scale = x_res ? __dadd_rn(1.0, eps ? *eps : 0.0) : 0.0;
  const int64_t total = n_rows * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    double acc = 0.0;
    for (int m = rowptr[r]; m < rowptr[r + 1]; ++m)
      acc = __dadd_rn(acc, act_fwd_d(act, __dadd_rn(P[(int64_t)src[m] * ld_p + c], Q[(int64_t)cob[m] * ld_q + c])));
    if (x_res) acc = __dadd_rn(acc, __dmul_rn(scale, x_res[r * ld_res + c]));
    out[r * ld_out + c] = acc;
  }
}

// gA[r] = SUM_i G[dst[i]] * act'(A[r] + B[oth[i]])
__global__ void __launch_bounds__(kThreads)
cob_bwd_kernel(const double* __restrict__ G, int64_t ld_g, const double* __restrict__ A, int64_t ld_a,
               const double* __restrict__ B, int64_t ld_b, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ dst, const int32_t* __restrict__ oth, int64_t n_rows, int F, int act,
               double* __restrict__ gA, int64_t ld_ga) {
  const int64_t total = n_rows * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    const int beg = rowptr[r], end = rowptr[r + 1];
    const double a = beg < end ? A[r * ld_a + c] : 0.0;
    double acc = 0.0;
    for (int m = beg; m < end; ++m) {
      const double pre = __dadd_rn(a, B[(int64_t)oth[m] * ld_b + c]);
      const double g = G[(int64_t)dst[m] * ld_g + c];
      acc = __dadd_rn(acc, act == CWN_ACT_RELU ? (pre > 0.0 ? g : 0.0) : __dmul_rn(g, act_bwd_d(act, pre)));
    }
    gA[r * ld_ga + c] = acc;
  }
}

static int grid_for(int64_t total) {
  int64_t need = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (need > cap) need = cap;
  return (int)(need < 1 ? 1 : need);
}
static int check_mat(const double* p, int64_t ld, int F, const char* name) {
  if (!p) return fail(CWN_E_NULL, name);
  if (ld < F) return fail(CWN_E_SHAPE, "leading dimension smaller than F");
  if ((reinterpret_cast<uintptr_t>(p) & 7u) != 0) return fail(CWN_E_ALIGN, name);
  return CWN_OK;
}

}  // namespace f64
}  // namespace cwn

using namespace cwn;
using namespace cwn::f64;

extern "C" int cwn_csr_gather_reduce_f64(const double* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                                         int64_t n_rows, int32_t F, const double* x_res, int64_t ld_res,
                                         const double* eps, double* out, int64_t ld_out, int32_t reduce,
                                         cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_gather_reduce_f64: bad n_rows/F");
  if (reduce < CWN_REDUCE_ADD || reduce > CWN_REDUCE_MAX) return fail(CWN_E_ENUM, "unknown reduce");
  if (x_res && reduce != CWN_REDUCE_ADD) return fail(CWN_E_ENUM, "x_res requires CWN_REDUCE_ADD");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr) return fail(CWN_E_NULL, "rowptr");
  int rc;
  if ((rc = check_mat(out, ld_out, F, "out"))) return rc;
  if (x_src && (rc = check_mat(x_src, ld_src, F, "x_src"))) return rc;
  if (x_res && (rc = check_mat(x_res, ld_res, F, "x_res"))) return rc;
  gather_reduce_kernel<<<grid_for(n_rows * F), kThreads, 0, (cudaStream_t)stream>>>(x_src, ld_src, rowptr, idx, n_rows, F, x_res,
                                                                                    ld_res, eps, out, ld_out, reduce);
  return launched("cwn_csr_gather_reduce_f64");
}

extern "C" int cwn_gather_rows_f64(const double* x, int64_t ld_x, const int64_t* idx, int64_t E, int32_t F, double scale,
                                   double* out, int64_t ld_out, cwn_stream_t stream) {
  if (E < 0 || F <= 0) return fail(CWN_E_SHAPE, "cwn_gather_rows_f64: bad E/F");
  if (E == 0) return CWN_OK;
  if (!idx) return fail(CWN_E_NULL, "idx");
  int rc;
  if ((rc = check_mat(x, ld_x, F, "x")) || (rc = check_mat(out, ld_out, F, "out"))) return rc;
  gather_rows_kernel<<<grid_for(E * F), kThreads, 0, (cudaStream_t)stream>>>(x, ld_x, idx, E, F, scale, out, ld_out);
  return launched("cwn_gather_rows_f64");
}

extern "C" int cwn_csr_cob_fwd_f64(const double* P, int64_t ld_p, const double* Q, int64_t ld_q, const int32_t* rowptr,
                                   const int32_t* src, const int32_t* cob, int64_t n_rows, int32_t F, int32_t act,
                                   const double* x_res, int64_t ld_res, const double* eps, double* out, int64_t ld_out,
                                   cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_cob_fwd_f64: bad n_rows/F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr || !src || !cob) return fail(CWN_E_NULL, "cwn_csr_cob_fwd_f64: plan");
  int rc;
  if ((rc = check_mat(P, ld_p, F, "P")) || (rc = check_mat(Q, ld_q, F, "Q")) || (rc = check_mat(out, ld_out, F, "out"))) return rc;
  if (x_res && (rc = check_mat(x_res, ld_res, F, "x_res"))) return rc;
  cob_fwd_kernel<<<grid_for(n_rows * F), kThreads, 0, (cudaStream_t)stream>>>(P, ld_p, Q, ld_q, rowptr, src, cob, n_rows, F, act,
                                                                              x_res, ld_res, eps, out, ld_out);
  return launched("cwn_csr_cob_fwd_f64");
}

extern "C" int cwn_csr_cob_bwd_f64(const double* G, int64_t ld_g, const double* A, int64_t ld_a, const double* B,
                                   int64_t ld_b, const int32_t* rowptr, const int32_t* dst, const int32_t* oth,
                                   int64_t n_rows, int32_t F, int32_t act, double* gA, int64_t ld_ga, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0 || n_rows > INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_cob_bwd_f64: bad n_rows/F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!rowptr || !dst || !oth) return fail(CWN_E_NULL, "cwn_csr_cob_bwd_f64: plan");
  int rc;
  if ((rc = check_mat(G, ld_g, F, "G")) || (rc = check_mat(A, ld_a, F, "A")) || (rc = check_mat(B, ld_b, F, "B")) ||
      (rc = check_mat(gA, ld_ga, F, "gA")))
    return rc;
  cob_bwd_kernel<<<grid_for(n_rows * F), kThreads, 0, (cudaStream_t)stream>>>(G, ld_g, A, ld_a, B, ld_b, rowptr, dst, oth, n_rows,
                                                                              F, act, gA, ld_ga);
  return launched("cwn_csr_cob_bwd_f64");
}
