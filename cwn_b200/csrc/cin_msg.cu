// K5 — the message MLP of the dense CIN layers with BatchNorm over the MESSAGE population
// (reference mp/layers.py:94-103 `msg_nn(cat[x_j, attr])`, nets mp/models.py:40-47: Linear(2F -> F), act, BatchNorm1d(F)).
//
// The Linear is applied in split-weight form, pre_e = P[src_e] + Q[att_e] with P = x W1^T and Q = y W2^T + b per CELL
// (csrc/gsa.cu, cwn_csr_cob_fwd_f32), so a message is a_e = act(pre_e) and its BatchNorm is affine in a_e:
//     out[t] = SUM_{e -> t} (scale (a_e - mean) + beta) = scale * S[t] + deg(t) (beta - scale * mean),   S[t] = SUM_{e -> t} a_e
// S is the existing fused pass; the statistics over the E messages are mean = colsum(S) / E and
// var = colsum(S2) / E with S2[t] = SUM_{e -> t} (a_e - mean)^2 (second pass, centred: no E[a^2] - E[a]^2 cancellation)
// -> cwn_cin_msg_sq_f32. Backward through the statistics (c1 = mean_e G_e, c2 = mean_e G_e ahat_e are row reductions of
// G, S and deg) needs the per-message gradient g_pre_e = scale (G[t_e] - c1 - ahat_e c2) act'(pre_e) summed per source /
// per attribute cell -> cwn_cin_msg_bwd_f32 on the by-source and by-attribute plans. No [E, F] tensor is materialised.
//
// Thread mapping: one thread per (row, 4-feature chunk), 128-bit loads, a sequential walk over the row's messages in plan
// order (deterministic, no atomics). These passes are not on the benchmark's path (SparseCIN is); they are written for
// clarity, not tuned like gsa.cu.
#include "common.cuh"

namespace cwn {

__device__ __forceinline__ float act_f(int act, float v) {
  switch (act) {
    case CWN_ACT_RELU: return fmaxf(v, 0.f);
    case CWN_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case CWN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case CWN_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__device__ __forceinline__ float act_df(int act, float v) {
  switch (act) {
    case CWN_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case CWN_ACT_ELU: return v > 0.f ? 1.f : expf(v);
    case CWN_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
    case CWN_ACT_TANH: { const float t = tanhf(v); return 1.f - t * t; }
    default: return 1.f;
  }
}

template <int W>  // W = 4: float4 chunks (F % 4 == 0, aligned rows); W = 1: scalar
struct Chunk {
  float v[W];
  __device__ __forceinline__ static Chunk load(const float* p) {
    Chunk c;
    if (W == 4) { const float4 t = ldg_f4(p); c.v[0] = t.x; c.v[1] = t.y; c.v[2] = t.z; c.v[3] = t.w; }
    else c.v[0] = __ldg(p);
    return c;
  }
  __device__ __forceinline__ void store(float* p) const {
    if (W == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
  }
};

// out[t] = SUM_{i in row t} (act(P[src[i]] + Q[att[i]]) - mu)^2
template <int W>
__global__ void __launch_bounds__(256) cin_msg_sq_kernel(const float* __restrict__ P, int64_t ld_p, const float* __restrict__ Q,
                                                         int64_t ld_q, const int32_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ src, const int32_t* __restrict__ att,
                                                         int64_t n_rows, int F, int act, const float* __restrict__ mu,
                                                         float* __restrict__ out, int64_t ld_out) {
  const int chunks = F / W;
  const int64_t total = n_rows * chunks;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = e / chunks;
    const int c = (int)(e - t * chunks) * W;
    const Chunk<W> m = Chunk<W>::load(mu + c);
    Chunk<W> acc;
#pragma unroll
    for (int k = 0; k < W; ++k) acc.v[k] = 0.f;
    for (int i = __ldg(rowptr + t), end = __ldg(rowptr + t + 1); i < end; ++i) {
      const Chunk<W> p = Chunk<W>::load(P + (int64_t)__ldg(src + i) * ld_p + c);
      const Chunk<W> q = Chunk<W>::load(Q + (int64_t)__ldg(att + i) * ld_q + c);
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float d = act_f(act, p.v[k] + q.v[k]) - m.v[k];
        acc.v[k] = fmaf(d, d, acc.v[k]);
      }
    }
    acc.store(out + t * ld_out + c);
  }
}

// gA[r] = SUM_{i in row r} scale (G[dst[i]] - c1 - ahat_i c2) act'(pre_i),  pre_i = A[r] + B[oth[i]],
//         ahat_i = (act(pre_i) - mu) rstd
template <int W>
__global__ void __launch_bounds__(256) cin_msg_bwd_kernel(const float* __restrict__ G, int64_t ld_g, const float* __restrict__ A,
                                                          int64_t ld_a, const float* __restrict__ B, int64_t ld_b,
                                                          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dst,
                                                          const int32_t* __restrict__ oth, int64_t n_rows, int F, int act,
                                                          const float* __restrict__ scale, const float* __restrict__ mu,
                                                          const float* __restrict__ rstd, const float* __restrict__ c1,
                                                          const float* __restrict__ c2, float* __restrict__ gA, int64_t ld_ga) {
  const int chunks = F / W;
  const int64_t total = n_rows * chunks;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / chunks;
    const int c = (int)(e - r * chunks) * W;
    const Chunk<W> sc = Chunk<W>::load(scale + c), m = Chunk<W>::load(mu + c), rs = Chunk<W>::load(rstd + c);
    const Chunk<W> k1 = Chunk<W>::load(c1 + c), k2 = Chunk<W>::load(c2 + c);
    const Chunk<W> a = Chunk<W>::load(A + r * ld_a + c);
    Chunk<W> acc;
#pragma unroll
    for (int k = 0; k < W; ++k) acc.v[k] = 0.f;
    for (int i = __ldg(rowptr + r), end = __ldg(rowptr + r + 1); i < end; ++i) {
      const Chunk<W> g = Chunk<W>::load(G + (int64_t)__ldg(dst + i) * ld_g + c);
      const Chunk<W> b = Chunk<W>::load(B + (int64_t)__ldg(oth + i) * ld_b + c);
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float pre = a.v[k] + b.v[k];
        const float ahat = (act_f(act, pre) - m.v[k]) * rs.v[k];
        acc.v[k] += sc.v[k] * (g.v[k] - k1.v[k] - ahat * k2.v[k]) * act_df(act, pre);
      }
    }
    acc.store(gA + r * ld_ga + c);
  }
}

static int grid_of(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cwn

using namespace cwn;

extern "C" int cwn_cin_msg_sq_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q, const int32_t* rowptr,
                                  const int32_t* src, const int32_t* att, int64_t n_rows, int32_t F, int32_t act,
                                  const float* mu, float* out, int64_t ld_out, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0) return fail(CWN_E_SHAPE, "cwn_cin_msg_sq_f32: bad n_rows / F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "cwn_cin_msg_sq_f32: unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!P || !Q || !rowptr || !mu || !out) return fail(CWN_E_NULL, "cwn_cin_msg_sq_f32: operand");
  const bool vec = F % 4 == 0 && aligned16(P) && aligned16(Q) && aligned16(mu) && aligned16(out) && ld_p % 4 == 0 &&
                   ld_q % 4 == 0 && ld_out % 4 == 0;
  if (vec)
    cin_msg_sq_kernel<4><<<grid_of(n_rows * (F / 4)), 256, 0, (cudaStream_t)stream>>>(P, ld_p, Q, ld_q, rowptr, src, att, n_rows, F,
                                                                                    act, mu, out, ld_out);
  else
    cin_msg_sq_kernel<1><<<grid_of(n_rows * F), 256, 0, (cudaStream_t)stream>>>(P, ld_p, Q, ld_q, rowptr, src, att, n_rows, F, act,
                                                                              mu, out, ld_out);
  return launched("cin_msg_sq_kernel");
}

extern "C" int cwn_cin_msg_bwd_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B, int64_t ld_b,
                                   const int32_t* rowptr, const int32_t* dst, const int32_t* oth, int64_t n_rows, int32_t F,
                                   int32_t act, const float* scale, const float* mu, const float* rstd, const float* c1,
                                   const float* c2, float* gA, int64_t ld_ga, cwn_stream_t stream) {
  if (n_rows < 0 || F <= 0) return fail(CWN_E_SHAPE, "cwn_cin_msg_bwd_f32: bad n_rows / F");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "cwn_cin_msg_bwd_f32: unknown activation");
  if (n_rows == 0) return CWN_OK;
  if (!G || !A || !B || !rowptr || !scale || !mu || !rstd || !c1 || !c2 || !gA) return fail(CWN_E_NULL, "cwn_cin_msg_bwd_f32: operand");
  const bool vec = F % 4 == 0 && aligned16(G) && aligned16(A) && aligned16(B) && aligned16(gA) && aligned16(scale) &&
                   aligned16(mu) && aligned16(rstd) && aligned16(c1) && aligned16(c2) && ld_g % 4 == 0 && ld_a % 4 == 0 &&
                   ld_b % 4 == 0 && ld_ga % 4 == 0;
  if (vec)
    cin_msg_bwd_kernel<4><<<grid_of(n_rows * (F / 4)), 256, 0, (cudaStream_t)stream>>>(G, ld_g, A, ld_a, B, ld_b, rowptr, dst, oth,
                                                                                     n_rows, F, act, scale, mu, rstd, c1, c2, gA, ld_ga);
  else
    cin_msg_bwd_kernel<1><<<grid_of(n_rows * F), 256, 0, (cudaStream_t)stream>>>(G, ld_g, A, ld_a, B, ld_b, rowptr, dst, oth, n_rows,
                                                                               F, act, scale, mu, rstd, c1, c2, gA, ld_ga);
  return launched("cin_msg_bwd_kernel");
}
