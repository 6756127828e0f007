// Readout head of the SparseCIN-family models in one launch forward and two backward
// (reference mp/nn.py:50-60 pool_complex, mp/models.py:230-254, mp/molec_models.py:137-161):
//     pooled_d[b] = SUM | MEAN_{i : batch_d[i] = b} x_d[i]                     (per cochain dimension d)
//     z_d[b]      = pooled_d[b] W1_d^T + b1_d ;   h[b] = SUM | MEAN_d act(z_d[b])
//     out[b]      = h[b] W2^T + b2
// Through torch this tail is ~45 tiny launches per training step (3 pooled reductions, stack, 3 addmm, 3 relu, stack,
// sum, addmm and their backward), all on the critical path between the last layer and the first backward kernel.
// Here one CTA owns one complex: its cells are a contiguous run of the row plan of `batch`, the pooled vectors and
// the hidden vector live in shared memory, and the weights (H2 x K per dimension, ~100 KB in all) stream from L2.
// Parameter gradients are sums over the complexes: a second backward kernel gives every gradient element to one
// thread that walks the complexes in order — deterministic, no atomics — and can accumulate straight into `.grad`.
#include "common.cuh"

namespace cwn {

constexpr int kHeadThreads = 256;

// the per-dimension descriptors travel by value as a kernel argument (no device allocation; capturable in a CUDA graph)
struct HeadDims {
  cwn_head_dim d[CWN_MAX_HEAD_DIMS];
};

template <int ACT>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const __grid_constant__ HeadDims dims, int n_dims, int K, int H2, int out_size, int pool_mean,
                int final_mean, const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ h_out,
                float* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  extern __shared__ float sm[];  // pooled[n_dims][K] | h[H2]
  float* s_pooled = sm;
  float* s_h = sm + n_dims * K;
  const int b = blockIdx.x;
  // ---- pooling: thread (d, k) walks the cells of complex b in plan order (the order of segment_pool's kernel)
  for (int e = threadIdx.x; e < n_dims * K; e += kHeadThreads) {
    const int d = e / K, k = e - d * K;
    const cwn_head_dim& D = dims.d[d];
    float acc = 0.f;
    if (D.rowptr) {
      const int beg = __ldg(D.rowptr + b), end = __ldg(D.rowptr + b + 1);
      for (int i = beg; i < end; ++i) {
        const int64_t row = D.perm ? (int64_t)__ldg(D.perm + i) : (int64_t)i;
        acc = __fadd_rn(acc, __ldg(D.x + row * D.ld_x + k));
      }
      if (pool_mean) acc = __fdiv_rn(acc, (float)max(end - beg, 1));
    }
    s_pooled[e] = acc;
    D.pooled[(int64_t)b * K + k] = acc;
  }
  __syncthreads();
  // ---- per-dimension Linear + activation, combined over the dimensions
  for (int j = threadIdx.x; j < H2; j += kHeadThreads) {
    float hsum = 0.f;
    for (int d = 0; d < n_dims; ++d) {
      const cwn_head_dim& D = dims.d[d];
      const float* w = D.w1 + (int64_t)j * K;
      const float* p = s_pooled + d * K;
      float z = 0.f;
      if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15u) == 0) {
        for (int k = 0; k < K; k += 4) {
          const float4 wv = ldg_f4(w + k);
          z = fmaf(p[k], wv.x, z);
          z = fmaf(p[k + 1], wv.y, z);
          z = fmaf(p[k + 2], wv.z, z);
          z = fmaf(p[k + 3], wv.w, z);
        }
      } else {
        for (int k = 0; k < K; ++k) z = fmaf(p[k], __ldg(w + k), z);
      }
      if (D.b1) z += __ldg(D.b1 + j);
      D.z[(int64_t)b * H2 + j] = z;
      hsum += act_fwd<ACT>(z);
    }
    if (final_mean) hsum = hsum / (float)n_dims;
    s_h[j] = hsum;
    h_out[(int64_t)b * H2 + j] = hsum;
  }
  __syncthreads();
  // ---- output Linear: one warp per output column
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < out_size; o += kHeadThreads / 32) {
    float acc = 0.f;
    for (int j = lane; j < H2; j += 32) acc = fmaf(s_h[j], __ldg(w2 + (int64_t)o * H2 + j), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[(int64_t)b * out_size + o] = acc + (b2 ? __ldg(b2 + o) : 0.f);
  }
}

// g_z_d[b] = (g_out[b] W2) * act'(z_d[b]) [/ n_dims] ;  g_pooled_d[b] = g_z_d[b] W1_d ;  g_x_d[i] = g_pooled_d[batch[i]] [/ count]
template <int ACT>
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_input_kernel(const __grid_constant__ HeadDims dims, int n_dims, int K, int H2, int out_size, int pool_mean,
                      int final_mean, const float* __restrict__ w2, const float* __restrict__ g_out) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  extern __shared__ float sm[];  // g_z[n_dims][H2]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < H2; j += kHeadThreads) {
    float gh = 0.f;
    for (int o = 0; o < out_size; ++o) gh = fmaf(__ldg(g_out + (int64_t)b * out_size + o), __ldg(w2 + (int64_t)o * H2 + j), gh);
    if (final_mean) gh = gh / (float)n_dims;
    for (int d = 0; d < n_dims; ++d) {
      const cwn_head_dim& D = dims.d[d];
      const float gz = gh * act_bwd<ACT>(D.z[(int64_t)b * H2 + j]);
      sm[d * H2 + j] = gz;
      D.g_z[(int64_t)b * H2 + j] = gz;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n_dims * K; e += kHeadThreads) {
    const int d = e / K, k = e - d * K;
    const cwn_head_dim& D = dims.d[d];
    if (!D.g_x || !D.rowptr) continue;
    const float* gz = sm + d * H2;
    float acc = 0.f;
    for (int j = 0; j < H2; ++j) acc = fmaf(gz[j], __ldg(D.w1 + (int64_t)j * K + k), acc);  // coalesced over k
    const int beg = __ldg(D.rowptr + b), end = __ldg(D.rowptr + b + 1);
    if (pool_mean) acc = acc / (float)max(end - beg, 1);
    for (int i = beg; i < end; ++i) {
      const int64_t row = D.perm ? (int64_t)__ldg(D.perm + i) : (int64_t)i;
      D.g_x[row * D.ld_gx + k] = acc;
    }
  }
}

// parameter gradients: element e of [ g_W1_0 | g_b1_0 | g_W1_1 | ... | g_W2 | g_b2 ] is a sum over the B complexes.
// A CTA owns 32 consecutive elements (lane -> element: the `pooled` / `h` reads of a warp are coalesced, the g_z / g_out
// reads are warp broadcasts) and its kHeadThreads / 32 warps each sum one contiguous range of complexes; the ranges are
// then added in order by warp 0 — a fixed two-level order, so the result does not depend on the launch. (One thread per
// element walking all 128 complexes took 41 us, the longest kernel of the step.)
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_param_kernel(const __grid_constant__ HeadDims dims, int n_dims, int64_t B, int K, int H2, int out_size,
                      const float* __restrict__ h, const float* __restrict__ g_out, float* __restrict__ g_w2,
                      float* __restrict__ g_b2, int accumulate_out) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  constexpr int kWarps = kHeadThreads / 32;
  __shared__ float part[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t per_dim = (int64_t)H2 * K + H2;
  const int64_t total = per_dim * n_dims + (int64_t)out_size * H2 + out_size;
  const int64_t chunk = (B + kWarps - 1) / kWarps;
  const int64_t b0 = warp * chunk, b1 = (b0 + chunk < B) ? b0 + chunk : B;
  for (int64_t e0 = (int64_t)blockIdx.x * 32; e0 < total; e0 += (int64_t)gridDim.x * 32) {
    const int64_t e = e0 + lane;
    float acc = 0.f;
    float* dst = nullptr;
    bool accumulate = false;
    if (e < per_dim * n_dims) {
      const int d = (int)(e / per_dim);
      const int64_t r = e - d * per_dim;
      const cwn_head_dim& D = dims.d[d];
      accumulate = D.accumulate != 0;
      if (r < (int64_t)H2 * K) {
        const int j = (int)(r / K), k = (int)(r - (int64_t)j * K);
        dst = D.g_w1 ? D.g_w1 + r : nullptr;
        if (dst)
          for (int64_t b = b0; b < b1; ++b) acc = fmaf(D.g_z[b * H2 + j], D.pooled[b * K + k], acc);
      } else {
        const int j = (int)(r - (int64_t)H2 * K);
        dst = D.g_b1 ? D.g_b1 + j : nullptr;
        if (dst)
          for (int64_t b = b0; b < b1; ++b) acc += D.g_z[b * H2 + j];
      }
    } else if (e < total) {
      const int64_t r = e - per_dim * n_dims;
      accumulate = accumulate_out != 0;
      if (r < (int64_t)out_size * H2) {
        const int o = (int)(r / H2), j = (int)(r - (int64_t)o * H2);
        dst = g_w2 ? g_w2 + r : nullptr;
        if (dst)
          for (int64_t b = b0; b < b1; ++b) acc = fmaf(__ldg(g_out + b * out_size + o), __ldg(h + b * H2 + j), acc);
      } else {
        const int o = (int)(r - (int64_t)out_size * H2);
        dst = g_b2 ? g_b2 + o : nullptr;
        if (dst)
          for (int64_t b = b0; b < b1; ++b) acc += __ldg(g_out + b * out_size + o);
      }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && dst) {
      float s = part[0][lane];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) s += part[w][lane];
      *dst = (accumulate ? *dst : 0.f) + s;
    }
    __syncthreads();
  }
}

static int check_head(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2, int32_t out_size,
                      int32_t act, bool forward) {
  if (!dims) return fail(CWN_E_NULL, "cwn_readout_head: dims");
  if (n_dims < 1 || n_dims > CWN_MAX_HEAD_DIMS) return fail(CWN_E_SHAPE, "cwn_readout_head: 1..CWN_MAX_HEAD_DIMS dimensions");
  if (B < 0 || B > INT32_MAX || K <= 0 || H2 <= 0 || out_size <= 0) return fail(CWN_E_SHAPE, "cwn_readout_head: bad sizes");
  if ((size_t)(n_dims * (size_t)K + H2) * 4 > 200 * 1024 || (size_t)n_dims * H2 * 4 > 200 * 1024)
    return fail(CWN_E_SHAPE, "cwn_readout_head: K / H2 too large for the shared-memory vectors");
  if (act < CWN_ACT_ID || act > CWN_ACT_TANH) return fail(CWN_E_ENUM, "unknown activation");
  for (int d = 0; d < n_dims; ++d) {
    if (!dims[d].w1 || !dims[d].pooled || !dims[d].z) return fail(CWN_E_NULL, "cwn_readout_head: w1 / pooled / z");
    if (forward && dims[d].rowptr && !dims[d].x) return fail(CWN_E_NULL, "cwn_readout_head: x");
    if (forward && dims[d].rowptr && dims[d].ld_x < K) return fail(CWN_E_SHAPE, "cwn_readout_head: ld_x < K");
  }
  return CWN_OK;
}

}  // namespace cwn

using namespace cwn;

#define CWN_HEAD_BY_ACT(ACTV, ...)                                                     \
  switch (ACTV) {                                                                      \
    case CWN_ACT_ID: { constexpr int ACT = CWN_ACT_ID; __VA_ARGS__; } break;           \
    case CWN_ACT_RELU: { constexpr int ACT = CWN_ACT_RELU; __VA_ARGS__; } break;       \
    case CWN_ACT_ELU: { constexpr int ACT = CWN_ACT_ELU; __VA_ARGS__; } break;         \
    case CWN_ACT_SIGMOID: { constexpr int ACT = CWN_ACT_SIGMOID; __VA_ARGS__; } break; \
    default: { constexpr int ACT = CWN_ACT_TANH; __VA_ARGS__; } break;                 \
  }

extern "C" int cwn_readout_head_fwd(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2,
                                    int32_t out_size, int32_t act, int32_t pool_mean, int32_t final_mean,
                                    const float* w2, const float* b2, float* h, float* out, cwn_stream_t stream) {
  int rc;
  if ((rc = check_head(dims, n_dims, B, K, H2, out_size, act, true))) return rc;
  if (!w2 || !h || !out) return fail(CWN_E_NULL, "cwn_readout_head_fwd: w2 / h / out");
  if (B == 0) return CWN_OK;
  HeadDims hd;
  for (int d = 0; d < n_dims; ++d) hd.d[d] = dims[d];
  const size_t smem = ((size_t)n_dims * K + H2) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  CWN_HEAD_BY_ACT(act, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_fwd_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(head_fwd_kernel<ACT>, (int)B, kHeadThreads, smem, st, hd, n_dims, K, H2, out_size, pool_mean, final_mean, w2, b2, h, out);
  })
  return launched("cwn_readout_head_fwd");
}

extern "C" int cwn_readout_head_bwd_parts(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2,
                                          int32_t out_size, int32_t act, int32_t pool_mean, int32_t final_mean,
                                          const float* w2, const float* h, const float* g_out, float* g_w2, float* g_b2,
                                          int32_t accumulate_out, int32_t parts, cwn_stream_t stream);

extern "C" int cwn_readout_head_bwd(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2,
                                    int32_t out_size, int32_t act, int32_t pool_mean, int32_t final_mean,
                                    const float* w2, const float* h, const float* g_out, float* g_w2, float* g_b2,
                                    int32_t accumulate_out, cwn_stream_t stream) {
  return cwn_readout_head_bwd_parts(dims, n_dims, B, K, H2, out_size, act, pool_mean, final_mean, w2, h, g_out, g_w2, g_b2,
                                    accumulate_out, 3, stream);
}

// parts: bit 0 = input gradients (g_z, g_x: what the rest of the backward pass waits for), bit 1 = parameter gradients
// (ordered sums over the complexes, ~40 us at B = 128: nothing reads them before the optimizer, so a caller may issue
// them on another stream once the first part has been launched)
extern "C" int cwn_readout_head_bwd_parts(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2,
                                          int32_t out_size, int32_t act, int32_t pool_mean, int32_t final_mean,
                                          const float* w2, const float* h, const float* g_out, float* g_w2, float* g_b2,
                                          int32_t accumulate_out, int32_t parts, cwn_stream_t stream) {
  int rc;
  if (!(parts & 3)) return fail(CWN_E_ENUM, "cwn_readout_head_bwd_parts: parts must have bit 0 and / or bit 1 set");
  if ((rc = check_head(dims, n_dims, B, K, H2, out_size, act, false))) return rc;
  if (!w2 || !h || !g_out) return fail(CWN_E_NULL, "cwn_readout_head_bwd: w2 / h / g_out");
  for (int d = 0; d < n_dims; ++d) {
    if (!dims[d].g_z) return fail(CWN_E_NULL, "cwn_readout_head_bwd: g_z");
    if (dims[d].g_x && dims[d].ld_gx < K) return fail(CWN_E_SHAPE, "cwn_readout_head_bwd: ld_gx < K");
  }
  if (B == 0) return CWN_OK;
  HeadDims hd;
  for (int d = 0; d < n_dims; ++d) hd.d[d] = dims[d];
  const size_t smem = (size_t)n_dims * H2 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (parts & 1) {
    CWN_HEAD_BY_ACT(act, {
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(head_bwd_input_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(head_bwd_input_kernel<ACT>, (int)B, kHeadThreads, smem, st, hd, n_dims, K, H2, out_size, pool_mean, final_mean, w2, g_out);
    })
  }
  if (parts & 2) {
    const int64_t total = ((int64_t)H2 * K + H2) * n_dims + (int64_t)out_size * H2 + out_size;
    int64_t grid = (total + 31) / 32;  // 32 elements per CTA (see head_bwd_param_kernel)
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    launch_pdl(head_bwd_param_kernel, (int)grid, kHeadThreads, 0, st, hd, n_dims, B, K, H2, out_size, h, g_out, g_w2, g_b2, accumulate_out);
  }
  return launched("cwn_readout_head_bwd", (unsigned)((parts & 1) + ((parts >> 1) & 1)));
}
