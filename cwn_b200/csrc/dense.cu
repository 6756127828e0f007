// Dense update / combine nets of SparseCINConv as grouped, fused fp32 kernels (forward and backward).
//
// Reference: mp/layers.py:191-199 (forward of SparseCINCochainConv after propagate) with the default nets of
// SparseCINConv, :303-325:  Linear -> BatchNorm -> act -> Linear -> BatchNorm -> act (two branches), then
// Linear(2H -> H) -> BatchNorm -> act over their concatenation. In the reference (PyTorch) every one of those is its
// own library kernel plus autograd bookkeeping: ~40 launches per cochain per layer, ~1 000 per training step at
// 128 molecules, where the step is purely launch/latency bound (profiles/r1_ncu_launches_eager_step_v1.summary.txt).
//
// Here a "unit" z = f_in(X) W^T + b is one 64x64-tiled FFMA GEMM whose input tile applies the PREVIOUS unit's
// BatchNorm + activation on load (normalised activations never touch HBM), whose input may be the virtual
// concatenation [X0 | X1] (no torch.cat), and whose epilogue emits per-tile BatchNorm partials (mean, M2; merged with
// Chan's formula, so no E[x^2]-E[x]^2 cancellation). Every entry point takes up to CWN_MAX_GROUP problems (the two
// branches of the three cochain dimensions) in ONE launch. Backward = column reductions for BatchNorm, then one
// kernel that forms g_z, the input gradient g_z W and per-CTA partials of the weight gradient g_z^T f_in(X); a final
// ordered sum makes the weight gradients deterministic (no atomics anywhere).
//
// Tensor cores are deliberately not used: TF32 (10-bit mantissa) breaks the 1e-5 rtol parity gate at K = 64..128
// (SURVEY 7), and at the real-data shape these GEMMs are ~25 MFLOP each — launch latency, not flops, is the cost.
#include "common.cuh"

namespace cwn {

constexpr int TM = 64;    // rows per tile
constexpr int TN = 64;    // columns per tile
constexpr int DT = 256;   // threads per CTA (16 x 16, 4x4 outputs each)
constexpr int LDT = TN + 4;

template <class D>
struct Group {
  D d[CWN_MAX_GROUP];
  int start[CWN_MAX_GROUP + 1];  // first CTA of each problem
  int n;
};

template <class D>
__device__ __forceinline__ int find_problem(const Group<D>& g, int cta) {
  int p = 0;
#pragma unroll
  for (int i = 1; i < CWN_MAX_GROUP; ++i)
    if (i < g.n && cta >= g.start[i]) p = i;
  return p;
}

__device__ __forceinline__ float act_apply(int act, float v) {
  switch (act) {
    case CWN_ACT_RELU: return fmaxf(v, 0.f);
    case CWN_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case CWN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case CWN_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__device__ __forceinline__ float act_grad(int act, float v) {  // derivative as a function of the pre-activation
  switch (act) {
    case CWN_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case CWN_ACT_ELU: return v > 0.f ? 1.f : expf(v);
    case CWN_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
    case CWN_ACT_TANH: { float t = tanhf(v); return 1.f - t * t; }
    default: return 1.f;
  }
}

__host__ __device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// acc[4][4] += A[64 x inner] * B[inner x 64] with A row-major in shared memory (rows = output rows, float4 along the
// inner dimension) and B inner-major (float4 along the output columns). Thread (ty, tx) owns rows 4ty.., cols 4tx..
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                         int inner4, int ty, int tx, float (&acc)[4][4]) {
  const float* a0 = As + (ty * 4) * lda;
  const float* b0 = Bs + tx * 4;
  for (int k = 0; k < inner4; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + k);
#pragma unroll
    for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(b0 + (k + q) * ldb);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[i][0] = fmaf(av[q], b[q].x, acc[i][0]);
        acc[i][1] = fmaf(av[q], b[q].y, acc[i][1]);
        acc[i][2] = fmaf(av[q], b[q].z, acc[i][2]);
        acc[i][3] = fmaf(av[q], b[q].w, acc[i][3]);
      }
    }
  }
}

// f_in(X)[row][k] for the (possibly concatenated, possibly BatchNorm+activation-transformed) unit input
template <class D>
__device__ __forceinline__ float unit_input(const D& d, int64_t row, int k) {
  float v;
  if (k < d.k0) {
    v = __ldg(d.x0 + row * d.ld_x0 + k);
    if (d.in_scale0) v = (v - __ldg(d.in_mean0 + k)) * __ldg(d.in_scale0 + k) + __ldg(d.in_beta0 + k);
  } else {
    const int k1 = k - d.k0;
    v = __ldg(d.x1 + row * d.ld_x1 + k1);
    if (d.in_scale1) v = (v - __ldg(d.in_mean1 + k1)) * __ldg(d.in_scale1 + k1) + __ldg(d.in_beta1 + k1);
  }
  return act_apply(d.in_act, v);
}

// ------------------------------------------------------------------------------------------------ forward unit
__global__ void __launch_bounds__(DT) linear_fwd_kernel(const __grid_constant__ Group<cwn_linear_desc> g) {
  extern __shared__ __align__(16) float smem[];
  const int p = find_problem(g, blockIdx.x);
  const cwn_linear_desc& d = g.d[p];
  const int t = blockIdx.x - g.start[p];
  const int col_tiles = (d.h + TN - 1) / TN;
  const int rt = t / col_tiles, ct = t % col_tiles;
  const int K = d.k0 + d.k1, K4 = round4(K), lda = K4 + 4;
  float* As = smem;             // [TM][lda]
  float* Bs = smem + TM * lda;  // [K4][LDT]   Bs[k][n] = W[col0 + n][k]
  const int64_t row0 = (int64_t)rt * TM;
  const int col0 = ct * TN;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;

  for (int i = tid; i < TN * K4; i += DT) {
    const int n = i / K4, k = i % K4;
    float w = 0.f;
    if (k < K && col0 + n < d.h) w = __ldg(d.w + (int64_t)(col0 + n) * d.ld_w + k);
    Bs[k * LDT + n] = w;
  }
  for (int i = tid; i < TM * K4; i += DT) {
    const int r = i / K4, k = i % K4;
    float v = 0.f;
    if (k < K && row0 + r < d.n_rows) v = unit_input(d, row0 + r, k);
    As[r * lda + k] = v;
  }
  __syncthreads();
  float acc[4][4] = {};
  tile_mma(As, lda, Bs, LDT, K4, ty, tx, acc);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = col0 + tx * 4 + j;
    const float b = (d.bias && c < d.h) ? __ldg(d.bias + c) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][j] += b;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= d.n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < d.h) d.z[row * d.ld_z + c] = acc[i][j];
    }
  }
  if (!d.stats) return;
  __syncthreads();  // As/Bs are dead: reuse the front of shared memory for the output tile
  float* Ys = smem;  // [TM][LDT]
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(Ys + (ty * 4 + i) * LDT + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  if (tid < TN && col0 + tid < d.h) {
    const int cnt = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
    float s = 0.f;
    for (int r = 0; r < cnt; ++r) s += Ys[r * LDT + tid];
    const float mean = s / (float)cnt;
    float m2 = 0.f;
    for (int r = 0; r < cnt; ++r) {
      const float dv = Ys[r * LDT + tid] - mean;
      m2 = fmaf(dv, dv, m2);
    }
    d.stats[((int64_t)rt * 2 + 0) * d.h + col0 + tid] = mean;
    d.stats[((int64_t)rt * 2 + 1) * d.h + col0 + tid] = m2;
  }
}

// ------------------------------------------------------------------------------------------------ BN statistics
__global__ void __launch_bounds__(DT) bn_finalize_kernel(const __grid_constant__ Group<cwn_bn_desc> g) {
  const cwn_bn_desc& d = g.d[blockIdx.x];
  for (int c = threadIdx.x; c < d.h; c += DT) {
    float mean, rstd;
    if (d.training) {
      float n = 0.f, m2 = 0.f;
      mean = 0.f;
      for (int t = 0; t < d.n_tiles; ++t) {  // Chan's parallel-variance merge, tiles in order (deterministic)
        const int64_t left = d.n_rows - (int64_t)t * TM;
        const float cnt = (float)(left < TM ? left : TM);
        const float mt = d.stats[((int64_t)t * 2 + 0) * d.h + c];
        const float m2t = d.stats[((int64_t)t * 2 + 1) * d.h + c];
        const float delta = mt - mean, tot = n + cnt;
        mean += delta * (cnt / tot);
        m2 += m2t + delta * delta * (n * cnt / tot);
        n = tot;
      }
      const float var = m2 / (float)d.n_rows;
      rstd = 1.f / sqrtf(var + d.eps);
      if (d.running_mean) d.running_mean[c] = (1.f - d.momentum) * d.running_mean[c] + d.momentum * mean;
      if (d.running_var) {
        const float unbiased = d.n_rows > 1 ? m2 / (float)(d.n_rows - 1) : var;
        d.running_var[c] = (1.f - d.momentum) * d.running_var[c] + d.momentum * unbiased;
      }
    } else {
      mean = d.running_mean[c];
      rstd = 1.f / sqrtf(d.running_var[c] + d.eps);
    }
    d.mean[c] = mean;
    d.rstd[c] = rstd;
    d.scale[c] = d.gamma ? d.gamma[c] * rstd : rstd;
  }
  if (threadIdx.x == 0 && d.training && d.num_batches_tracked) *d.num_batches_tracked += 1;
}

__global__ void __launch_bounds__(DT) bn_act_kernel(const __grid_constant__ Group<cwn_bn_act_desc> g) {
  const int p = find_problem(g, blockIdx.x);
  const cwn_bn_act_desc& d = g.d[p];
  const int64_t row0 = (int64_t)(blockIdx.x - g.start[p]) * TM;
  const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
  for (int i = threadIdx.x; i < rows * d.h; i += DT) {
    const int r = i / d.h, c = i % d.h;
    float v = d.z[(row0 + r) * d.ld_z + c];
    if (d.scale) v = (v - d.mean[c]) * d.scale[c] + (d.beta ? d.beta[c] : 0.f);
    d.out[(row0 + r) * d.ld_out + c] = act_apply(d.act, v);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// g_out * act'(y) and zhat for one element of a unit's output
__device__ __forceinline__ void unit_gy(const cwn_unit_bwd_desc& d, int64_t row, int c, float& gy, float& zhat) {
  const float z = __ldg(d.z + row * d.ld_z + c);
  const float g = __ldg(d.g_out + row * d.ld_g + c);
  if (d.has_bn) {
    const float zc = z - __ldg(d.mean + c);
    const float y = zc * __ldg(d.scale + c) + (d.beta ? __ldg(d.beta + c) : 0.f);
    zhat = zc * __ldg(d.rstd + c);
    gy = g * act_grad(d.act, y);
  } else {
    zhat = 0.f;
    gy = g * act_grad(d.act, z);
  }
}

__global__ void __launch_bounds__(DT) unit_bwd_reduce_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  __shared__ float part[2][4][TN];
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc& d = g.d[p];
  const int tile = blockIdx.x - g.start[p];
  const int64_t row0 = (int64_t)tile * TM;
  const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
  const int lane_c = threadIdx.x & (TN - 1), rg = threadIdx.x >> 6;  // 4 row groups x 64 columns
  for (int col0 = 0; col0 < d.h; col0 += TN) {
    const int c = col0 + lane_c;
    float s1 = 0.f, s2 = 0.f;
    if (c < d.h)
      for (int r = rg; r < rows; r += 4) {
        float gy, zhat;
        unit_gy(d, row0 + r, c, gy, zhat);
        s1 += gy;
        s2 = fmaf(gy, zhat, s2);
      }
    part[0][rg][lane_c] = s1;
    part[1][rg][lane_c] = s2;
    __syncthreads();
    if (rg == 0 && c < d.h) {
      d.red_partials[((int64_t)tile * 2 + 0) * d.h + c] = ((part[0][0][lane_c] + part[0][1][lane_c]) + part[0][2][lane_c]) + part[0][3][lane_c];
      d.red_partials[((int64_t)tile * 2 + 1) * d.h + c] = ((part[1][0][lane_c] + part[1][1][lane_c]) + part[1][2][lane_c]) + part[1][3][lane_c];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(DT) unit_bwd_finalize_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  const cwn_unit_bwd_desc& d = g.d[blockIdx.x];
  if (!d.has_bn) return;
  const int n_tiles = (int)((d.n_rows + TM - 1) / TM);
  for (int c = threadIdx.x; c < d.h; c += DT) {
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      s1 += d.red_partials[((int64_t)t * 2 + 0) * d.h + c];
      s2 += d.red_partials[((int64_t)t * 2 + 1) * d.h + c];
    }
    d.c1[c] = s1 / (float)d.n_rows;
    d.c2[c] = s2 / (float)d.n_rows;
    if (d.g_gamma) d.g_gamma[c] = d.accumulate_affine ? d.g_gamma[c] + s2 : s2;
    if (d.g_beta) d.g_beta[c] = d.accumulate_affine ? d.g_beta[c] + s1 : s1;
  }
}

// g_z tile -> input gradient (g_z W) and per-CTA partial weight gradient (g_z^T f_in(X)); CTA j of a problem strides
// over the row tiles j, j + n_ctas, ... and owns slab j of the partial buffers (plain read-modify-write, no atomics).
__global__ void __launch_bounds__(DT) unit_bwd_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  extern __shared__ __align__(16) float smem[];
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc& d = g.d[p];
  const int j = blockIdx.x - g.start[p];
  const int K = d.k0 + d.k1, H4 = round4(d.h), ldg = H4 + 4;
  const int m_tiles = (d.h + TM - 1) / TM;
  float* Gz = smem;                        // [TM][ldg]            g_z[r][c]
  float* GzT = Gz + TM * ldg;              // [m_tiles*TM][LDT]    g_z^T[c][r]
  float* Ain = GzT + m_tiles * TM * LDT;   // [TM][LDT]            f_in(X)[r][k chunk]
  float* Ws = Ain + TM * LDT;              // [H4][LDT]            W[c][k chunk]
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int n_tiles = (int)((d.n_rows + TM - 1) / TM);
  float* wpart = d.w_partials + (int64_t)j * d.h * K;
  float* bpart = d.b_partials + (int64_t)j * d.h;
  bool first = true;
  for (int tile = j; tile < n_tiles; tile += d.n_ctas, first = false) {
    const int64_t row0 = (int64_t)tile * TM;
    const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
    __syncthreads();
    for (int i = tid; i < TM * H4; i += DT) {  // g_z of this tile
      const int r = i / H4, c = i % H4;
      float gz = 0.f;
      if (r < rows && c < d.h) {
        float gy, zhat;
        unit_gy(d, row0 + r, c, gy, zhat);
        gz = d.has_bn ? __ldg(d.scale + c) * (gy - d.c1[c] - zhat * d.c2[c]) : gy;
      }
      Gz[r * ldg + c] = gz;
    }
    for (int i = tid; i < m_tiles * TM * TM; i += DT) GzT[(i / TM) * LDT + (i % TM)] = 0.f;
    __syncthreads();
    for (int i = tid; i < TM * d.h; i += DT) {
      const int r = i / d.h, c = i % d.h;
      GzT[c * LDT + r] = Gz[r * ldg + c];
    }
    if (tid < d.h || (d.h > DT)) {
      for (int c = tid; c < d.h; c += DT) {  // bias gradient partial: column sums of g_z
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += Gz[r * ldg + c];
        bpart[c] = first ? s : bpart[c] + s;
      }
    }
    for (int kc = 0; kc < K; kc += TN) {
      __syncthreads();
      for (int i = tid; i < H4 * TN; i += DT) {  // W[c][kc + k]
        const int c = i / TN, k = i % TN;
        Ws[c * LDT + k] = (c < d.h && kc + k < K) ? __ldg(d.w + (int64_t)c * d.ld_w + kc + k) : 0.f;
      }
      for (int i = tid; i < TM * TN; i += DT) {  // f_in(X)[r][kc + k]
        const int r = i / TN, k = i % TN;
        Ain[r * LDT + k] = (r < rows && kc + k < K) ? unit_input(d, row0 + r, kc + k) : 0.f;
      }
      __syncthreads();
      if (d.g_in0 || d.g_in1) {  // input gradient chunk: [64 rows] x [64 k] = Gz [64 x h] * Ws [h x 64]
        float acc[4][4] = {};
        tile_mma(Gz, ldg, Ws, LDT, H4, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = ty * 4 + i;
          if (r >= rows) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = kc + tx * 4 + q;
            if (k >= K) continue;
            if (k < d.k0) { if (d.g_in0) d.g_in0[(row0 + r) * d.ld_gi0 + k] = acc[i][q]; }
            else if (d.g_in1) d.g_in1[(row0 + r) * d.ld_gi1 + (k - d.k0)] = acc[i][q];
          }
        }
      }
      for (int mt = 0; mt < m_tiles; ++mt) {  // weight gradient chunk: [64 c] x [64 k] = GzT [64 x rows] * Ain [rows x 64]
        float acc[4][4] = {};
        tile_mma(GzT + mt * TM * LDT, LDT, Ain, LDT, TM, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = mt * TM + ty * 4 + i;
          if (c >= d.h) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = kc + tx * 4 + q;
            if (k >= K) continue;
            float* dst = wpart + (int64_t)c * K + k;
            *dst = first ? acc[i][q] : *dst + acc[i][q];
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(DT) wgrad_finalize_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc& d = g.d[p];
  const int K = d.k0 + d.k1;
  const int64_t i = (int64_t)(blockIdx.x - g.start[p]) * DT + threadIdx.x;
  const int64_t total = (int64_t)d.h * K;
  if (i < total) {
    float s = 0.f;
    for (int j = 0; j < d.n_ctas; ++j) s += d.w_partials[(int64_t)j * total + i];
    float* dst = d.g_w + (i / K) * d.ld_gw + (i % K);
    *dst = d.accumulate_w ? *dst + s : s;
  } else if (i < total + d.h && d.g_b) {
    const int c = (int)(i - total);
    float s = 0.f;
    for (int j = 0; j < d.n_ctas; ++j) s += d.b_partials[(int64_t)j * d.h + c];
    d.g_b[c] = d.accumulate_w ? d.g_b[c] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <class Kernel>
static int ensure_smem(Kernel kernel, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return CWN_OK;
  if (bytes > 227 * 1024) return fail(CWN_E_SHAPE, "dense kernels: K or h too large for shared memory");
  return cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what);
}

static int check_group(const void* descs, int n, const char* what) {
  if (n < 0 || n > CWN_MAX_GROUP) return fail(CWN_E_SHAPE, what);
  if (n > 0 && !descs) return fail(CWN_E_NULL, what);
  return CWN_OK;
}

}  // namespace cwn

using namespace cwn;

extern "C" int cwn_linear_fwd_grouped(const cwn_linear_desc* descs, int32_t n, cwn_stream_t stream) {
  int rc = check_group(descs, n, "cwn_linear_fwd_grouped");
  if (rc || n == 0) return rc;
  Group<cwn_linear_desc> g;
  g.n = n;
  int total = 0;
  size_t smem = (size_t)TM * LDT * sizeof(float);
  for (int i = 0; i < n; ++i) {
    const cwn_linear_desc& d = descs[i];
    if (d.n_rows < 0 || d.h <= 0 || d.k0 <= 0 || d.k1 < 0 || d.n_rows > (int64_t)INT32_MAX * TM)
      return fail(CWN_E_SHAPE, "cwn_linear_fwd_grouped: bad shape");
    if (d.n_rows > 0 && (!d.x0 || !d.w || !d.z || (d.k1 > 0 && !d.x1))) return fail(CWN_E_NULL, "cwn_linear_fwd_grouped: operand");
    g.d[i] = d;
    g.start[i] = total;
    total += (int)((d.n_rows + TM - 1) / TM) * ((d.h + TN - 1) / TN);
    const int K4 = round4(d.k0 + d.k1);
    const size_t need = ((size_t)TM * (K4 + 4) + (size_t)K4 * LDT) * sizeof(float);
    if (need > smem) smem = need;
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  if ((rc = ensure_smem(linear_fwd_kernel, smem, "cudaFuncSetAttribute(linear_fwd_kernel)"))) return rc;
  linear_fwd_kernel<<<total, DT, smem, (cudaStream_t)stream>>>(g);
  return launched("linear_fwd_kernel");
}

extern "C" int cwn_bn_finalize_grouped(const cwn_bn_desc* descs, int32_t n, cwn_stream_t stream) {
  int rc = check_group(descs, n, "cwn_bn_finalize_grouped");
  if (rc || n == 0) return rc;
  Group<cwn_bn_desc> g;
  g.n = n;
  for (int i = 0; i < n; ++i) {
    const cwn_bn_desc& d = descs[i];
    if (d.h <= 0 || d.n_rows < 0) return fail(CWN_E_SHAPE, "cwn_bn_finalize_grouped: bad shape");
    if (!d.mean || !d.scale || !d.rstd) return fail(CWN_E_NULL, "cwn_bn_finalize_grouped: outputs");
    if (d.training ? !d.stats : (!d.running_mean || !d.running_var)) return fail(CWN_E_NULL, "cwn_bn_finalize_grouped: statistics");
    g.d[i] = d;
    g.start[i] = i;
  }
  bn_finalize_kernel<<<n, DT, 0, (cudaStream_t)stream>>>(g);
  return launched("bn_finalize_kernel");
}

extern "C" int cwn_bn_act_grouped(const cwn_bn_act_desc* descs, int32_t n, cwn_stream_t stream) {
  int rc = check_group(descs, n, "cwn_bn_act_grouped");
  if (rc || n == 0) return rc;
  Group<cwn_bn_act_desc> g;
  g.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_bn_act_desc& d = descs[i];
    if (d.h <= 0 || d.n_rows < 0) return fail(CWN_E_SHAPE, "cwn_bn_act_grouped: bad shape");
    if (d.n_rows > 0 && (!d.z || !d.out)) return fail(CWN_E_NULL, "cwn_bn_act_grouped: operand");
    g.d[i] = d;
    g.start[i] = total;
    total += (int)((d.n_rows + TM - 1) / TM);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  bn_act_kernel<<<total, DT, 0, (cudaStream_t)stream>>>(g);
  return launched("bn_act_kernel");
}

static int load_bwd_group(const cwn_unit_bwd_desc* descs, int n, Group<cwn_unit_bwd_desc>& g, const char* what) {
  int rc = check_group(descs, n, what);
  if (rc) return rc;
  g.n = n;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = descs[i];
    if (d.n_rows < 0 || d.h <= 0 || d.k0 <= 0 || d.k1 < 0 || d.n_ctas < 0) return fail(CWN_E_SHAPE, what);
    if (d.n_rows > 0 && (!d.x0 || !d.w || !d.z || !d.g_out || (d.k1 > 0 && !d.x1))) return fail(CWN_E_NULL, what);
    if (d.has_bn && (!d.mean || !d.scale || !d.rstd || !d.red_partials || !d.c1 || !d.c2)) return fail(CWN_E_NULL, what);
    g.d[i] = d;
  }
  return CWN_OK;
}

extern "C" int cwn_unit_bwd_reduce_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_unit_bwd_reduce_grouped");
  if (rc || n == 0) return rc;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    g.start[i] = total;
    if (g.d[i].has_bn) total += (int)((g.d[i].n_rows + TM - 1) / TM);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  unit_bwd_reduce_kernel<<<total, DT, 0, (cudaStream_t)stream>>>(g);
  return launched("unit_bwd_reduce_kernel");
}

extern "C" int cwn_unit_bwd_finalize_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_unit_bwd_finalize_grouped");
  if (rc || n == 0) return rc;
  for (int i = 0; i <= n; ++i) g.start[i] = i;
  unit_bwd_finalize_kernel<<<n, DT, 0, (cudaStream_t)stream>>>(g);
  return launched("unit_bwd_finalize_kernel");
}

extern "C" int cwn_unit_bwd_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_unit_bwd_grouped");
  if (rc || n == 0) return rc;
  int total = 0;
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    const int n_tiles = (int)((d.n_rows + TM - 1) / TM);
    if (d.n_ctas > n_tiles || (n_tiles > 0 && d.n_ctas == 0)) return fail(CWN_E_SHAPE, "cwn_unit_bwd_grouped: n_ctas must be in [1, row tiles]");
    if (d.n_rows > 0 && (!d.w_partials || !d.b_partials)) return fail(CWN_E_NULL, "cwn_unit_bwd_grouped: partial buffers");
    g.start[i] = total;
    total += d.n_ctas;
    const int H4 = round4(d.h), m_tiles = (d.h + TM - 1) / TM;
    const size_t need = ((size_t)TM * (H4 + 4) + (size_t)m_tiles * TM * LDT + (size_t)TM * LDT + (size_t)H4 * LDT) * sizeof(float);
    if (need > smem) smem = need;
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  if ((rc = ensure_smem(unit_bwd_kernel, smem, "cudaFuncSetAttribute(unit_bwd_kernel)"))) return rc;
  unit_bwd_kernel<<<total, DT, smem, (cudaStream_t)stream>>>(g);
  return launched("unit_bwd_kernel");
}

extern "C" int cwn_wgrad_finalize_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_wgrad_finalize_grouped");
  if (rc || n == 0) return rc;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    if (!d.g_w) return fail(CWN_E_NULL, "cwn_wgrad_finalize_grouped: g_w");
    g.start[i] = total;
    if (d.n_ctas > 0) total += (int)(((int64_t)d.h * (d.k0 + d.k1) + d.h + DT - 1) / DT);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  wgrad_finalize_kernel<<<total, DT, 0, (cudaStream_t)stream>>>(g);
  return launched("wgrad_finalize_kernel");
}
