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
#include <algorithm>

#ifndef CWN_MMA_PIPELINE
#define CWN_MMA_PIPELINE 1  // register double-buffering of the shared-memory operands in tile_mma
#endif
#ifndef CWN_FWD_MIN_CTAS
#define CWN_FWD_MIN_CTAS 3  // __launch_bounds__ occupancy target of linear_fwd_kernel: the 408 CTAs of the real-data
                            // shape fit ONE wave (2 CTAs/SM made it two: 29.6 us vs 14.7 us at 304 CTAs); 80 registers
#endif
#ifndef CWN_BWD_MIN_CTAS
#define CWN_BWD_MIN_CTAS 2  // __launch_bounds__ occupancy target of unit_bwd_kernel (caps registers at 128)
#endif

namespace cwn {

// Profiling aid (variant build -DCWN_PHASE_TIMING only): thread 0 of every CTA stamps clock64() at phase boundaries
// into a caller-provided buffer [CTA][16]; slot 15 holds %globaltimer at CTA start. tools/phase_timing.py reads it.
#ifdef CWN_PHASE_TIMING
__device__ long long* g_phase_buf = nullptr;
#define CWN_PHASE(i)                                                                    \
  do {                                                                                  \
    if (threadIdx.x == 0 && g_phase_buf) {                                              \
      g_phase_buf[(size_t)blockIdx.x * 16 + (i)] = clock64();                           \
      if ((i) == 0) {                                                                   \
        unsigned long long t_;                                                          \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
        g_phase_buf[(size_t)blockIdx.x * 16 + 15] = (long long)t_;                      \
      }                                                                                 \
    }                                                                                   \
  } while (0)
#else
#define CWN_PHASE(i)
#endif

constexpr int TM = 64;    // rows per tile (large problems); small ones use 32-row tiles for twice the CTAs
constexpr int TN = 64;    // columns per tile
constexpr int DT = 256;   // threads per CTA (16 x 16, 4x4 outputs each)
constexpr int LDT = TN + 4;

template <class D>
struct Group {
  D d[CWN_MAX_GROUP];
  int start[CWN_MAX_GROUP + 1];  // first CTA of each problem
  int n;
};

// rows of a problem that are real (the rest pads a fixed-capacity batch, see cwn_linear_desc::n_rows_live)
template <class D>
__device__ __forceinline__ int64_t live_rows(const D& d) {
  if (!d.n_rows_live) return d.n_rows;
  int64_t n = (int64_t)__ldg(d.n_rows_live);
  n = n < 1 ? 1 : n;
  return n < d.n_rows ? n : d.n_rows;
}

template <class D>
__device__ __forceinline__ int find_problem(const Group<D>& g, int cta) {
  int p = 0;
#pragma unroll
  for (int i = 1; i < CWN_MAX_GROUP; ++i)
    if (i < g.n && cta >= g.start[i]) p = i;
  return p;
}

// Activations. The runtime-dispatched versions are deliberately NOT inlined: expm1f / tanhf / expf expand to
// ~100 instructions each, and inlining the 5-way switch at every element of every unrolled tile loader made these
// kernels 70-100 KB of straight-line code that each CTA executes once — ncu's top stall was `no_instruction`
// (instruction-cache misses). The kernels are instantiated for the two common codes (identity, ReLU) at compile time
// and fall back to the out-of-line switch for ELU / sigmoid / tanh.
__device__ __noinline__ float act_apply_rt(int act, float v) {
  switch (act) {
    case CWN_ACT_RELU: return fmaxf(v, 0.f);
    case CWN_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case CWN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case CWN_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__device__ __noinline__ float act_grad_rt(int act, float v) {  // derivative as a function of the pre-activation
  switch (act) {
    case CWN_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case CWN_ACT_ELU: return v > 0.f ? 1.f : expf(v);
    case CWN_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
    case CWN_ACT_TANH: { float t = tanhf(v); return 1.f - t * t; }
    default: return 1.f;
  }
}
constexpr int kActRuntime = -1;
template <int A>
__device__ __forceinline__ float act_apply(int act, float v) {
  if (A == CWN_ACT_ID) return v;
  if (A == CWN_ACT_RELU) return fmaxf(v, 0.f);
  return act_apply_rt(act, v);
}
template <int A>
__device__ __forceinline__ float act_grad(int act, float v) {
  if (A == CWN_ACT_ID) return 1.f;
  if (A == CWN_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  return act_grad_rt(act, v);
}

__host__ __device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// acc[R][4] += A[(16 R) x inner] * B[inner x 64] with A row-major in shared memory (rows = output rows, float4 along
// the inner dimension) and B inner-major (float4 along the output columns). Thread (ty, tx) owns rows R*ty.., cols 4tx..
// The operands of step k+4 are fetched into a second register set before the FMAs of step k are issued: with only a
// couple of warps per scheduler (these launches are small) the shared-memory latency would otherwise sit exposed.
template <int R>
__device__ __forceinline__ void mma_step(const float4 (&a)[R], const float4 (&b)[4], float (&acc)[R][4]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
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

template <int R>
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                         int inner4, int ty, int tx, float (&acc)[R][4]) {
  const float* a0 = As + (ty * R) * lda;
  const float* b0 = Bs + tx * 4;
  float4 a[R], b[4], an[R], bn[4];
  auto fetch = [&](float4 (&aa)[R], float4 (&bb)[4], int k) {
#pragma unroll
    for (int i = 0; i < R; ++i) aa[i] = *reinterpret_cast<const float4*>(a0 + i * lda + k);
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = *reinterpret_cast<const float4*>(b0 + (k + q) * ldb);
  };
  fetch(a, b, 0);
#if CWN_MMA_PIPELINE
  // two k-steps per trip with ping-pong operand registers: the loads of the next step are in flight while the FMAs
  // of the current one issue, and no register copies are needed
  for (int k = 0; k < inner4; k += 8) {
    if (k + 4 < inner4) fetch(an, bn, k + 4);
    mma_step<R>(a, b, acc);
    if (k + 4 < inner4) {
      if (k + 8 < inner4) fetch(a, b, k + 8);
      mma_step<R>(an, bn, acc);
    }
  }
#else
  for (int k = 0; k < inner4; k += 4) {
    if (k > 0) fetch(a, b, k);
    mma_step<R>(a, b, acc);
  }
#endif
}

// f_in(X)[row][k] for the (possibly concatenated, possibly BatchNorm+activation-transformed) unit input.
// `vin` = the per-column (mean, scale, beta) of the input transform staged in shared memory: vin[0..K4) mean,
// vin[K4..2K4) scale, vin[2K4..3K4) beta (identity = 0, 1, 0).
template <class D>
__device__ __forceinline__ void stage_input_vectors(const D& d, float* vin, int K, int K4) {
  for (int k = threadIdx.x; k < K4; k += DT) {
    float m = 0.f, sc = 1.f, b = 0.f;
    if (k < d.k0) {
      if (d.in_scale0) { m = __ldg(d.in_mean0 + k); sc = __ldg(d.in_scale0 + k); b = __ldg(d.in_beta0 + k); }
    } else if (k < K) {
      const int k1 = k - d.k0;
      if (d.in_scale1) { m = __ldg(d.in_mean1 + k1); sc = __ldg(d.in_scale1 + k1); b = __ldg(d.in_beta1 + k1); }
    }
    vin[k] = m; vin[K4 + k] = sc; vin[2 * K4 + k] = b;
  }
}

template <class D>
__device__ __forceinline__ bool input_vec_ok(const D& d) {
  auto ok = [](const float* p, int64_t ld) { return ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) && (ld % 4 == 0); };
  return (d.k0 % 4 == 0) && (d.k1 % 4 == 0) && ok(d.x0, d.ld_x0) && (d.k1 == 0 || ok(d.x1, d.ld_x1));
}

// dst[r * ldd + (k - kc)] = f_in(X)[row0 + r][k] for r < TM, k in [kc, kc + width), zero outside the matrix.
// 128-bit loads when the layout allows; (r, column) advance incrementally so there is one division per thread.
template <int A_IN, class D>
__device__ __forceinline__ void load_input_tile(const D& d, const float* vin, int K, int K4, int64_t row0, int rows,
                                                int kc, int width, float* dst, int ldd, bool vec, int tile_rows) {
  const int act = d.in_act;
  if (vec) {
    // All global loads of a batch are issued before any of them is consumed (ncu showed the transform of each
    // float4 stalling on its own load when loads and shared-memory stores were interleaved).
    const int nc4 = width / 4;  // width is a multiple of 4
    const int total = tile_rows * nc4;
    constexpr int NB = 4;
    for (int base = threadIdx.x; base < total; base += NB * DT) {
      float4 v[NB];
      int rr[NB], kk[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int i = base + j * DT;
        rr[j] = i / nc4;
        kk[j] = (i - rr[j] * nc4) * 4;
        const int k = kc + kk[j];
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && rr[j] < rows && k < K) {
          const float* src = (k < d.k0) ? d.x0 + (row0 + rr[j]) * d.ld_x0 + k : d.x1 + (row0 + rr[j]) * d.ld_x1 + (k - d.k0);
          v[j] = ldg_f4(src);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int i = base + j * DT;
        if (i >= total) continue;
        const int k = kc + kk[j];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr[j] < rows && k < K) {
          o.x = act_apply<A_IN>(act, (v[j].x - vin[k]) * vin[K4 + k] + vin[2 * K4 + k]);
          o.y = act_apply<A_IN>(act, (v[j].y - vin[k + 1]) * vin[K4 + k + 1] + vin[2 * K4 + k + 1]);
          o.z = act_apply<A_IN>(act, (v[j].z - vin[k + 2]) * vin[K4 + k + 2] + vin[2 * K4 + k + 2]);
          o.w = act_apply<A_IN>(act, (v[j].w - vin[k + 3]) * vin[K4 + k + 3] + vin[2 * K4 + k + 3]);
        }
        *reinterpret_cast<float4*>(dst + rr[j] * ldd + kk[j]) = o;
      }
    }
  } else {
    for (int i = threadIdx.x; i < tile_rows * width; i += DT) {
      const int r = i / width, kk = i % width, k = kc + kk;
      float v = 0.f;
      if (r < rows && k < K) {
        v = (k < d.k0) ? __ldg(d.x0 + (row0 + r) * d.ld_x0 + k) : __ldg(d.x1 + (row0 + r) * d.ld_x1 + (k - d.k0));
        v = act_apply<A_IN>(act, (v - vin[k]) * vin[K4 + k] + vin[2 * K4 + k]);
      }
      dst[r * ldd + kk] = v;
    }
  }
}

// The descriptor of this CTA's problem, copied once from the kernel-parameter (constant) space to shared memory:
// the Group<> parameter block is 2-3 KB, each CTA needs ~300 B of it, and touching its fields one by one through
// dynamically indexed LDC exposed one constant-cache miss after the other at the top of the kernel.
// The copy is then read BY VALUE into registers: a reference into shared memory made the compiler re-load fields
// (k0, x0, ld_x0 ...) after every barrier and inside the tile loaders.
template <class D>
__device__ __forceinline__ D stage_desc(const Group<D>& g, int p, D* sd) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&g.d[p]);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sd);
  for (int i = threadIdx.x; i < (int)(sizeof(D) / 4); i += DT) dst[i] = src[i];
  __syncthreads();
  return *sd;
}

// "last CTA done" hand-off: every CTA of a problem publishes its partials, bumps the problem's counter, and the one
// that observes the final count runs the (cheap, ordered => deterministic) merge. Saves a dependent launch.
// Release/acquire through ONE thread: the CTA barrier orders every thread's partial stores before thread 0's
// device-scope fence + counter increment (causality is cumulative, the pattern cooperative-groups grid.sync uses);
// 256 fences per CTA were measured at ~2 000 cycles.
__device__ __forceinline__ bool last_cta_of_problem(int32_t* counter, int total) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int last = (atomicAdd(counter, 1) == total - 1);
    if (last) __threadfence();
    s_last = last;
  }
  __syncthreads();
  return s_last != 0;
}

constexpr int kStageFloats = 8192;  // 32 KB of per-tile partials staged in shared memory at a time (generic path)

// Ordered merge of per-tile records [n_tiles][2][h] by ONE CTA (the last one of a problem). This is the serial tail
// of the launch and it is COLD code (executed once, by one CTA, on an SM that has never fetched it), so it is built
// to be both short in round trips and short in instructions: thread (q, g) owns float4 q of the record and the tiles
// g, g + G, ... (G = DT / (h/2)), eight 128-bit loads in flight per trip of a rolled loop, ONE pass.
//   v1 staged the records through shared memory with scalar loads: ~26 dependent L2 round trips, 17 800 cycles
//      (half of the whole forward kernel at the real-data shape);
//   v2 kept 16 tiles per thread in registers, fully unrolled, two passes: one round trip but ~2 000 instructions of
//      straight-line cold code: 10 400 cycles, dominated by instruction fetch.
constexpr int kMergeBatch = 8;  // 128-bit loads in flight per thread per trip

struct MergeMap {
  int P, G, q, g;
  bool live;
  __device__ __forceinline__ MergeMap(int h) {
    P = h / 2;
    G = DT / P;
    q = threadIdx.x % P;
    g = threadIdx.x / P;
    live = g < G;
  }
};

__device__ __forceinline__ bool merge_fast_ok(const float* partials, int h) {
  return (h % 4 == 0) && (h / 2 <= DT) && ((reinterpret_cast<uintptr_t>(partials) & 15u) == 0);
}

// column totals of comb[G][ld]: thread c sums its column over the G groups in order
__device__ __forceinline__ float merge_column(const float* comb, int G, int ld, int c) {
  float t = 0.f;
  for (int g = 0; g < G; ++g) t += comb[g * ld + c];
  return t;
}

// Exact group-combination of the per-tile (mean, M2) in one pass about a pivot p (the mean of tile 0, i.e. a value
// within a fraction of a standard deviation of the global mean, so nothing cancels):
//   mean = SUM cnt_t mean_t / N,   M2 = SUM M2_t + [SUM cnt_t (mean_t - p)^2 - N (mean - p)^2].
// Returns (mean, M2) of column threadIdx.x (< h). comb: 6 DT floats of shared memory.
__device__ __forceinline__ void bn_merge_fast(const float* stats, int n_tiles, int64_t n_rows, int h, int tile_rows,
                                              float* comb, float& mean, float& m2) {
  const MergeMap m(h);
  const bool is_mean = m.q < h / 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;  // a: SUM cnt*mean_t | SUM M2_t;  b: SUM cnt*(mean_t - p)^2
  if (m.live) {
    const float4* rec = reinterpret_cast<const float4*>(stats) + m.q;
    const int h2 = h / 2;  // float4 per record
    const float4 pv = __ldcg(rec);
    CWN_PHASE(9);
    for (int t0 = m.g; t0 < n_tiles; t0 += kMergeBatch * m.G) {
      float4 v[kMergeBatch];
#pragma unroll
      for (int j = 0; j < kMergeBatch; ++j) {
        const int t = t0 + j * m.G;
        v[j] = t < n_tiles ? __ldcg(rec + (int64_t)t * h2) : pv;
      }
#pragma unroll
      for (int j = 0; j < kMergeBatch; ++j) {
        const int t = t0 + j * m.G;
        const int64_t left = n_rows - (int64_t)t * tile_rows;
        float w = t < n_tiles ? (float)(left < tile_rows ? left : tile_rows) : 0.f;
        if (!is_mean) w = t < n_tiles ? 1.f : 0.f;
        a.x = fmaf(w, v[j].x, a.x); a.y = fmaf(w, v[j].y, a.y); a.z = fmaf(w, v[j].z, a.z); a.w = fmaf(w, v[j].w, a.w);
        const float dx = v[j].x - pv.x, dy = v[j].y - pv.y, dz = v[j].z - pv.z, dw = v[j].w - pv.w;
        b.x = fmaf(w * dx, dx, b.x); b.y = fmaf(w * dy, dy, b.y); b.z = fmaf(w * dz, dz, b.z); b.w = fmaf(w * dw, dw, b.w);
      }
    }
    reinterpret_cast<float4*>(comb)[m.g * m.P + m.q] = a;                    // [G][2h]
    if (is_mean) reinterpret_cast<float4*>(comb + 4 * DT)[m.g * (h / 4) + m.q] = b;  // [G][h]
    CWN_PHASE(10);
  }
  __syncthreads();
  if (threadIdx.x < h) {
    const int c = threadIdx.x;
    const float n = (float)n_rows;
    mean = merge_column(comb, m.G, 2 * h, c) / n;
    const float within = merge_column(comb, m.G, 2 * h, h + c);
    const float about_p = merge_column(comb + 4 * DT, m.G, h, c);
    const float dm = mean - __ldcg(stats + c);
    m2 = within + fmaxf(about_p - n * dm * dm, 0.f);
  }
  CWN_PHASE(11);
}

// Plain ordered column sums of [n_tiles][2][h] records: (s1, s2) of column threadIdx.x (< h). comb: 4 DT floats.
__device__ __forceinline__ void sums_merge_fast(const float* partials, int n_tiles, int h, float* comb, float& s1,
                                                float& s2) {
  const MergeMap m(h);
  if (m.live) {
    const float4* rec = reinterpret_cast<const float4*>(partials) + m.q;
    const int h2 = h / 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t0 = m.g; t0 < n_tiles; t0 += kMergeBatch * m.G) {
      float4 v[kMergeBatch];
#pragma unroll
      for (int j = 0; j < kMergeBatch; ++j) {
        const int t = t0 + j * m.G;
        v[j] = t < n_tiles ? __ldcg(rec + (int64_t)t * h2) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < kMergeBatch; ++j) acc = f4_add(acc, v[j]);
    }
    reinterpret_cast<float4*>(comb)[m.g * m.P + m.q] = acc;
  }
  __syncthreads();
  if (threadIdx.x < h) {
    s1 = merge_column(comb, m.G, 2 * h, threadIdx.x);
    s2 = merge_column(comb, m.G, 2 * h, h + threadIdx.x);
  }
}

// BatchNorm statistics of one problem from the per-tile (mean, M2) partials (training) or the running statistics.
// `stage`: kStageFloats floats of shared memory. Column = threadIdx.x (h <= DT).
__device__ __forceinline__ void bn_finalize_body(const float* stats, int n_tiles, int64_t n_rows, int h, const float* gamma,
                                                 float eps, float momentum, int training, float* running_mean,
                                                 float* running_var, int64_t* nbt, float* mean_out, float* scale_out,
                                                 float* rstd_out, float* stage, int stage_floats = kStageFloats,
                                                 int tile_rows = TM) {
  if (training && merge_fast_ok(stats, h)) {
    __shared__ __align__(16) float comb[6 * DT];
    float mean = 0.f, m2 = 0.f;
    // requested before the merge, consumed after it (one round trip less on the serial tail)
    const bool col = threadIdx.x < h;
    const float rm0 = (col && running_mean) ? running_mean[threadIdx.x] : 0.f;
    const float rv0 = (col && running_var) ? running_var[threadIdx.x] : 0.f;
    const float gm = (col && gamma) ? gamma[threadIdx.x] : 1.f;
    __syncthreads();
    bn_merge_fast(stats, n_tiles, n_rows, h, tile_rows, comb, mean, m2);
    if (col) {
      const int c = threadIdx.x;
      const float var = m2 / (float)n_rows;
      const float rstd = 1.f / sqrtf(var + eps);
      if (running_mean) running_mean[c] = (1.f - momentum) * rm0 + momentum * mean;
      if (running_var) {
        const float unbiased = n_rows > 1 ? m2 / (float)(n_rows - 1) : var;
        running_var[c] = (1.f - momentum) * rv0 + momentum * unbiased;
      }
      mean_out[c] = mean;
      rstd_out[c] = rstd;
      scale_out[c] = gm * rstd;
    }
    if (threadIdx.x == 0 && nbt) *nbt += 1;
  } else if (training) {
    // Exact group-combination of the per-tile (mean, M2):  mean = SUM cnt_t mean_t / N,
    //   M2 = SUM [M2_t + cnt_t (mean_t - mean)^2]  — two passes of plain ordered sums (deterministic). A sequential
    // Chan merge was measured first: its dependent chain of two IEEE divisions per tile, run by ONE CTA after all the
    // others had finished, cost more than the GEMM itself. The partials are staged in shared memory with coalesced
    // loads; a column is summed by DT/h threads over interleaved tiles and the partial sums are combined in order.
    const int per_tile = 2 * h;
    const int chunk = stage_floats / per_tile > 0 ? stage_floats / per_tile : 1;
    const int parts = (h <= DT / 4) ? 4 : (h <= DT / 2 ? 2 : 1);  // threads per column
    const int c = threadIdx.x % h, part = threadIdx.x / h;
    const bool live = threadIdx.x < parts * h;
    __shared__ float comb[DT];
    float mean = 0.f, m2 = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      float acc = 0.f;
      for (int t0 = 0; t0 < n_tiles; t0 += chunk) {
        const int nt = (n_tiles - t0 < chunk) ? n_tiles - t0 : chunk;
        __syncthreads();
        for (int i = threadIdx.x; i < nt * per_tile; i += DT) stage[i] = __ldcg(stats + (int64_t)t0 * per_tile + i);
        __syncthreads();
        if (live)
          for (int t = part; t < nt; t += parts) {
            const int64_t left = n_rows - (int64_t)(t0 + t) * tile_rows;
            const float cnt = (float)(left < tile_rows ? left : tile_rows);
            const float mt = stage[t * per_tile + c];
            if (pass == 0) {
              acc = fmaf(cnt, mt, acc);
            } else {
              const float dm = mt - mean;
              acc += stage[t * per_tile + h + c] + cnt * dm * dm;
            }
          }
      }
      __syncthreads();
      comb[threadIdx.x] = acc;
      __syncthreads();
      float total = 0.f;
      if (threadIdx.x < h)
        for (int q = 0; q < parts; ++q) total += comb[q * h + threadIdx.x];
      if (pass == 0) {  // every thread of a column needs the mean for the second pass
        __syncthreads();
        if (threadIdx.x < h) comb[threadIdx.x] = total / (float)n_rows;
        __syncthreads();
        mean = comb[c];
      } else {
        m2 = total;
      }
    }
    if (threadIdx.x < h) {
      const float var = m2 / (float)n_rows;
      const float rstd = 1.f / sqrtf(var + eps);
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      if (running_var) {
        const float unbiased = n_rows > 1 ? m2 / (float)(n_rows - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      }
      mean_out[c] = mean;
      rstd_out[c] = rstd;
      scale_out[c] = gamma ? gamma[c] * rstd : rstd;
    }
    if (threadIdx.x == 0 && nbt) *nbt += 1;
  } else if (threadIdx.x < h) {
    const int c = threadIdx.x;
    const float mean = running_mean[c];
    const float rstd = 1.f / sqrtf(running_var[c] + eps);
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    scale_out[c] = gamma ? gamma[c] * rstd : rstd;
  }
}

// ------------------------------------------------------------------------------------------------ forward unit
template <int TR, int A_IN>  // tile rows: 64 or 32; input activation: compile-time code or kActRuntime
__global__ void __launch_bounds__(DT, CWN_FWD_MIN_CTAS) linear_fwd_kernel(const __grid_constant__ Group<cwn_linear_desc> g) {
  constexpr int TM = TR;
  constexpr int R = TR / 16;
  extern __shared__ __align__(16) float smem[];
  __shared__ cwn_linear_desc sd;
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  const cwn_linear_desc d = stage_desc(g, p, &sd);
  CWN_PHASE(1);
  const int t = blockIdx.x - g.start[p];
  const int col_tiles = (d.h + TN - 1) / TN;
  const int rt = t / col_tiles, ct = t % col_tiles;
  const int K = d.k0 + d.k1, K4 = round4(K), lda = K4 + 4;
  float* As = smem;               // [TM][lda]
  float* Bs = As + TM * lda;      // [K4][LDT]   Bs[k][n] = W[col0 + n][k]
  float* vin = Bs + K4 * LDT;     // [3][K4]
  const int64_t row0 = (int64_t)rt * TM;
  const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
  const int col0 = ct * TN;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;

  stage_input_vectors(d, vin, K, K4);
  {  // weight tile, transposed into Bs: lanes run along n so the shared-memory stores are conflict-free
    const bool wvec = ((reinterpret_cast<uintptr_t>(d.w) & 15u) == 0) && (d.ld_w % 4 == 0) && (K % 4 == 0);
    const int n = tid % TN;
    const bool live = col0 + n < d.h;
    const float* wrow = d.w + (int64_t)(col0 + n) * d.ld_w;
    if (wvec) {
      constexpr int NB = 4;
      for (int c0 = tid / TN; c0 < K4 / 4; c0 += NB * (DT / TN)) {
        float4 w[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int c4 = c0 + j * (DT / TN);
          w[j] = (live && c4 < K4 / 4) ? ldg_f4(wrow + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int c4 = c0 + j * (DT / TN);
          if (c4 >= K4 / 4) continue;
          Bs[(c4 * 4 + 0) * LDT + n] = w[j].x;
          Bs[(c4 * 4 + 1) * LDT + n] = w[j].y;
          Bs[(c4 * 4 + 2) * LDT + n] = w[j].z;
          Bs[(c4 * 4 + 3) * LDT + n] = w[j].w;
        }
      }
    } else {
      for (int k = tid / TN; k < K4; k += DT / TN) Bs[k * LDT + n] = (live && k < K) ? __ldg(wrow + k) : 0.f;
    }
  }
  __syncthreads();  // vin ready
  CWN_PHASE(2);
  load_input_tile<A_IN>(d, vin, K, K4, row0, rows, 0, K4, As, lda, input_vec_ok(d), TM);
  __syncthreads();
  CWN_PHASE(3);
  float acc[R][4] = {};
  tile_mma<R>(As, lda, Bs, LDT, K4, ty, tx, acc);
  CWN_PHASE(4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = col0 + tx * 4 + j;
    const float b = (d.bias && c < d.h) ? __ldg(d.bias + c) : 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i][j] += b;
  }
  const bool zvec = ((reinterpret_cast<uintptr_t>(d.z) & 15u) == 0) && (d.ld_z % 4 == 0) && (col0 + tx * 4 + 3 < d.h);
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = ty * R + i;
    if (r >= rows) continue;
    float* zrow = d.z + (row0 + r) * d.ld_z + col0 + tx * 4;
    if (zvec) {
      *reinterpret_cast<float4*>(zrow) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + tx * 4 + j < d.h) zrow[j] = acc[i][j];
    }
  }
  if (!d.stats) {
    if (d.bn_mean && !d.bn_training && blockIdx.x == g.start[p])  // eval: statistics are the running ones
      bn_finalize_body(nullptr, 0, d.n_rows, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 0, d.bn_running_mean,
                       d.bn_running_var, nullptr, d.bn_mean, d.bn_scale, d.bn_rstd, nullptr);
    return;
  }
  __syncthreads();  // As/Bs are dead: reuse the front of shared memory for the output tile
  CWN_PHASE(5);
  float* Ys = smem;              // [TM][LDT]
  float* red = smem + TM * LDT;  // [4][TN]
#pragma unroll
  for (int i = 0; i < R; ++i)
    *reinterpret_cast<float4*>(Ys + (ty * R + i) * LDT + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  // per-column (mean, M2) of this tile: 4 row groups per column, combined in a fixed order
  const int c = tid & (TN - 1), part = tid >> 6;
  float s = 0.f;
  for (int r = part; r < rows; r += 4) s += Ys[r * LDT + c];
  red[part * TN + c] = s;
  __syncthreads();
  const float mean = (((red[c] + red[TN + c]) + red[2 * TN + c]) + red[3 * TN + c]) / (float)rows;
  __syncthreads();
  float m2 = 0.f;
  for (int r = part; r < rows; r += 4) {
    const float dv = Ys[r * LDT + c] - mean;
    m2 = fmaf(dv, dv, m2);
  }
  red[part * TN + c] = m2;
  __syncthreads();
  if (part == 0 && col0 + c < d.h) {
    d.stats[((int64_t)rt * 2 + 0) * d.h + col0 + c] = mean;
    d.stats[((int64_t)rt * 2 + 1) * d.h + col0 + c] = ((red[c] + red[TN + c]) + red[2 * TN + c]) + red[3 * TN + c];
  }
  CWN_PHASE(6);
  if (d.bn_mean && d.counter) {
    const int total = g.start[p + 1] - g.start[p];
    const bool last_ = last_cta_of_problem(d.counter, total);
    CWN_PHASE(7);
    if (last_) {
      const int n_tiles = (int)((d.n_rows + TM - 1) / TM);
      bn_finalize_body(d.stats, n_tiles, d.n_rows, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 1, d.bn_running_mean,
                       d.bn_running_var, d.bn_num_batches_tracked, d.bn_mean, d.bn_scale, d.bn_rstd, smem,
                       TM * lda + K4 * LDT, TM);  // (this problem's share of the dynamic shared memory)
      if (threadIdx.x == 0) *d.counter = 0;
      CWN_PHASE(8);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward, fast path
// The same unit for the layouts the models actually produce (16-byte aligned operands, k0, k1, h multiples of 4).
// Per-phase clock64() stamps of the first version at the real-data shape (tools/phase_timing.py) showed a CTA spending
// its life in a chain of dependent phases — descriptor 1 400 cycles, vectors + W 3 600, X tile 2 500, FMAs 7 400,
// stores + statistics 2 100, fence 1 800 — i.e. latency, not throughput. Here every global operand of the tile (W rows,
// X rows, the input-transform vectors) is requested up front with cp.async and awaited ONCE; W stays in its natural
// [n][k] layout (no transposing scatter: a thread owns columns tx, tx+16, tx+32, tx+48, rows of W (K+4 floats apart)
// fall on distinct bank groups for 128-bit reads); the input transform runs in place in shared memory; the BatchNorm
// partials come straight from the accumulator registers (shuffle + one small shared array) instead of a staged tile.
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;  // 0 source bytes = zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int TR, int A_IN>
__global__ void __launch_bounds__(DT, CWN_FWD_MIN_CTAS) linear_fwd_fast_kernel(const __grid_constant__ Group<cwn_linear_desc> g) {
  constexpr int R = TR / 16;
  extern __shared__ __align__(16) float smem[];
  __shared__ cwn_linear_desc sd;
  __shared__ float red[8][TN];
  __shared__ float meanv[TN];
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  const cwn_linear_desc d = stage_desc(g, p, &sd);
  CWN_PHASE(1);
  const int t = blockIdx.x - g.start[p];
  const int col_tiles = (d.h + TN - 1) / TN;
  const int rt = t / col_tiles, ct = t - rt * col_tiles;
  const int K = d.k0 + d.k1, K4 = K / 4, ld = K + 4;
  float* Xs = smem;             // [TR][ld]   f_in(X) tile
  float* Ws = Xs + TR * ld;     // [TN][ld]   W[col0 + n][k]
  float* vin = Ws + TN * ld;    // [3][K]     mean | scale | beta of the input transform
  const int64_t row0 = (int64_t)rt * TR;
  const int rows = (int)((d.n_rows - row0 < TR) ? d.n_rows - row0 : TR);
  const int col0 = ct * TN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const bool transform = d.in_scale0 || d.in_scale1 || d.in_act != CWN_ACT_ID;

  for (int i = tid; i < TN * K4; i += DT) {
    const int n = i / K4, c = (i - n * K4) * 4;
    const bool ok = col0 + n < d.h;
    cp_async16(Ws + n * ld + c, ok ? d.w + (int64_t)(col0 + n) * d.ld_w + c : d.w, ok);
  }
  for (int i = tid; i < TR * K4; i += DT) {
    const int r = i / K4, c = (i - r * K4) * 4;
    const bool ok = r < rows;
    const float* src = (c < d.k0) ? d.x0 + (row0 + r) * d.ld_x0 + c : d.x1 + (row0 + r) * d.ld_x1 + (c - d.k0);
    cp_async16(Xs + r * ld + c, ok ? src : d.x0, ok);
  }
  if (transform)
    for (int i = tid; i < 3 * K4; i += DT) {
      const int which = i / K4, c = (i - which * K4) * 4;
      const bool first = c < d.k0;
      const float* base = which == 0 ? (first ? d.in_mean0 : d.in_mean1)
                        : which == 1 ? (first ? d.in_scale0 : d.in_scale1) : (first ? d.in_beta0 : d.in_beta1);
      const bool has = first ? d.in_scale0 != nullptr : d.in_scale1 != nullptr;
      if (has && base) {
        cp_async16(vin + which * K + c, base + (first ? c : c - d.k0), true);
      } else {
        const float f = which == 1 ? 1.f : 0.f;
        *reinterpret_cast<float4*>(vin + which * K + c) = make_float4(f, f, f, f);
      }
    }
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = col0 + tx + 16 * j;
    bias[j] = (d.bias && c < d.h) ? __ldg(d.bias + c) : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  CWN_PHASE(2);
  if (transform) {
    for (int i = tid; i < rows * K4; i += DT) {
      const int r = i / K4, c = (i - r * K4) * 4;
      float4 v = *reinterpret_cast<const float4*>(Xs + r * ld + c);
      const float4 mu = *reinterpret_cast<const float4*>(vin + c);
      const float4 sc = *reinterpret_cast<const float4*>(vin + K + c);
      const float4 be = *reinterpret_cast<const float4*>(vin + 2 * K + c);
      v.x = act_apply<A_IN>(d.in_act, (v.x - mu.x) * sc.x + be.x);
      v.y = act_apply<A_IN>(d.in_act, (v.y - mu.y) * sc.y + be.y);
      v.z = act_apply<A_IN>(d.in_act, (v.z - mu.z) * sc.z + be.z);
      v.w = act_apply<A_IN>(d.in_act, (v.w - mu.w) * sc.w + be.w);
      *reinterpret_cast<float4*>(Xs + r * ld + c) = v;
    }
    __syncthreads();
  }
  {
    CWN_PHASE(3);
    float acc[R][4] = {};
    {
      const float* a0 = Xs + (ty * R) * ld;
      const float* b0 = Ws + tx * ld;
  #pragma unroll 2
      for (int k = 0; k < K; k += 4) {
        float4 a[R], b[4];
  #pragma unroll
        for (int i = 0; i < R; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * ld + k);
  #pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(b0 + j * 16 * ld + k);
  #pragma unroll
        for (int i = 0; i < R; ++i)
  #pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
  #pragma unroll
        for (int i = 0; i < R; ++i)
  #pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
  #pragma unroll
        for (int i = 0; i < R; ++i)
  #pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
  #pragma unroll
        for (int i = 0; i < R; ++i)
  #pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
    }
    CWN_PHASE(4);
  #pragma unroll
    for (int i = 0; i < R; ++i) {
      const int r = ty * R + i;
  #pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] += bias[j];
        const int c = col0 + tx + 16 * j;
        if (r < rows && c < d.h) d.z[(row0 + r) * d.ld_z + c] = acc[i][j];
      }
    }
    CWN_PHASE(5);
    if (!d.stats) {
      if (d.bn_mean && !d.bn_training && blockIdx.x == g.start[p])  // eval: statistics are the running ones
        bn_finalize_body(nullptr, 0, d.n_rows, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 0, d.bn_running_mean,
                         d.bn_running_var, nullptr, d.bn_mean, d.bn_scale, d.bn_rstd, nullptr);
      return;
    }
    {  // per-column (mean, M2) of this tile from the accumulators: rows of a thread, the lane 16 apart, then the warps
      const int w = tid >> 5, lane = tid & 31;
      float sj[4];
  #pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
  #pragma unroll
        for (int i = 0; i < R; ++i) v += (ty * R + i < rows) ? acc[i][j] : 0.f;
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        sj[j] = v;
      }
      if (lane < 16)
  #pragma unroll
        for (int j = 0; j < 4; ++j) red[w][tx + 16 * j] = sj[j];
      __syncthreads();
      if (tid < TN) {
        float tot = 0.f;
  #pragma unroll
        for (int q = 0; q < 8; ++q) tot += red[q][tid];
        meanv[tid] = tot / (float)rows;
      }
      __syncthreads();
  #pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float mu = meanv[tx + 16 * j];
        float v = 0.f;
  #pragma unroll
        for (int i = 0; i < R; ++i) {
          const float dv = acc[i][j] - mu;
          v = (ty * R + i < rows) ? fmaf(dv, dv, v) : v;
        }
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        sj[j] = v;
      }
      if (lane < 16)
  #pragma unroll
        for (int j = 0; j < 4; ++j) red[w][tx + 16 * j] = sj[j];
      __syncthreads();
      if (tid < TN && col0 + tid < d.h) {
        float m2 = 0.f;
  #pragma unroll
        for (int q = 0; q < 8; ++q) m2 += red[q][tid];
        d.stats[((int64_t)rt * 2 + 0) * d.h + col0 + tid] = meanv[tid];
        d.stats[((int64_t)rt * 2 + 1) * d.h + col0 + tid] = m2;
      }
    }
  }
  CWN_PHASE(6);
  if (d.bn_mean && d.counter) {
    const int total = g.start[p + 1] - g.start[p];
    const bool last_ = last_cta_of_problem(d.counter, total);
    CWN_PHASE(7);
    if (last_) {
      const int n_tiles = (int)((d.n_rows + TR - 1) / TR);
      bn_finalize_body(d.stats, n_tiles, d.n_rows, d.h, d.bn_gamma, d.bn_eps, d.bn_momentum, 1, d.bn_running_mean,
                       d.bn_running_var, d.bn_num_batches_tracked, d.bn_mean, d.bn_scale, d.bn_rstd, smem,
                       TR * ld + TN * ld, TR);
      if (threadIdx.x == 0) *d.counter = 0;
      CWN_PHASE(8);
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN statistics
__global__ void __launch_bounds__(DT) bn_finalize_kernel(const __grid_constant__ Group<cwn_bn_desc> g) {
  __shared__ __align__(16) float stage[kStageFloats];
  const cwn_bn_desc& d = g.d[blockIdx.x];
  bn_finalize_body(d.stats, d.n_tiles, d.n_rows, d.h, d.gamma, d.eps, d.momentum, d.training, d.running_mean,
                   d.running_var, d.num_batches_tracked, d.mean, d.scale, d.rstd, stage, kStageFloats,
                   d.tile_rows > 0 ? d.tile_rows : TM);
}

// out = act((z - mean) * scale + beta): the layer output the neighbours gather from. One float4 per thread per
// step, every load of a step issued before the first use (the scalar version with an out-of-line activation call
// serialised 16 L2 round trips per thread: 9.5 us for 1.6 MB).
constexpr int kBnActV4 = 512;  // float4 elements per CTA on the vector path
template <int ACT>
__global__ void __launch_bounds__(DT) bn_act_kernel(const __grid_constant__ Group<cwn_bn_act_desc> g, int vec) {
  pdl_trigger();
  pdl_wait();
  const int p = find_problem(g, blockIdx.x);
  const cwn_bn_act_desc& d = g.d[p];
  if (vec) {
    const int h4 = d.h / 4;
    const int64_t total = d.n_rows * h4;
    const int64_t i0 = (int64_t)(blockIdx.x - g.start[p]) * kBnActV4 + threadIdx.x;
    constexpr int NB = kBnActV4 / DT;
    float4 z[NB], mu[NB], sc[NB], be[NB];
    int64_t off_out[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int64_t i = i0 + (int64_t)j * DT;
      const int64_t r = i / h4;
      const int c = (int)(i - r * h4) * 4;
      off_out[j] = -1;
      if (i < total) {
        z[j] = __ldcg(reinterpret_cast<const float4*>(d.z + r * d.ld_z + c));
        if (d.scale) {
          mu[j] = __ldcg(reinterpret_cast<const float4*>(d.mean + c));
          sc[j] = __ldcg(reinterpret_cast<const float4*>(d.scale + c));
          be[j] = d.beta ? ldg_f4(d.beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        off_out[j] = r * d.ld_out + c;
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (off_out[j] < 0) continue;
      float4 v = z[j];
      if (d.scale) {
        v.x = (v.x - mu[j].x) * sc[j].x + be[j].x; v.y = (v.y - mu[j].y) * sc[j].y + be[j].y;
        v.z = (v.z - mu[j].z) * sc[j].z + be[j].z; v.w = (v.w - mu[j].w) * sc[j].w + be[j].w;
      }
      v.x = act_apply<ACT>(d.act, v.x); v.y = act_apply<ACT>(d.act, v.y);
      v.z = act_apply<ACT>(d.act, v.z); v.w = act_apply<ACT>(d.act, v.w);
      *reinterpret_cast<float4*>(d.out + off_out[j]) = v;
    }
    return;
  }
  const int64_t row0 = (int64_t)(blockIdx.x - g.start[p]) * TM;
  const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
  for (int i = threadIdx.x; i < rows * d.h; i += DT) {
    const int r = i / d.h, c = i % d.h;
    float v = d.z[(row0 + r) * d.ld_z + c];
    if (d.scale) v = (v - d.mean[c]) * d.scale[c] + (d.beta ? d.beta[c] : 0.f);
    d.out[(row0 + r) * d.ld_out + c] = act_apply<ACT>(d.act, v);
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
    gy = g * act_grad_rt(d.act, y);
  } else {
    zhat = 0.f;
    gy = g * act_grad_rt(d.act, z);
  }
}

__device__ __forceinline__ void unit_bwd_finalize_body(const cwn_unit_bwd_desc& d, float* stage, int tile_rows) {
  const int n_tiles = (int)((d.n_rows + tile_rows - 1) / tile_rows);
  const int per_tile = 2 * d.h;
  const int chunk = kStageFloats / per_tile > 0 ? kStageFloats / per_tile : 1;
  float s1 = 0.f, s2 = 0.f;  // column threadIdx.x
  const bool fast = merge_fast_ok(d.red_partials, d.h);
  if (fast) {
    __syncthreads();
    sums_merge_fast(d.red_partials, n_tiles, d.h, stage, s1, s2);  // stage: >= 4 DT floats
  }
  for (int t0 = 0; t0 < n_tiles && !fast; t0 += chunk) {
    const int nt = (n_tiles - t0 < chunk) ? n_tiles - t0 : chunk;
    __syncthreads();
    for (int i = threadIdx.x; i < nt * per_tile; i += DT) stage[i] = __ldcg(d.red_partials + (int64_t)t0 * per_tile + i);
    __syncthreads();
    if (threadIdx.x < d.h)
      for (int t = 0; t < nt; ++t) {
        s1 += stage[t * per_tile + threadIdx.x];
        s2 += stage[t * per_tile + d.h + threadIdx.x];
      }
  }
  if (threadIdx.x < d.h) {
    const int c = threadIdx.x;
    const float n_live = (float)live_rows(d);  // (padding rows carry zero gradients: they add nothing to s1, s2)
    d.c1[c] = s1 / n_live;
    d.c2[c] = s2 / n_live;
    if (d.g_gamma) d.g_gamma[c] = d.accumulate_affine ? d.g_gamma[c] + s2 : s2;
    if (d.g_beta) d.g_beta[c] = d.accumulate_affine ? d.g_beta[c] + s1 : s1;
  }
}

template <int TR, int A_OUT>
__global__ void __launch_bounds__(DT) unit_bwd_reduce_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  constexpr int TM = TR;
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
    if (c < d.h) {
      const float mean = __ldg(d.mean + c), scale = __ldg(d.scale + c), rstd = __ldg(d.rstd + c);
      const float beta = d.beta ? __ldg(d.beta + c) : 0.f;
      const float* zp = d.z + row0 * d.ld_z + c;
      const float* gp = d.g_out + row0 * d.ld_g + c;
#pragma unroll 4
      for (int r = rg; r < rows; r += 4) {
        const float zc = __ldg(zp + r * d.ld_z) - mean;
        const float gy = __ldg(gp + r * d.ld_g) * act_grad<A_OUT>(d.act, zc * scale + beta);
        s1 += gy;
        s2 = fmaf(gy, zc * rstd, s2);
      }
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
  if (d.counter) {
    __shared__ __align__(16) float stage[kStageFloats];
    if (last_cta_of_problem(d.counter, g.start[p + 1] - g.start[p])) {
      unit_bwd_finalize_body(d, stage, TM);
      if (threadIdx.x == 0) *d.counter = 0;
    }
  }
}

// BN-backward column sums, vector path (h % 4 == 0, 16-byte aligned rows): thread (q, rg) owns float4 column q and the
// rows rg, rg + RG, ... of the tile (RG = DT / (h/4) row groups) and has its 128-bit loads of z and g_out in flight
// together (the scalar version walked 16 rows per thread four loads at a time: 8-10 us per launch at the real-data
// shape for 3 MB of input); the row groups are then combined in order through shared memory.
template <int TR, int A_OUT>
__global__ void __launch_bounds__(DT) unit_bwd_reduce_fast_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  __shared__ __align__(16) float part[2 * 4 * DT];   // [2][RG][h], RG * h = 4 DT
  __shared__ __align__(16) float stage[6 * DT];      // merge scratch of the last CTA
  pdl_trigger();
  pdl_wait();
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc& d = g.d[p];
  const int tile = blockIdx.x - g.start[p];
  const int h = d.h, h4 = h / 4, RG = DT / h4;
  const int64_t row0 = (int64_t)tile * TR;
  const int rows = (int)((d.n_rows - row0 < TR) ? d.n_rows - row0 : TR);
  const int q = threadIdx.x % h4, rg = threadIdx.x / h4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (rg < RG) {
    const int c = q * 4;
    const float4 mean = ldg_f4(d.mean + c), scale = ldg_f4(d.scale + c), rstd = ldg_f4(d.rstd + c);
    const float4 beta = d.beta ? ldg_f4(d.beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* zp = d.z + row0 * d.ld_z + c;
    const float* gp = d.g_out + row0 * d.ld_g + c;
    constexpr int NB = 4;
    for (int r0 = rg; r0 < rows; r0 += NB * RG) {
      float4 zv[NB], gv[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int r = r0 + j * RG;
        zv[j] = gv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
          zv[j] = __ldcg(reinterpret_cast<const float4*>(zp + (int64_t)r * d.ld_z));
          gv[j] = __ldcg(reinterpret_cast<const float4*>(gp + (int64_t)r * d.ld_g));
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (r0 + j * RG >= rows) continue;
        float zc, gy;
        zc = zv[j].x - mean.x; gy = gv[j].x * act_grad<A_OUT>(d.act, zc * scale.x + beta.x); s1.x += gy; s2.x = fmaf(gy, zc * rstd.x, s2.x);
        zc = zv[j].y - mean.y; gy = gv[j].y * act_grad<A_OUT>(d.act, zc * scale.y + beta.y); s1.y += gy; s2.y = fmaf(gy, zc * rstd.y, s2.y);
        zc = zv[j].z - mean.z; gy = gv[j].z * act_grad<A_OUT>(d.act, zc * scale.z + beta.z); s1.z += gy; s2.z = fmaf(gy, zc * rstd.z, s2.z);
        zc = zv[j].w - mean.w; gy = gv[j].w * act_grad<A_OUT>(d.act, zc * scale.w + beta.w); s1.w += gy; s2.w = fmaf(gy, zc * rstd.w, s2.w);
      }
    }
    reinterpret_cast<float4*>(part)[rg * h4 + q] = s1;
    reinterpret_cast<float4*>(part + 4 * DT)[rg * h4 + q] = s2;
  }
  __syncthreads();
  if (threadIdx.x < h) {
    const int c = threadIdx.x;
    d.red_partials[((int64_t)tile * 2 + 0) * h + c] = merge_column(part, RG, h, c);
    d.red_partials[((int64_t)tile * 2 + 1) * h + c] = merge_column(part + 4 * DT, RG, h, c);
  }
  if (d.counter && last_cta_of_problem(d.counter, g.start[p + 1] - g.start[p])) {
    unit_bwd_finalize_body(d, stage, TR);
    if (threadIdx.x == 0) *d.counter = 0;
  }
}

__global__ void __launch_bounds__(DT) unit_bwd_finalize_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  __shared__ __align__(16) float stage[kStageFloats];
  const cwn_unit_bwd_desc& d = g.d[blockIdx.x];
  if (!d.has_bn) return;
  unit_bwd_finalize_body(d, stage, d.tile_rows > 0 ? d.tile_rows : TM);
}

// g_z tile -> input gradient (g_z W) and per-CTA partial weight gradient (g_z^T f_in(X)); CTA j of a problem strides
// over the row tiles j, j + n_ctas, ... and owns slab j of the partial buffers (plain read-modify-write, no atomics).
template <int TR, int A_IN, int A_OUT>
__global__ void __launch_bounds__(DT, CWN_BWD_MIN_CTAS) unit_bwd_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  constexpr int TM = TR;          // rows per tile
  constexpr int R = TR / 16;      // rows per thread of the input-gradient tile
  constexpr int LDR = TR + 4;     // leading dimension of g_z^T (inner dimension = rows)
  extern __shared__ __align__(16) float smem[];
  __shared__ cwn_unit_bwd_desc sd;
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc d = stage_desc(g, p, &sd);
  CWN_PHASE(1);
  const int j = blockIdx.x - g.start[p];
  const int K = d.k0 + d.k1, K4 = round4(K), H4 = round4(d.h), ldg = H4 + 4;
  const int m_tiles = (d.h + 63) / 64;
  float* Gz = smem;                        // [TM][ldg]            g_z[r][c]
  float* GzT = Gz + TM * ldg;              // [m_tiles*64][LDR]    g_z^T[c][r]
  float* Ain = GzT + m_tiles * 64 * LDR;   // [TM][LDT]            f_in(X)[r][k chunk]
  float* Ws = Ain + TM * LDT;              // [H4][LDT]            W[c][k chunk]
  float* vin = Ws + H4 * LDT;              // [3][K4]              input transform
  float* vout = vin + 3 * K4;              // [6][H4]              mean, scale, rstd, beta, c1, c2 of this unit
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int n_tiles = (int)((d.n_rows + TM - 1) / TM);
  float* wpart = d.w_partials + (int64_t)j * d.h * K;
  float* bpart = d.b_partials + (int64_t)j * d.h;
  const bool in_vec = input_vec_ok(d);
  const bool w_vec = ((reinterpret_cast<uintptr_t>(d.w) & 15u) == 0) && (d.ld_w % 4 == 0) && (K % 4 == 0);
  const bool g_vec = ((reinterpret_cast<uintptr_t>(d.z) & 15u) == 0) && (d.ld_z % 4 == 0) && (d.h % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(d.g_out) & 15u) == 0) && (d.ld_g % 4 == 0);
  const bool gi_vec = (d.k0 % 4 == 0) && (d.k1 % 4 == 0) &&
                      (!d.g_in0 || (((reinterpret_cast<uintptr_t>(d.g_in0) & 15u) == 0) && d.ld_gi0 % 4 == 0)) &&
                      (!d.g_in1 || (((reinterpret_cast<uintptr_t>(d.g_in1) & 15u) == 0) && d.ld_gi1 % 4 == 0));

  stage_input_vectors(d, vin, K, K4);
  for (int c = tid; c < H4; c += DT) {
    const bool live = c < d.h && d.has_bn;
    vout[c] = live ? __ldg(d.mean + c) : 0.f;
    vout[H4 + c] = live ? __ldg(d.scale + c) : 1.f;
    vout[2 * H4 + c] = live ? __ldg(d.rstd + c) : 0.f;
    vout[3 * H4 + c] = (live && d.beta) ? __ldg(d.beta + c) : 0.f;
    vout[4 * H4 + c] = live ? d.c1[c] : 0.f;
    vout[5 * H4 + c] = live ? d.c2[c] : 0.f;
  }
  for (int i = tid; i < m_tiles * 64 * LDR; i += DT) GzT[i] = 0.f;  // rows c >= h stay zero for good
  bool first = true;
  for (int tile = j; tile < n_tiles; tile += d.n_ctas, first = false) {
    const int64_t row0 = (int64_t)tile * TM;
    const int rows = (int)((d.n_rows - row0 < TM) ? d.n_rows - row0 : TM);
    __syncthreads();
    if (first) CWN_PHASE(2);
    // g_z = scale * (g_out act'(y) - c1 - zhat c2) of this tile (plain g_out act'(z) without BatchNorm)
    if (g_vec) {
      const int nc4 = H4 / 4;
      const int total = TM * nc4;
      constexpr int NB = 2;  // two operands per element: 4 loads in flight
      for (int base = tid; base < total; base += NB * DT) {
        float4 zv[NB], gv[NB];
        int rr[NB], cc[NB];
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          const int i = base + jj * DT;
          rr[jj] = i / nc4;
          cc[jj] = (i - rr[jj] * nc4) * 4;
          zv[jj] = gv[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < total && rr[jj] < rows) {
            zv[jj] = ldg_f4(d.z + (row0 + rr[jj]) * d.ld_z + cc[jj]);
            gv[jj] = ldg_f4(d.g_out + (row0 + rr[jj]) * d.ld_g + cc[jj]);
          }
        }
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          const int i = base + jj * DT;
          if (i >= total) continue;
          const int c = cc[jj];
          float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr[jj] < rows) {
            const float zz[4] = {zv[jj].x, zv[jj].y, zv[jj].z, zv[jj].w}, gg[4] = {gv[jj].x, gv[jj].y, gv[jj].z, gv[jj].w};
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float zc = zz[q] - vout[c + q];
              const float gy = gg[q] * act_grad<A_OUT>(d.act, zc * vout[H4 + c + q] + vout[3 * H4 + c + q]);
              o[q] = d.has_bn ? vout[H4 + c + q] * (gy - vout[4 * H4 + c + q] - zc * vout[2 * H4 + c + q] * vout[5 * H4 + c + q]) : gy;
            }
            out = make_float4(o[0], o[1], o[2], o[3]);
          }
          *reinterpret_cast<float4*>(Gz + rr[jj] * ldg + c) = out;
        }
      }
    } else {
      for (int i = tid; i < TM * H4; i += DT) {
        const int r = i / H4, c = i % H4;
        float gz = 0.f;
        if (r < rows && c < d.h) {
          const float zc = __ldg(d.z + (row0 + r) * d.ld_z + c) - vout[c];
          const float gy = __ldg(d.g_out + (row0 + r) * d.ld_g + c) * act_grad<A_OUT>(d.act, zc * vout[H4 + c] + vout[3 * H4 + c]);
          gz = d.has_bn ? vout[H4 + c] * (gy - vout[4 * H4 + c] - zc * vout[2 * H4 + c] * vout[5 * H4 + c]) : gy;
        }
        Gz[r * ldg + c] = gz;
      }
    }
    __syncthreads();
    if (first) CWN_PHASE(3);
    {  // transpose into GzT (lanes along r: conflict-free stores) and the bias-gradient partial (column sums)
      const int r = tid & (TM - 1);
      for (int c = tid / TM; c < d.h; c += DT / TM) GzT[c * LDR + r] = Gz[r * ldg + c];
      for (int c = tid; c < d.h; c += DT) {
        float s_ = 0.f;
        for (int rr = 0; rr < rows; ++rr) s_ += Gz[rr * ldg + c];
        bpart[c] = first ? s_ : bpart[c] + s_;
      }
    }
    if (first) CWN_PHASE(4);
    for (int kc = 0; kc < K; kc += TN) {
      __syncthreads();
      if (w_vec) {  // W[c][kc + k], natural layout
        constexpr int NB = 4;
        for (int base = tid; base < H4 * (TN / 4); base += NB * DT) {
          float4 w[NB];
#pragma unroll
          for (int jj = 0; jj < NB; ++jj) {
            const int i = base + jj * DT;
            const int c = i / (TN / 4), k = (i % (TN / 4)) * 4;
            w[jj] = (i < H4 * (TN / 4) && c < d.h && kc + k < K) ? ldg_f4(d.w + (int64_t)c * d.ld_w + kc + k)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int jj = 0; jj < NB; ++jj) {
            const int i = base + jj * DT;
            if (i < H4 * (TN / 4)) *reinterpret_cast<float4*>(Ws + (i / (TN / 4)) * LDT + (i % (TN / 4)) * 4) = w[jj];
          }
        }
      } else {
        for (int i = tid; i < H4 * TN; i += DT) {
          const int c = i / TN, k = i % TN;
          Ws[c * LDT + k] = (c < d.h && kc + k < K) ? __ldg(d.w + (int64_t)c * d.ld_w + kc + k) : 0.f;
        }
      }
      load_input_tile<A_IN>(d, vin, K, K4, row0, rows, kc, TN, Ain, LDT, in_vec, TM);
      __syncthreads();
      if (first && kc == 0) CWN_PHASE(5);
      if (d.g_in0 || d.g_in1) {  // input gradient chunk: [TM rows] x [64 k] = Gz [TM x h] * Ws [h x 64]
        float acc[R][4] = {};
        tile_mma<R>(Gz, ldg, Ws, LDT, H4, ty, tx, acc);
        const int kq = kc + tx * 4;  // first of this thread's four columns
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int r = ty * R + i;
          if (r >= rows) continue;
          if (gi_vec && kq + 3 < K) {  // the four columns lie in one block (k0 % 4 == 0): one 128-bit store
            float* base = (kq < d.k0) ? (d.g_in0 ? d.g_in0 + (row0 + r) * d.ld_gi0 + kq : nullptr)
                                      : (d.g_in1 ? d.g_in1 + (row0 + r) * d.ld_gi1 + (kq - d.k0) : nullptr);
            if (base) {
              float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
              if (d.accumulate_in) v = f4_add(*reinterpret_cast<const float4*>(base), v);
              *reinterpret_cast<float4*>(base) = v;
            }
            continue;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = kq + q;
            if (k >= K) continue;
            float* dst = (k < d.k0) ? (d.g_in0 ? d.g_in0 + (row0 + r) * d.ld_gi0 + k : nullptr)
                                    : (d.g_in1 ? d.g_in1 + (row0 + r) * d.ld_gi1 + (k - d.k0) : nullptr);
            if (dst) *dst = d.accumulate_in ? *dst + acc[i][q] : acc[i][q];
          }
        }
      }
      if (first && kc == 0) CWN_PHASE(6);
      for (int mt = 0; mt < m_tiles; ++mt) {  // weight gradient chunk: [64 c] x [64 k] = GzT [64 x TM] * Ain [TM x 64]
        float acc[4][4] = {};
        tile_mma<4>(GzT + mt * 64 * LDR, LDR, Ain, LDT, TM, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = mt * 64 + ty * 4 + i;
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
      if (first && kc == 0) CWN_PHASE(7);
    }
  }
  CWN_PHASE(8);
}

// ------------------------------------------------------------------------------------------------ backward, fast path
// unit_bwd_kernel for 16-byte aligned operands with k0, k1, h multiples of 4, restructured like linear_fwd_fast_kernel:
// W (whole), the input-transform and BatchNorm vectors are requested once, the z / g_out / X tiles of a row tile with
// one cp.async batch; g_z is formed in place over the g_out tile; its transpose reuses the z tile's shared memory; the
// bias partial is read off the transposed tile; weight-gradient partials leave as 128-bit stores. Same FMA order as
// the generic kernel, so the two agree bit for bit.
template <int TR, int A_IN, int A_OUT>
__global__ void __launch_bounds__(DT, CWN_BWD_MIN_CTAS) unit_bwd_fast_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g) {
  constexpr int R = TR / 16;
  constexpr int LDR = TR + 4;
  extern __shared__ __align__(16) float smem[];
  __shared__ cwn_unit_bwd_desc sd;
  CWN_PHASE(0);
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc d = stage_desc(g, p, &sd);
  CWN_PHASE(1);
  const int j = blockIdx.x - g.start[p];
  const int K = d.k0 + d.k1, K4 = K / 4, ldk = K + 4, h = d.h, h4 = h / 4, ldh = h + 4;
  const int m_tiles = (h + 63) / 64;
  const int zt_floats = (TR * ldh > m_tiles * 64 * LDR) ? TR * ldh : m_tiles * 64 * LDR;
  float* Gz = smem;                 // [TR][ldh]            g_out tile, then g_z
  float* ZT = Gz + TR * ldh;        // [TR][ldh] z tile, then [m_tiles*64][LDR] g_z^T
  float* Xs = ZT + zt_floats;       // [TR][ldk]            f_in(X)
  float* Ws = Xs + TR * ldk;        // [h][ldk]             W
  float* vin = Ws + h * ldk;        // [3][K]
  float* vout = vin + 3 * K;        // [6][h]   mean, scale, rstd, beta, c1, c2
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int n_tiles = (int)((d.n_rows + TR - 1) / TR);
  float* wpart = d.w_partials + (int64_t)j * h * K;
  float* bpart = d.b_partials + (int64_t)j * h;
  const bool transform = d.in_scale0 || d.in_scale1 || d.in_act != CWN_ACT_ID;

  for (int i = tid; i < h * K4; i += DT) {
    const int c = i / K4, k = (i - c * K4) * 4;
    cp_async16(Ws + c * ldk + k, d.w + (int64_t)c * d.ld_w + k, true);
  }
  if (transform)
    for (int i = tid; i < 3 * K4; i += DT) {
      const int which = i / K4, c = (i - which * K4) * 4;
      const bool first = c < d.k0;
      const float* base = which == 0 ? (first ? d.in_mean0 : d.in_mean1)
                        : which == 1 ? (first ? d.in_scale0 : d.in_scale1) : (first ? d.in_beta0 : d.in_beta1);
      const bool has = first ? d.in_scale0 != nullptr : d.in_scale1 != nullptr;
      if (has && base) {
        cp_async16(vin + which * K + c, base + (first ? c : c - d.k0), true);
      } else {
        const float f = which == 1 ? 1.f : 0.f;
        *reinterpret_cast<float4*>(vin + which * K + c) = make_float4(f, f, f, f);
      }
    }
  for (int c = tid; c < h; c += DT) {
    const bool live = d.has_bn;
    vout[c] = live ? __ldg(d.mean + c) : 0.f;
    vout[h + c] = live ? __ldg(d.scale + c) : 1.f;
    vout[2 * h + c] = live ? __ldg(d.rstd + c) : 0.f;
    vout[3 * h + c] = (live && d.beta) ? __ldg(d.beta + c) : 0.f;
    vout[4 * h + c] = live ? __ldcg(d.c1 + c) : 0.f;
    vout[5 * h + c] = live ? __ldcg(d.c2 + c) : 0.f;
  }
  bool first = true;
  for (int tile = j; tile < n_tiles; tile += d.n_ctas, first = false) {
    const int64_t row0 = (int64_t)tile * TR;
    const int rows = (int)((d.n_rows - row0 < TR) ? d.n_rows - row0 : TR);
    __syncthreads();  // the previous tile's readers are done with the shared tiles
    for (int i = tid; i < TR * h4; i += DT) {
      const int r = i / h4, c = (i - r * h4) * 4;
      const bool ok = r < rows;
      cp_async16(ZT + r * ldh + c, ok ? d.z + (row0 + r) * d.ld_z + c : d.z, ok);
      cp_async16(Gz + r * ldh + c, ok ? d.g_out + (row0 + r) * d.ld_g + c : d.g_out, ok);
    }
    for (int i = tid; i < TR * K4; i += DT) {
      const int r = i / K4, c = (i - r * K4) * 4;
      const bool ok = r < rows;
      const float* src = (c < d.k0) ? d.x0 + (row0 + r) * d.ld_x0 + c : d.x1 + (row0 + r) * d.ld_x1 + (c - d.k0);
      cp_async16(Xs + r * ldk + c, ok ? src : d.x0, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    if (first) CWN_PHASE(2);
    // g_z = scale * (g_out act'(y) - c1 - zhat c2), in place over the g_out tile (plain g_out act'(z) without BatchNorm)
    for (int i = tid; i < TR * h4; i += DT) {
      const int r = i / h4, c = (i - r * h4) * 4;
      float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        const float4 zv = *reinterpret_cast<const float4*>(ZT + r * ldh + c);
        const float4 gv = *reinterpret_cast<const float4*>(Gz + r * ldh + c);
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float zc = zz[q] - vout[c + q];
          const float gy = gg[q] * act_grad<A_OUT>(d.act, zc * vout[h + c + q] + vout[3 * h + c + q]);
          o[q] = d.has_bn ? vout[h + c + q] * (gy - vout[4 * h + c + q] - zc * vout[2 * h + c + q] * vout[5 * h + c + q]) : gy;
        }
        out = make_float4(o[0], o[1], o[2], o[3]);
      }
      *reinterpret_cast<float4*>(Gz + r * ldh + c) = out;
    }
    if (transform)
      for (int i = tid; i < rows * K4; i += DT) {
        const int r = i / K4, c = (i - r * K4) * 4;
        float4 v = *reinterpret_cast<const float4*>(Xs + r * ldk + c);
        const float4 mu = *reinterpret_cast<const float4*>(vin + c);
        const float4 sc = *reinterpret_cast<const float4*>(vin + K + c);
        const float4 be = *reinterpret_cast<const float4*>(vin + 2 * K + c);
        v.x = act_apply<A_IN>(d.in_act, (v.x - mu.x) * sc.x + be.x);
        v.y = act_apply<A_IN>(d.in_act, (v.y - mu.y) * sc.y + be.y);
        v.z = act_apply<A_IN>(d.in_act, (v.z - mu.z) * sc.z + be.z);
        v.w = act_apply<A_IN>(d.in_act, (v.w - mu.w) * sc.w + be.w);
        *reinterpret_cast<float4*>(Xs + r * ldk + c) = v;
      }
    __syncthreads();
    if (first) CWN_PHASE(3);
    {  // transpose g_z into the (dead) z tile; rows c >= h of the last 64-row block are zero
      const int r = tid & (TR - 1);
      for (int c = tid / TR; c < m_tiles * 64; c += DT / TR) ZT[c * LDR + r] = c < h ? Gz[r * ldh + c] : 0.f;
    }
    __syncthreads();
    if (first) CWN_PHASE(4);
    if (tid < h) {  // bias-gradient partial = column sums of g_z, rows in order
      const float4* row = reinterpret_cast<const float4*>(ZT + tid * LDR);
      float s_ = 0.f;
#pragma unroll 4
      for (int q = 0; q < TR / 4; ++q) {
        const float4 v = row[q];
        s_ += v.x; s_ += v.y; s_ += v.z; s_ += v.w;
      }
      bpart[tid] = first ? s_ : bpart[tid] + s_;
    }
    for (int kc = 0; kc < K; kc += TN) {
      {
        if (d.g_in0 || d.g_in1) {  // input gradient chunk: [TR rows] x [64 k] = Gz [TR x h] * Ws [h x 64]
          float acc[R][4] = {};
          tile_mma<R>(Gz, ldh, Ws + kc, ldk, h, ty, tx, acc);
          const int kq = kc + tx * 4;
          if (kq < K) {
            float* base = (kq < d.k0) ? (d.g_in0 ? d.g_in0 + row0 * d.ld_gi0 + kq : nullptr)
                                      : (d.g_in1 ? d.g_in1 + row0 * d.ld_gi1 + (kq - d.k0) : nullptr);
            const int64_t ldo = (kq < d.k0) ? d.ld_gi0 : d.ld_gi1;
            if (base)
  #pragma unroll
              for (int i = 0; i < R; ++i) {
                const int r = ty * R + i;
                if (r < rows) {
                  float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                  if (d.accumulate_in) v = f4_add(*reinterpret_cast<const float4*>(base + r * ldo), v);
                  *reinterpret_cast<float4*>(base + r * ldo) = v;
                }
              }
          }
        }
        if (first && kc == 0) CWN_PHASE(5);
        for (int mt = 0; mt < m_tiles; ++mt) {  // weight gradient chunk: [64 c] x [64 k] = GzT [64 x TR] * Xs [TR x 64]
          float acc[4][4] = {};
          tile_mma<4>(ZT + mt * 64 * LDR, LDR, Xs + kc, ldk, TR, ty, tx, acc);
          const int k = kc + tx * 4;
          if (k < K)
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = mt * 64 + ty * 4 + i;
              if (c >= h) continue;
              float4* dst = reinterpret_cast<float4*>(wpart + (int64_t)c * K + k);
              float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
              if (!first) v = f4_add(*dst, v);
              *dst = v;
            }
        }
        if (first && kc == 0) CWN_PHASE(6);
      }
    }
  }
  CWN_PHASE(8);
}

// Ordered sum of the per-CTA partial slabs -> dW, db. Vector path: a CTA covers 64 float4 units x 4 slab groups;
// thread (u, sg) sums slabs sg, sg + 4, ... with batched 128-bit loads, the four group sums are combined in order.
// (One thread per element walking ~100 slabs with scalar loads was 8-9 us per launch.)
__global__ void __launch_bounds__(DT) wgrad_finalize_kernel(const __grid_constant__ Group<cwn_unit_bwd_desc> g, int vec) {
  pdl_trigger();
  pdl_wait();
  const int p = find_problem(g, blockIdx.x);
  const cwn_unit_bwd_desc& d = g.d[p];
  const int K = d.k0 + d.k1;
  if (vec) {
    __shared__ float4 comb[3][64];
    const int u = threadIdx.x & 63, sg = threadIdx.x >> 6;
    const int64_t w4 = (int64_t)d.h * K / 4;
    const int b4 = d.g_b ? d.h / 4 : 0;
    const int64_t unit = (int64_t)(blockIdx.x - g.start[p]) * 64 + u;
    const bool is_w = unit < w4, is_b = !is_w && unit < w4 + b4;
    const float* src = is_w ? d.w_partials + unit * 4 : d.b_partials + (unit - w4) * 4;
    const int64_t slab = is_w ? (int64_t)d.h * K : d.h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (is_w || is_b) {
      constexpr int NB = 8;
      for (int j0 = sg; j0 < d.n_ctas; j0 += 4 * NB) {
        float4 v[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const int j = j0 + 4 * q;
          v[q] = j < d.n_ctas ? __ldcg(reinterpret_cast<const float4*>(src + (int64_t)j * slab)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) acc = f4_add(acc, v[q]);
      }
    }
    if (sg > 0) comb[sg - 1][u] = acc;
    __syncthreads();
    if (sg == 0 && (is_w || is_b)) {
      acc = f4_add(f4_add(f4_add(acc, comb[0][u]), comb[1][u]), comb[2][u]);
      float* dst = is_w ? d.g_w + (unit * 4 / K) * d.ld_gw + (unit * 4 % K) : d.g_b + (unit - w4) * 4;
      if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        *d4 = d.accumulate_w ? f4_add(*d4, acc) : acc;
      } else {
        const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = d.accumulate_w ? dst[e] + a[e] : a[e];
      }
    }
    return;
  }
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

}  // namespace cwn
#include "dense_tc5.cuh"
namespace cwn {

// ------------------------------------------------------------------------------------------------ host side
template <class Kernel>
static int ensure_smem(Kernel kernel, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return CWN_OK;
  if (bytes > 227 * 1024) return fail(CWN_E_SHAPE, "dense kernels: K or h too large for shared memory");
  return cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what);
}

// common compile-time activation code of a group (identity / ReLU), else the runtime fallback
template <class D, class Get>
static int group_act(const D* descs, int n, Get get) {
  const int a = get(descs[0]);
  if (a != CWN_ACT_ID && a != CWN_ACT_RELU) return kActRuntime;
  for (int i = 1; i < n; ++i)
    if (get(descs[i]) != a) return kActRuntime;
  return a;
}

// tile rows of a group: every problem must agree (0 means 64)
template <class D>
static int group_tile_rows(const D* descs, int n, int& tr) {
  tr = descs[0].tile_rows > 0 ? descs[0].tile_rows : 64;
  if (tr != 64 && tr != 32) return fail(CWN_E_SHAPE, "tile_rows must be 64 or 32");
  for (int i = 1; i < n; ++i)
    if ((descs[i].tile_rows > 0 ? descs[i].tile_rows : 64) != tr) return fail(CWN_E_SHAPE, "tile_rows differs inside a group");
  return CWN_OK;
}

static int check_group(const void* descs, int n, const char* what) {
  if (n < 0 || n > CWN_MAX_GROUP) return fail(CWN_E_SHAPE, what);
  if (n > 0 && !descs) return fail(CWN_E_NULL, what);
  return CWN_OK;
}

}  // namespace cwn

using namespace cwn;

// Test hook: route every grouped dense call through the generic (any shape / alignment) kernels, so the parity tests
// can compare the two paths on the same inputs.
static int g_force_generic_dense = 0;  // bit 0: forward entry points, bit 1: backward entry points
extern "C" int cwn_debug_force_generic_dense(int32_t mask) {
  g_force_generic_dense = mask;
  return CWN_OK;
}

#ifdef CWN_PHASE_TIMING
extern "C" int cwn_debug_set_phase_buffer(long long* buf) {
  return cuda_status(cudaMemcpyToSymbol(g_phase_buf, &buf, sizeof(buf)), "cwn_debug_set_phase_buffer");
}
#endif

extern "C" int cwn_linear_fwd_grouped(const cwn_linear_desc* descs, int32_t n, cwn_stream_t stream) {
  int rc = check_group(descs, n, "cwn_linear_fwd_grouped");
  if (rc || n == 0) return rc;
  int tr;
  if ((rc = group_tile_rows(descs, n, tr))) return rc;
  Group<cwn_linear_desc> g;
  g.n = n;
  int total = 0;
  size_t smem = ((size_t)tr * LDT + 4 * TN) * sizeof(float);
  for (int i = 0; i < n; ++i) {
    const cwn_linear_desc& d = descs[i];
    if (d.n_rows < 0 || d.h <= 0 || d.k0 <= 0 || d.k1 < 0 || d.n_rows > (int64_t)INT32_MAX * TM)
      return fail(CWN_E_SHAPE, "cwn_linear_fwd_grouped: bad shape");
    if (d.n_rows > 0 && (!d.x0 || !d.w || !d.z || (d.k1 > 0 && !d.x1))) return fail(CWN_E_NULL, "cwn_linear_fwd_grouped: operand");
    g.d[i] = d;
    g.start[i] = total;
    total += (int)((d.n_rows + tr - 1) / tr) * ((d.h + TN - 1) / TN);
    const int K4 = round4(d.k0 + d.k1);
    const size_t need = ((size_t)tr * (K4 + 4) + (size_t)K4 * LDT + 3 * (size_t)K4) * sizeof(float);
    if (need > smem) smem = need;
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  const int a_in = group_act(descs, n, [](const cwn_linear_desc& d) { return d.in_act; });
  // tensor-core path (tcgen05, 3xTF32 over split accumulators): 64-row tiles, every problem of the group must qualify
  if (tr == 64 && t5_enabled() && !(g_force_generic_dense & 1)) {
    bool ok = true;
    size_t smem5 = 0;
    for (int i = 0; i < n && ok; ++i) ok = descs[i].n_rows == 0 || t5_fwd_ok(descs[i], smem5);
    if (ok && smem5 > 0) {
      int total5 = 0;
      for (int i = 0; i < n; ++i) {
        g.start[i] = total5;
        total5 += (int)((descs[i].n_rows + T5R - 1) / T5R);
      }
      g.start[n] = total5;
#define CWN_LAUNCH_T5(AV)                                                                                            \
  {                                                                                                                  \
    if ((rc = ensure_smem(linear_fwd_tc5_kernel<AV>, smem5, "cudaFuncSetAttribute(linear_fwd_tc5_kernel)"))) return rc; \
    if ((rc = cuda_status(launch_pdl(linear_fwd_tc5_kernel<AV>, total5, T5T, smem5, (cudaStream_t)stream, g), "linear_fwd_tc5_kernel"))) return rc; \
  }
      if (a_in == CWN_ACT_ID) CWN_LAUNCH_T5(CWN_ACT_ID)
      else if (a_in == CWN_ACT_RELU) CWN_LAUNCH_T5(CWN_ACT_RELU)
      else CWN_LAUNCH_T5(kActRuntime)
#undef CWN_LAUNCH_T5
      return launched("linear_fwd_tc5_kernel");
    }
  }
  for (int i = 0; i < n; ++i)
    if (descs[i].n_rows_live) return fail(CWN_E_SHAPE, "cwn_linear_fwd_grouped: n_rows_live needs the tensor-core path (h, K powers of two <= 128, tile_rows 64)");
  bool fast = !(g_force_generic_dense & 1);
  size_t smem_fast = 0;
  for (int i = 0; i < n && fast; ++i) {
    const cwn_linear_desc& d = descs[i];
    if (d.n_rows == 0) continue;
    auto vec_ok = [](const float* a, const float* b, const float* c) { return aligned16(a) && aligned16(b) && aligned16(c); };
    fast = d.k0 % 4 == 0 && d.k1 % 4 == 0 && aligned16(d.x0) && d.ld_x0 % 4 == 0 &&
           (d.k1 == 0 || (aligned16(d.x1) && d.ld_x1 % 4 == 0)) && aligned16(d.w) && d.ld_w % 4 == 0 &&
           (!d.in_scale0 || vec_ok(d.in_mean0, d.in_scale0, d.in_beta0)) &&
           (!d.in_scale1 || vec_ok(d.in_mean1, d.in_scale1, d.in_beta1)) &&
           (!d.stats || (d.h % 4 == 0 && aligned16(d.stats)));
    const int K = d.k0 + d.k1;
    const size_t need = ((size_t)(tr + TN) * (K + 4) + 3 * (size_t)K) * sizeof(float);
    if (need > smem_fast) smem_fast = need;
  }
  if (fast && smem_fast <= 200 * 1024) {
#define CWN_LAUNCH_FAST(TRV, AV)                                                                                   \
  {                                                                                                                \
    if ((rc = ensure_smem(linear_fwd_fast_kernel<TRV, AV>, smem_fast, "cudaFuncSetAttribute(linear_fwd_fast_kernel)"))) return rc; \
    linear_fwd_fast_kernel<TRV, AV><<<total, DT, smem_fast, (cudaStream_t)stream>>>(g);                            \
  }
#define CWN_FAST_BY_ACT(TRV)                                     \
  if (a_in == CWN_ACT_ID) CWN_LAUNCH_FAST(TRV, CWN_ACT_ID)       \
  else if (a_in == CWN_ACT_RELU) CWN_LAUNCH_FAST(TRV, CWN_ACT_RELU) \
  else CWN_LAUNCH_FAST(TRV, kActRuntime)
    if (tr == 64) { CWN_FAST_BY_ACT(64) } else { CWN_FAST_BY_ACT(32) }
#undef CWN_FAST_BY_ACT
#undef CWN_LAUNCH_FAST
    return launched("linear_fwd_fast_kernel");
  }
#define CWN_LAUNCH_FWD(TRV, AV)                                                                             \
  {                                                                                                         \
    if ((rc = ensure_smem(linear_fwd_kernel<TRV, AV>, smem, "cudaFuncSetAttribute(linear_fwd_kernel)"))) return rc; \
    linear_fwd_kernel<TRV, AV><<<total, DT, smem, (cudaStream_t)stream>>>(g);                               \
  }
#define CWN_FWD_BY_ACT(TRV)                                    \
  if (a_in == CWN_ACT_ID) CWN_LAUNCH_FWD(TRV, CWN_ACT_ID)      \
  else if (a_in == CWN_ACT_RELU) CWN_LAUNCH_FWD(TRV, CWN_ACT_RELU) \
  else CWN_LAUNCH_FWD(TRV, kActRuntime)
  if (tr == 64) { CWN_FWD_BY_ACT(64) } else { CWN_FWD_BY_ACT(32) }
#undef CWN_FWD_BY_ACT
#undef CWN_LAUNCH_FWD
  return launched("linear_fwd_kernel");
}

extern "C" int cwn_bn_finalize_grouped(const cwn_bn_desc* descs, int32_t n, cwn_stream_t stream) {
  int rc = check_group(descs, n, "cwn_bn_finalize_grouped");
  if (rc || n == 0) return rc;
  Group<cwn_bn_desc> g;
  g.n = n;
  for (int i = 0; i < n; ++i) {
    const cwn_bn_desc& d = descs[i];
    if (d.h <= 0 || d.h > DT || d.n_rows < 0) return fail(CWN_E_SHAPE, "cwn_bn_finalize_grouped: bad shape (h <= 256)");
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
  bool vec = true;
  for (int i = 0; i < n; ++i) {
    const cwn_bn_act_desc& d = descs[i];
    if (d.h <= 0 || d.n_rows < 0) return fail(CWN_E_SHAPE, "cwn_bn_act_grouped: bad shape");
    if (d.n_rows > 0 && (!d.z || !d.out)) return fail(CWN_E_NULL, "cwn_bn_act_grouped: operand");
    vec = vec && d.h % 4 == 0 && d.ld_z % 4 == 0 && d.ld_out % 4 == 0 && aligned16(d.z) && aligned16(d.out) &&
          (!d.scale || (aligned16(d.mean) && aligned16(d.scale) && (!d.beta || aligned16(d.beta))));
  }
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_bn_act_desc& d = descs[i];
    g.d[i] = d;
    g.start[i] = total;
    total += vec ? (int)((d.n_rows * (d.h / 4) + kBnActV4 - 1) / kBnActV4) : (int)((d.n_rows + TM - 1) / TM);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  const int act = group_act(descs, n, [](const cwn_bn_act_desc& d) { return d.act; });
  cudaError_t err;
  if (act == CWN_ACT_ID) err = launch_pdl(bn_act_kernel<CWN_ACT_ID>, total, DT, 0, (cudaStream_t)stream, g, (int)vec);
  else if (act == CWN_ACT_RELU) err = launch_pdl(bn_act_kernel<CWN_ACT_RELU>, total, DT, 0, (cudaStream_t)stream, g, (int)vec);
  else err = launch_pdl(bn_act_kernel<kActRuntime>, total, DT, 0, (cudaStream_t)stream, g, (int)vec);
  if ((rc = cuda_status(err, "bn_act_kernel"))) return rc;
  return launched("bn_act_kernel");
}

static int load_bwd_group(const cwn_unit_bwd_desc* descs, int n, Group<cwn_unit_bwd_desc>& g, const char* what) {
  int rc = check_group(descs, n, what);
  if (rc) return rc;
  g.n = n;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = descs[i];
    if (d.n_rows < 0 || d.h <= 0 || d.h > DT || d.k0 <= 0 || d.k1 < 0 || d.n_ctas < 0) return fail(CWN_E_SHAPE, what);
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
  int tr;
  if ((rc = group_tile_rows(descs, n, tr))) return rc;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    g.start[i] = total;
    if (g.d[i].has_bn) total += (int)((g.d[i].n_rows + tr - 1) / tr);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  const int a_out = group_act(descs, n, [](const cwn_unit_bwd_desc& d) { return d.act; });
  bool fast = !(g_force_generic_dense & 2);
  for (int i = 0; i < n && fast; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    if (!d.has_bn || d.n_rows == 0) continue;
    fast = d.h % 4 == 0 && aligned16(d.z) && d.ld_z % 4 == 0 && aligned16(d.g_out) && d.ld_g % 4 == 0 &&
           aligned16(d.mean) && aligned16(d.scale) && aligned16(d.rstd) && (!d.beta || aligned16(d.beta)) &&
           aligned16(d.red_partials);
  }
  if (fast) {
#define CWN_REDF(TRV)                                                                                              \
  {                                                                                                                \
    cudaError_t err;                                                                                               \
    if (a_out == CWN_ACT_ID) err = launch_pdl(unit_bwd_reduce_fast_kernel<TRV, CWN_ACT_ID>, total, DT, 0, (cudaStream_t)stream, g); \
    else if (a_out == CWN_ACT_RELU) err = launch_pdl(unit_bwd_reduce_fast_kernel<TRV, CWN_ACT_RELU>, total, DT, 0, (cudaStream_t)stream, g); \
    else err = launch_pdl(unit_bwd_reduce_fast_kernel<TRV, kActRuntime>, total, DT, 0, (cudaStream_t)stream, g);   \
    if ((rc = cuda_status(err, "unit_bwd_reduce_fast_kernel"))) return rc;                                         \
  }
    if (tr == 64) { CWN_REDF(64) } else { CWN_REDF(32) }
#undef CWN_REDF
    return launched("unit_bwd_reduce_fast_kernel");
  }
#define CWN_RED(TRV)                                                                                          \
  if (a_out == CWN_ACT_ID) unit_bwd_reduce_kernel<TRV, CWN_ACT_ID><<<total, DT, 0, (cudaStream_t)stream>>>(g);    \
  else if (a_out == CWN_ACT_RELU) unit_bwd_reduce_kernel<TRV, CWN_ACT_RELU><<<total, DT, 0, (cudaStream_t)stream>>>(g); \
  else unit_bwd_reduce_kernel<TRV, kActRuntime><<<total, DT, 0, (cudaStream_t)stream>>>(g);
  if (tr == 64) { CWN_RED(64) } else { CWN_RED(32) }
#undef CWN_RED
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

// 1 if cwn_unit_bwd_grouped will serve this group with the tensor-core kernel (the only one that can take the upstream
// units' BatchNorm-backward reductions from its g_in tiles), else 0
extern "C" int cwn_unit_bwd_fuses_reduce(const cwn_unit_bwd_desc* descs, int32_t n) {
  if (!descs || n <= 0 || n > CWN_MAX_GROUP) return 0;
  int tr;
  if (group_tile_rows(descs, n, tr) || tr != 64 || !t5_enabled() || (g_force_generic_dense & 2)) return 0;
  size_t smem5 = 0;
  for (int i = 0; i < n; ++i)
    if (descs[i].n_rows != 0 && !t5_bwd_ok(descs[i], smem5)) return 0;
  return smem5 > 0 ? 1 : 0;
}

extern "C" int cwn_unit_bwd_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_unit_bwd_grouped");
  if (rc || n == 0) return rc;
  int tr;
  if ((rc = group_tile_rows(descs, n, tr))) return rc;
  int total = 0;
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    const int n_tiles = (int)((d.n_rows + tr - 1) / tr);
    if (d.n_ctas > n_tiles || (n_tiles > 0 && d.n_ctas == 0)) return fail(CWN_E_SHAPE, "cwn_unit_bwd_grouped: n_ctas must be in [1, row tiles]");
    if (d.n_rows > 0 && (!d.w_partials || !d.b_partials)) return fail(CWN_E_NULL, "cwn_unit_bwd_grouped: partial buffers");
    g.start[i] = total;
    total += d.n_ctas;
    const int H4 = round4(d.h), m_tiles = (d.h + 63) / 64, K4 = round4(d.k0 + d.k1);
    const size_t need = ((size_t)tr * (H4 + 4) + (size_t)m_tiles * 64 * (tr + 4) + (size_t)tr * LDT + (size_t)H4 * LDT +
                         3 * (size_t)K4 + 6 * (size_t)H4) * sizeof(float);
    if (need > smem) smem = need;
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  const int a_in = group_act(descs, n, [](const cwn_unit_bwd_desc& d) { return d.in_act; });
  const int a_out = group_act(descs, n, [](const cwn_unit_bwd_desc& d) { return d.act; });
  bool wants_fused_reduce = false;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    for (int hf = 0; hf < 2; ++hf) {
      if (!(hf ? d.next_red1 : d.next_red0)) continue;
      wants_fused_reduce = true;
      const bool ok = (hf ? d.g_in1 : d.g_in0) && !d.accumulate_in && (hf ? d.next_rstd1 : d.next_rstd0) &&
                      (hf ? d.next_c1_1 : d.next_c1_0) && (hf ? d.next_c2_1 : d.next_c2_0) && d.next_counter &&
                      (hf ? d.in_mean1 : d.in_mean0) && (hf ? d.in_scale1 : d.in_scale0) &&
                      aligned16(hf ? d.next_rstd1 : d.next_rstd0) && (hf ? d.k1 : d.k0) > 0;
      if (!ok) return fail(CWN_E_NULL, "cwn_unit_bwd_grouped: next_red needs g_in, accumulate_in == 0, next_rstd / c1 / c2 / counter and the input transform");
    }
  }
  if (tr == 64 && t5_enabled() && !(g_force_generic_dense & 2)) {  // tensor-core path (see dense_tc5.cuh)
    bool ok = true;
    size_t smem5 = 0;
    for (int i = 0; i < n && ok; ++i) ok = g.d[i].n_rows == 0 || t5_bwd_ok(g.d[i], smem5);
    if (ok && smem5 > 0) {
#define CWN_LAUNCH_T5B(AI, AO)                                                                                            \
  {                                                                                                                       \
    if ((rc = ensure_smem(unit_bwd_tc5_kernel<AI, AO>, smem5, "cudaFuncSetAttribute(unit_bwd_tc5_kernel)"))) return rc;   \
    if ((rc = cuda_status(launch_pdl(unit_bwd_tc5_kernel<AI, AO>, total, T5T, smem5, (cudaStream_t)stream, g), "unit_bwd_tc5_kernel"))) return rc; \
  }
#define CWN_T5B_BY_OUT(AI)                                        \
  if (a_out == CWN_ACT_ID) CWN_LAUNCH_T5B(AI, CWN_ACT_ID)         \
  else if (a_out == CWN_ACT_RELU) CWN_LAUNCH_T5B(AI, CWN_ACT_RELU) \
  else CWN_LAUNCH_T5B(AI, kActRuntime)
      if (a_in == CWN_ACT_ID) { CWN_T5B_BY_OUT(CWN_ACT_ID) }
      else if (a_in == CWN_ACT_RELU) { CWN_T5B_BY_OUT(CWN_ACT_RELU) }
      else { CWN_T5B_BY_OUT(kActRuntime) }
#undef CWN_T5B_BY_OUT
#undef CWN_LAUNCH_T5B
      return launched("unit_bwd_tc5_kernel");
    }
  }
  if (wants_fused_reduce) return fail(CWN_E_SHAPE, "cwn_unit_bwd_grouped: next_red (fused upstream reduction) needs the tensor-core path: ask cwn_unit_bwd_fuses_reduce first");
  for (int i = 0; i < n; ++i)
    if (g.d[i].n_rows_live) return fail(CWN_E_SHAPE, "cwn_unit_bwd_grouped: n_rows_live needs the tensor-core path (h = 64, K in {32, 64, 128}, tile_rows 64)");
  bool fast = !(g_force_generic_dense & 2);
  size_t smem_fast = 0;
  for (int i = 0; i < n && fast; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    if (d.n_rows == 0) continue;
    auto vec_ok = [](const float* a, const float* b, const float* c) { return aligned16(a) && aligned16(b) && aligned16(c); };
    const int K = d.k0 + d.k1;
    fast = d.k0 % 4 == 0 && d.k1 % 4 == 0 && d.h % 4 == 0 && aligned16(d.x0) && d.ld_x0 % 4 == 0 &&
           (d.k1 == 0 || (aligned16(d.x1) && d.ld_x1 % 4 == 0)) && aligned16(d.w) && d.ld_w % 4 == 0 &&
           aligned16(d.z) && d.ld_z % 4 == 0 && aligned16(d.g_out) && d.ld_g % 4 == 0 &&
           (!d.in_scale0 || vec_ok(d.in_mean0, d.in_scale0, d.in_beta0)) &&
           (!d.in_scale1 || vec_ok(d.in_mean1, d.in_scale1, d.in_beta1)) &&
           (!d.g_in0 || (aligned16(d.g_in0) && d.ld_gi0 % 4 == 0)) && (!d.g_in1 || (aligned16(d.g_in1) && d.ld_gi1 % 4 == 0)) &&
           aligned16(d.w_partials);
    const int m_tiles = (d.h + 63) / 64;
    const size_t zt = std::max((size_t)tr * (d.h + 4), (size_t)m_tiles * 64 * (tr + 4));
    // (+64: the 64-column chunks of the two products may read past column K of the last row; those lanes' results
    // are never stored, but the reads must stay inside the allocation)
    const size_t need = ((size_t)tr * (d.h + 4) + zt + (size_t)tr * (K + 4) + (size_t)d.h * (K + 4) + 3 * (size_t)K +
                         6 * (size_t)d.h + 64) * sizeof(float);
    if (need > smem_fast) smem_fast = need;
  }
  if (fast && smem_fast <= 210 * 1024) {
#define CWN_LAUNCH_BWDF(TRV, AI, AO)                                                                              \
  {                                                                                                               \
    if ((rc = ensure_smem(unit_bwd_fast_kernel<TRV, AI, AO>, smem_fast, "cudaFuncSetAttribute(unit_bwd_fast_kernel)"))) return rc; \
    unit_bwd_fast_kernel<TRV, AI, AO><<<total, DT, smem_fast, (cudaStream_t)stream>>>(g);                         \
  }
#define CWN_BWDF_BY_OUT(TRV, AI)                                     \
  if (a_out == CWN_ACT_ID) CWN_LAUNCH_BWDF(TRV, AI, CWN_ACT_ID)      \
  else if (a_out == CWN_ACT_RELU) CWN_LAUNCH_BWDF(TRV, AI, CWN_ACT_RELU) \
  else CWN_LAUNCH_BWDF(TRV, AI, kActRuntime)
#define CWN_BWDF_BY_IN(TRV)                                      \
  if (a_in == CWN_ACT_ID) { CWN_BWDF_BY_OUT(TRV, CWN_ACT_ID) }   \
  else if (a_in == CWN_ACT_RELU) { CWN_BWDF_BY_OUT(TRV, CWN_ACT_RELU) } \
  else { CWN_BWDF_BY_OUT(TRV, kActRuntime) }
    if (tr == 64) { CWN_BWDF_BY_IN(64) } else { CWN_BWDF_BY_IN(32) }
#undef CWN_BWDF_BY_IN
#undef CWN_BWDF_BY_OUT
#undef CWN_LAUNCH_BWDF
    return launched("unit_bwd_fast_kernel");
  }
#define CWN_LAUNCH_BWD(TRV, AI, AO)                                                                          \
  {                                                                                                          \
    if ((rc = ensure_smem(unit_bwd_kernel<TRV, AI, AO>, smem, "cudaFuncSetAttribute(unit_bwd_kernel)"))) return rc; \
    unit_bwd_kernel<TRV, AI, AO><<<total, DT, smem, (cudaStream_t)stream>>>(g);                              \
  }
#define CWN_BWD_BY_OUT(TRV, AI)                                     \
  if (a_out == CWN_ACT_ID) CWN_LAUNCH_BWD(TRV, AI, CWN_ACT_ID)      \
  else if (a_out == CWN_ACT_RELU) CWN_LAUNCH_BWD(TRV, AI, CWN_ACT_RELU) \
  else CWN_LAUNCH_BWD(TRV, AI, kActRuntime)
#define CWN_BWD_BY_IN(TRV)                                      \
  if (a_in == CWN_ACT_ID) { CWN_BWD_BY_OUT(TRV, CWN_ACT_ID) }   \
  else if (a_in == CWN_ACT_RELU) { CWN_BWD_BY_OUT(TRV, CWN_ACT_RELU) } \
  else { CWN_BWD_BY_OUT(TRV, kActRuntime) }
  if (tr == 64) { CWN_BWD_BY_IN(64) } else { CWN_BWD_BY_IN(32) }
#undef CWN_BWD_BY_IN
#undef CWN_BWD_BY_OUT
#undef CWN_LAUNCH_BWD
  return launched("unit_bwd_kernel");
}

extern "C" int cwn_wgrad_finalize_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream) {
  Group<cwn_unit_bwd_desc> g;
  int rc = load_bwd_group(descs, n, g, "cwn_wgrad_finalize_grouped");
  if (rc || n == 0) return rc;
  bool vec = true;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    if (!d.g_w) return fail(CWN_E_NULL, "cwn_wgrad_finalize_grouped: g_w");
    vec = vec && (d.k0 + d.k1) % 4 == 0 && d.h % 4 == 0 && aligned16(d.w_partials) && aligned16(d.b_partials);
  }
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const cwn_unit_bwd_desc& d = g.d[i];
    g.start[i] = total;
    if (d.n_ctas <= 0) continue;
    const int64_t elems = (int64_t)d.h * (d.k0 + d.k1) + d.h;
    total += vec ? (int)((elems / 4 + 63) / 64) : (int)((elems + DT - 1) / DT);
  }
  g.start[n] = total;
  if (total == 0) return CWN_OK;
  if ((rc = cuda_status(launch_pdl(wgrad_finalize_kernel, total, DT, 0, (cudaStream_t)stream, g, (int)vec), "wgrad_finalize_kernel"))) return rc;
  return launched("wgrad_finalize_kernel");
}
