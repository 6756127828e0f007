// CSR plan builder: stable grouping of the messages of one adjacency by one of their columns.
//
// The reference scatters with unsorted destinations and atomics (torch_scatter.scatter -> scatter_add_,
// mp/cell_mp.py:439-440) every layer, forward and backward. Index tensors are constant across layers and between
// forward and backward, so the B200 design sorts ONCE per batch and every later pass is an
// atomic-free, deterministic segmented reduction.
//
// Pipeline (all stream-ordered, no host sync, caller-provided workspace):
//   1. narrow_keys   : int64 key -> int32 key (out-of-range keys go to the dummy row n_rows and raise flag bit 0),
//                      value = message id
//   2. cub::DeviceRadixSort::SortPairs over ceil(log2(n_rows+1)) bits (LSD radix sort => stable)
//   3. finish_plan   : rowptr[r] = lower_bound(sorted keys, r); payload columns permuted + narrowed to int32
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace cwn {

thread_local char g_err[256] = "";
std::atomic<unsigned long long> g_launches{0};

__global__ void narrow_keys_kernel(const int64_t* __restrict__ key, int64_t E, int64_t n_rows,
                                   int32_t* __restrict__ key32, int32_t* __restrict__ val, int32_t* flags) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = key[e];
    if (k < 0 || k >= n_rows) {
      k = n_rows;
      if (flags) atomicOr(flags, 1);
    }
    key32[e] = (int32_t)k;
    val[e] = (int32_t)e;
  }
}

__global__ void finish_plan_kernel(const int32_t* __restrict__ key_sorted, const int32_t* __restrict__ perm,
                                   int64_t E, int64_t n_rows, const int64_t* __restrict__ pay0,
                                   const int64_t* __restrict__ pay1, int32_t* __restrict__ rowptr,
                                   int32_t* __restrict__ pay0_sorted, int32_t* __restrict__ pay1_sorted) {
  const int64_t total = (E > n_rows + 1) ? E : n_rows + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i <= n_rows) {  // first position whose key is >= i
      int64_t lo = 0, hi = E;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (key_sorted[mid] < (int32_t)i) lo = mid + 1; else hi = mid;
      }
      rowptr[i] = (int32_t)lo;
    }
    if (i < E) {
      const int32_t e = perm[i];
      if (pay0) pay0_sorted[i] = (int32_t)pay0[e];
      if (pay1) pay1_sorted[i] = (int32_t)pay1[e];
    }
  }
}

__global__ void fill_zero_kernel(int32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

// ---- all the (small) plans of one batch in ONE launch: one CTA per plan, block-wide stable radix sort in shared
// memory. At the reference's real-data shape (128 molecules: E <= 10 240 per adjacency) this replaces ~6 launches
// per plan (13 plans per batch) by a single kernel.
constexpr int kSmallThreads = 512;
constexpr int kSmallItemsMax = 24;
constexpr int kSmallItemsMin = 8;  // second instantiation for batches whose plans all have E <= 4096
constexpr int kSmallCapacity = kSmallThreads * kSmallItemsMax;  // 12 288 messages per plan
constexpr int kMaxPlansPerLaunch = 16;

struct PlanBatch {
  cwn_plan_desc d[kMaxPlansPerLaunch];
  int bits[kMaxPlansPerLaunch];
};

template <int kSmallItems>
__global__ void __launch_bounds__(kSmallThreads)
small_plans_kernel(const __grid_constant__ PlanBatch batch, int32_t* flags) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using Sort = cub::BlockRadixSort<int32_t, kSmallThreads, kSmallItems, int32_t>;
  extern __shared__ __align__(16) unsigned char smem[];
  typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(smem);
  int32_t* sorted = reinterpret_cast<int32_t*>(smem + ((sizeof(typename Sort::TempStorage) + 15) & ~size_t(15)));
  const cwn_plan_desc& p = batch.d[blockIdx.x];
  const int E = (int)p.E;
  const int n_rows = (int)p.n_rows;
  const int t = threadIdx.x;
  int32_t keys[kSmallItems], vals[kSmallItems];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kSmallItems; ++i) {
    const int e = t * kSmallItems + i;  // blocked arrangement: ascending message id => stability is meaningful
    int32_t k = n_rows;                 // padding and out-of-range keys sort behind every real row
    if (e < E) {
      const int64_t kk = p.key[e];
      if (kk >= 0 && kk < n_rows) k = (int32_t)kk; else bad = true;
    }
    keys[i] = k;
    vals[i] = e;
  }
  if (bad && flags) atomicOr(flags, 1);
  Sort(temp).Sort(keys, vals, 0, batch.bits[blockIdx.x]);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSmallItems; ++i) {
    const int pos = t * kSmallItems + i;
    sorted[pos] = keys[i];
    if (pos < E) {
      const int32_t e = vals[i];
      p.perm[pos] = e;
      if (p.pay0) p.pay0_sorted[pos] = (int32_t)p.pay0[e];
      if (p.pay1) p.pay1_sorted[pos] = (int32_t)p.pay1[e];
    }
  }
  __syncthreads();
  for (int r = t; r <= n_rows; r += kSmallThreads) {  // rowptr[r] = first position whose key is >= r
    int lo = 0, hi = E;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sorted[mid] < r) lo = mid + 1; else hi = mid;
    }
    p.rowptr[r] = lo;
  }
}

static int key_bits(int64_t n_rows) {  // keys take values 0..n_rows (n_rows = dummy row)
  int bits = 1;
  while ((int64_t(1) << bits) <= n_rows) ++bits;
  return bits;
}

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

static size_t cub_temp_bytes(int64_t E, int64_t n_rows) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E, 0, key_bits(n_rows));
  return bytes;
}

static int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace cwn

using namespace cwn;

extern "C" const char* cwn_version(void) { return "cwn_b200 0.1 (sm_100a)"; }
extern "C" const char* cwn_last_error_string(void) { return g_err; }
extern "C" unsigned long long cwn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t cwn_csr_plan_workspace_bytes(int64_t E, int64_t n_rows) {
  if (E <= 0 || n_rows < 0 || E > INT32_MAX) return 256;
  return 3 * align_up((size_t)E * sizeof(int32_t)) + align_up(cub_temp_bytes(E, n_rows)) + 256;
}

extern "C" int cwn_csr_plan_build(const int64_t* key, const int64_t* pay0, const int64_t* pay1, int64_t E,
                                  int64_t n_rows, int32_t* rowptr, int32_t* perm, int32_t* pay0_sorted,
                                  int32_t* pay1_sorted, int32_t* flags, void* workspace, size_t workspace_bytes,
                                  cwn_stream_t stream) {
  if (E < 0 || n_rows < 0 || E > INT32_MAX || n_rows >= INT32_MAX) return fail(CWN_E_SHAPE, "cwn_csr_plan_build: bad E/n_rows");
  if (!rowptr) return fail(CWN_E_NULL, "rowptr");
  cudaStream_t st = (cudaStream_t)stream;
  if (E == 0) {
    fill_zero_kernel<<<blocks_for(n_rows + 1), 256, 0, st>>>(rowptr, n_rows + 1);
    return launched("cwn_csr_plan_build(empty)");
  }
  if (!key || !perm) return fail(CWN_E_NULL, "key/perm");
  if ((pay0 && !pay0_sorted) || (pay1 && !pay1_sorted)) return fail(CWN_E_NULL, "payload output");
  if (!workspace || workspace_bytes < cwn_csr_plan_workspace_bytes(E, n_rows)) return fail(CWN_E_WORKSPACE, "workspace too small");
  char* ws = (char*)workspace;
  const size_t col = align_up((size_t)E * sizeof(int32_t));
  int32_t* key32 = (int32_t*)ws;
  int32_t* val = (int32_t*)(ws + col);
  int32_t* key_sorted = (int32_t*)(ws + 2 * col);
  void* cub_temp = ws + 3 * col;
  size_t cub_bytes = workspace_bytes - 3 * col;

  narrow_keys_kernel<<<blocks_for(E), 256, 0, st>>>(key, E, n_rows, key32, val, flags);
  int rc = launched("narrow_keys");
  if (rc) return rc;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, key32, key_sorted, val, perm, (int)E, 0,
                                                   key_bits(n_rows), st);
  g_launches.fetch_add(1 + (key_bits(n_rows) + 7) / 8, std::memory_order_relaxed);  // histogram + onesweep passes
  if ((rc = cuda_status(ce, "cub::DeviceRadixSort::SortPairs"))) return rc;
  finish_plan_kernel<<<blocks_for(E > n_rows + 1 ? E : n_rows + 1), 256, 0, st>>>(
      key_sorted, perm, E, n_rows, pay0, pay1, rowptr, pay0_sorted, pay1_sorted);
  return launched("finish_plan");
}

extern "C" int64_t cwn_csr_plan_small_capacity(void) { return kSmallCapacity; }

extern "C" int cwn_csr_plan_build_small(const cwn_plan_desc* descs, int32_t n_plans, int32_t* flags,
                                        cwn_stream_t stream) {
  if (n_plans < 0) return fail(CWN_E_SHAPE, "cwn_csr_plan_build_small: negative plan count");
  if (n_plans == 0) return CWN_OK;
  if (!descs) return fail(CWN_E_NULL, "descs");
  using SortMax = cub::BlockRadixSort<int32_t, kSmallThreads, kSmallItemsMax, int32_t>;
  using SortMin = cub::BlockRadixSort<int32_t, kSmallThreads, kSmallItemsMin, int32_t>;
  const size_t smem_max = ((sizeof(typename SortMax::TempStorage) + 15) & ~size_t(15)) + (size_t)kSmallThreads * kSmallItemsMax * sizeof(int32_t);
  const size_t smem_min = ((sizeof(typename SortMin::TempStorage) + 15) & ~size_t(15)) + (size_t)kSmallThreads * kSmallItemsMin * sizeof(int32_t);
  static std::atomic<bool> configured{false};
  if (!configured.load(std::memory_order_acquire)) {
    cudaError_t ce = cudaFuncSetAttribute(small_plans_kernel<kSmallItemsMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(small_plans_kernel<kSmallItemsMin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_min);
    if (ce != cudaSuccess) return cuda_status(ce, "cudaFuncSetAttribute(small_plans_kernel)");
    configured.store(true, std::memory_order_release);
  }
  for (int base = 0; base < n_plans; base += kMaxPlansPerLaunch) {
    PlanBatch batch;
    const int n = (n_plans - base < kMaxPlansPerLaunch) ? n_plans - base : kMaxPlansPerLaunch;
    for (int i = 0; i < n; ++i) {
      const cwn_plan_desc& d = descs[base + i];
      if (d.E < 0 || d.n_rows < 0 || d.E > kSmallCapacity || d.n_rows >= INT32_MAX)
        return fail(CWN_E_SHAPE, "cwn_csr_plan_build_small: plan exceeds cwn_csr_plan_small_capacity()");
      if (!d.rowptr || (d.E > 0 && (!d.key || !d.perm))) return fail(CWN_E_NULL, "plan descriptor");
      if ((d.pay0 && !d.pay0_sorted) || (d.pay1 && !d.pay1_sorted)) return fail(CWN_E_NULL, "payload output");
      batch.d[i] = d;
      batch.bits[i] = key_bits(d.n_rows);
    }
    int64_t max_e = 0;
    for (int i = 0; i < n; ++i) max_e = batch.d[i].E > max_e ? batch.d[i].E : max_e;
    if (max_e <= kSmallThreads * kSmallItemsMin)
      launch_pdl(small_plans_kernel<kSmallItemsMin>, n, kSmallThreads, smem_min, (cudaStream_t)stream, batch, flags);
    else
      launch_pdl(small_plans_kernel<kSmallItemsMax>, n, kSmallThreads, smem_max, (cudaStream_t)stream, batch, flags);
    int rc = launched("small_plans_kernel");
    if (rc) return rc;
  }
  return CWN_OK;
}
