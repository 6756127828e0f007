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
#include <cub/block/block_scan.cuh>
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

// ---- the same plans by COUNTING instead of sorting (default for the plans of a real batch). The radix kernel above moves
// 12 288 padded (key, value) pairs through four ranking + exchange passes whatever E is: 49 us at the benchmark's batch —
// the FIRST kernel of every step, fully on its critical path. A stable grouping needs less: histogram the keys
// (shared-memory atomics), scan the histogram into rowptr, scatter every message id to a slot of its row in ARBITRARY
// order (atomic cursor), then sort each row's few ids ascending — which makes the result the unique stable grouping,
// identical to the radix sort's, however the atomics interleaved. Rows of <= kInsertMax ids are sorted by one thread
// (insertion sort in shared memory), longer rows by a warp (rank by counting, L^2 / 32 steps per lane). Meant for short
// rows: the host sends plans with E / n_rows > kCountMaxAvgRow, or whose tables do not fit shared memory, to the radix
// kernel. ~8 us for the 13 plans of a 128-molecule batch.
constexpr int kCountThreads = 1024;
constexpr int kInsertMax = 32;
constexpr int kCountMaxAvgRow = 64;
constexpr size_t kCountSmemMax = 200 * 1024;

static size_t count_smem_bytes(int64_t E, int64_t n_rows) { return (size_t)(2 * (n_rows + 2) + E) * sizeof(int32_t); }
static bool count_path_ok(const cwn_plan_desc& d) {
  return d.E > 0 && d.E <= kCountMaxAvgRow * (d.n_rows > 0 ? d.n_rows : 1) && count_smem_bytes(d.E, d.n_rows) <= kCountSmemMax;
}

__global__ void __launch_bounds__(kCountThreads)
small_plans_count_kernel(const __grid_constant__ PlanBatch batch, int32_t* flags) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  using Scan = cub::BlockScan<int32_t, kCountThreads>;
  __shared__ typename Scan::TempStorage scan_temp;
  extern __shared__ __align__(16) unsigned char smem[];
  const cwn_plan_desc& p = batch.d[blockIdx.x];
  const int E = (int)p.E, n_rows = (int)p.n_rows, t = threadIdx.x;
  const int n_keys = n_rows + 1;                      // row n_rows collects out-of-range keys (behind every real row)
  int32_t* start = reinterpret_cast<int32_t*>(smem);  // [n_keys + 1]: counts, then their exclusive prefix (= rowptr)
  int32_t* cur = start + (n_keys + 1);                // [n_keys]: next free slot of every row
  int32_t* perm_s = cur + (n_keys + 1);               // [E]
  for (int r = t; r <= n_keys; r += kCountThreads) start[r] = 0;
  __syncthreads();
  bool bad = false;
  for (int e = t; e < E; e += kCountThreads) {
    const int64_t kk = p.key[e];
    const bool ok = kk >= 0 && kk < n_rows;
    bad |= !ok;
    atomicAdd(&start[ok ? (int)kk : n_rows], 1);
  }
  if (bad && flags) atomicOr(flags, 1);
  __syncthreads();
  {  // exclusive prefix of the n_keys counts; thread t owns a contiguous segment
    const int seg = (n_keys + kCountThreads - 1) / kCountThreads;
    const int lo = min(t * seg, n_keys), hi = min(lo + seg, n_keys);
    int32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += start[i];
    int32_t base;
    Scan(scan_temp).ExclusiveSum(sum, base);
    for (int i = lo; i < hi; ++i) {
      const int32_t c = start[i];
      start[i] = base;
      cur[i] = base;
      base += c;
    }
    if (t == kCountThreads - 1) start[n_keys] = E;
  }
  __syncthreads();
  for (int r = t; r <= n_rows; r += kCountThreads) p.rowptr[r] = start[r];
  for (int e = t; e < E; e += kCountThreads) {
    const int64_t kk = p.key[e];
    const int k = (kk >= 0 && kk < n_rows) ? (int)kk : n_rows;
    perm_s[atomicAdd(&cur[k], 1)] = e;
  }
  __syncthreads();
  // ascending ids inside every row: short rows by one thread ...
  for (int r = t; r < n_keys; r += kCountThreads) {
    const int b = start[r], n = start[r + 1] - b;
    if (n < 2 || n > kInsertMax) continue;
    for (int i = 1; i < n; ++i) {
      const int32_t x = perm_s[b + i];
      int j = i - 1;
      while (j >= 0 && perm_s[b + j] > x) {
        perm_s[b + j + 1] = perm_s[b + j];
        --j;
      }
      perm_s[b + j + 1] = x;
    }
  }
  // ... long rows by a warp: the rank of an id is the number of smaller ids in its row (ids are distinct). The ranked
  // ids pass through the output array in global memory and come back, so that the write-out below is uniform.
  {
    const int warp = t >> 5, lane = t & 31;
    for (int r = warp; r < n_keys; r += kCountThreads / 32) {
      const int b = start[r], n = start[r + 1] - b;
      if (n <= kInsertMax) continue;
      for (int i = lane; i < n; i += 32) {
        const int32_t x = perm_s[b + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += perm_s[b + j] < x;
        p.perm[b + rank] = x;
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) perm_s[b + i] = p.perm[b + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int pos = t; pos < E; pos += kCountThreads) {
    const int32_t e = perm_s[pos];
    p.perm[pos] = e;
    if (p.pay0) p.pay0_sorted[pos] = (int32_t)p.pay0[e];
    if (p.pay1) p.pay1_sorted[pos] = (int32_t)p.pay1[e];
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
  for (int i = 0; i < n_plans; ++i) {
    const cwn_plan_desc& d = descs[i];
    if (d.E < 0 || d.n_rows < 0 || d.E > kSmallCapacity || d.n_rows >= INT32_MAX)
      return fail(CWN_E_SHAPE, "cwn_csr_plan_build_small: plan exceeds cwn_csr_plan_small_capacity()");
    if (!d.rowptr || (d.E > 0 && (!d.key || !d.perm))) return fail(CWN_E_NULL, "plan descriptor");
    if ((d.pay0 && !d.pay0_sorted) || (d.pay1 && !d.pay1_sorted)) return fail(CWN_E_NULL, "payload output");
  }
  static const bool count_enabled = [] { const char* v = getenv("CWN_B200_PLAN_COUNT"); return !(v && v[0] == '0'); }();
  static std::atomic<bool> count_configured{false};
  if (count_enabled && !count_configured.load(std::memory_order_acquire)) {
    cudaError_t ce = cudaFuncSetAttribute(small_plans_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCountSmemMax);
    if (ce != cudaSuccess) return cuda_status(ce, "cudaFuncSetAttribute(small_plans_count_kernel)");
    count_configured.store(true, std::memory_order_release);
  }
  // short-row plans (what a batch of complexes produces) go to the counting kernel, the rest to the radix kernel
  for (int pass = 0; pass < 2; ++pass) {
    PlanBatch batch;
    int n = 0;
    int64_t max_e = 0;
    size_t max_smem = 0;
    auto flush = [&]() -> int {
      if (n == 0) return CWN_OK;
      if (pass == 0) {
        launch_pdl(small_plans_count_kernel, n, kCountThreads, max_smem, (cudaStream_t)stream, batch, flags);
      } else if (max_e <= kSmallThreads * kSmallItemsMin) {
        launch_pdl(small_plans_kernel<kSmallItemsMin>, n, kSmallThreads, smem_min, (cudaStream_t)stream, batch, flags);
      } else {
        launch_pdl(small_plans_kernel<kSmallItemsMax>, n, kSmallThreads, smem_max, (cudaStream_t)stream, batch, flags);
      }
      n = 0;
      max_e = 0;
      max_smem = 0;
      return launched("small_plans_kernel");
    };
    for (int i = 0; i < n_plans; ++i) {
      const cwn_plan_desc& d = descs[i];
      const bool by_count = count_enabled && count_path_ok(d);
      if (by_count != (pass == 0)) continue;
      batch.d[n] = d;
      batch.bits[n] = key_bits(d.n_rows);
      max_e = d.E > max_e ? d.E : max_e;
      const size_t sm = count_smem_bytes(d.E, d.n_rows);
      max_smem = sm > max_smem ? sm : max_smem;
      if (++n == kMaxPlansPerLaunch) {
        int rc = flush();
        if (rc) return rc;
      }
    }
    int rc = flush();
    if (rc) return rc;
  }
  return CWN_OK;
}
