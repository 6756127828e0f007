// GPU-side collation: build the packed tensors of a ComplexBatch directly in HBM from a device-resident dataset.
//
// The reference collates on the CPU in Python (ComplexBatch.from_complex_list -> CochainBatch.from_cochain_list,
// data/complex.py:323-458, 690-728; data/data_loading.py:44-82): per key a loop over the complexes that adds the
// running cell-count offsets (`__inc__`, :148-169) and concatenates. With the training step at ~1.5 ms that loop
// (milliseconds per batch) would cap end-to-end throughput. Here the dataset lives in HBM as flat arrays with
// per-complex segment pointers; a batch is a list of complex ids; every output tensor is a concatenation of segments,
// each shifted by a per-segment constant — ONE kernel does all tensors of a batch ("jobs"), writing straight into the
// packed batch buffers (which may be the static buffers of the captured CUDA graph).
#include "common.cuh"

namespace cwn {

constexpr int kMaxJobs = 24;

struct CollateBatch {
  cwn_collate_job job[kMaxJobs];
  int64_t start[kMaxJobs + 1];  // first output element (row) of each job in the flattened work list
  int n;
};

// binary search: segment j with dst_start[j] <= pos < dst_start[j+1]
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ dst_start, int n_seg, int64_t pos) {
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(dst_start + mid) <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) collate_kernel(const __grid_constant__ CollateBatch cb) {
  const int64_t total = cb.start[cb.n];
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    int j = 0;
#pragma unroll
    for (int i = 1; i < kMaxJobs; ++i)
      if (i < cb.n && w >= cb.start[i]) j = i;
    const cwn_collate_job& job = cb.job[j];
    const int64_t pos = w - cb.start[j];                 // output row of this job
    const int seg = find_segment(job.dst_start, job.n_segments, pos);
    const int64_t t = pos - __ldg(job.dst_start + seg);  // position inside the segment
    if (job.kind == CWN_COLLATE_FILL) {                  // batch vector: the segment number
      reinterpret_cast<int64_t*>(job.dst)[pos] = seg;
      continue;
    }
    const int64_t s = __ldg(job.src_start + seg) + t;
    if (job.kind == CWN_COLLATE_I64) {
      const int64_t add = job.add ? __ldg(job.add + seg) : 0;
      reinterpret_cast<int64_t*>(job.dst)[pos] = __ldg(reinterpret_cast<const int64_t*>(job.src) + s) + add;
    } else {  // rows of `row_elems` 4-byte elements (features), copied as they are
      const float* src = reinterpret_cast<const float*>(job.src) + s * job.row_elems;
      float* dst = reinterpret_cast<float*>(job.dst) + pos * job.row_elems;
      for (int c = 0; c < job.row_elems; ++c) dst[c] = __ldg(src + c);
    }
  }
}

}  // namespace cwn

using namespace cwn;

extern "C" int cwn_collate(const cwn_collate_job* jobs, int32_t n_jobs, cwn_stream_t stream) {
  if (n_jobs < 0) return fail(CWN_E_SHAPE, "cwn_collate: negative job count");
  if (n_jobs == 0) return CWN_OK;
  if (!jobs) return fail(CWN_E_NULL, "cwn_collate: jobs");
  for (int base = 0; base < n_jobs; base += kMaxJobs) {
    CollateBatch cb;
    cb.n = (n_jobs - base < kMaxJobs) ? n_jobs - base : kMaxJobs;
    int64_t total = 0;
    for (int i = 0; i < cb.n; ++i) {
      const cwn_collate_job& j = jobs[base + i];
      if (j.n_segments < 0 || j.n_out < 0 || j.row_elems < 0) return fail(CWN_E_SHAPE, "cwn_collate: bad job");
      if (j.kind < CWN_COLLATE_I64 || j.kind > CWN_COLLATE_FILL) return fail(CWN_E_ENUM, "cwn_collate: kind");
      if (j.n_out > 0 && (!j.dst || !j.dst_start || (j.kind != CWN_COLLATE_FILL && (!j.src || !j.src_start))))
        return fail(CWN_E_NULL, "cwn_collate: job pointers");
      cb.job[i] = j;
      cb.start[i] = total;
      total += (j.n_segments > 0) ? j.n_out : 0;
    }
    cb.start[cb.n] = total;
    if (total == 0) continue;
    int64_t blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    collate_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(cb);
    int rc = launched("collate_kernel");
    if (rc) return rc;
  }
  return CWN_OK;
}
