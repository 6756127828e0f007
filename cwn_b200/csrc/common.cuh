// Shared helpers of the cwn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <stdlib.h>
#include <utility>
#include "../../include/cwn_b200.h"

namespace cwn {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized as multiples of this

extern thread_local char g_err[256];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

inline int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CWN_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

// count + check a kernel launch (no sync: errors of the launch itself only)
inline int launched(const char* what, unsigned n = 1) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return cuda_status(cudaGetLastError(), what);
}

// ---- programmatic dependent launch (PDL). At the real-data shape the step is a chain of ~70 dependent launches of a
// few microseconds each; between two kernel nodes of a CUDA graph the GPU idles ~1.5-2 us (launch latency + the next
// kernel's prologue). A kernel launched with the programmatic-serialization attribute may start as soon as every CTA of
// its stream predecessor has called pdl_trigger() (or exited): its prologue (descriptor staging, mbarrier / TMEM setup)
// then overlaps the predecessor's tail, and pdl_wait() blocks until the predecessor grid has completed and its writes
// are visible. Contract: a kernel launched through launch_pdl() calls pdl_wait() before its first access to global
// memory (reads AND writes: the predecessor may still be reading what this kernel overwrites). Both calls are no-ops in
// a kernel launched the ordinary way. CWN_B200_PDL=0 launches everything the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("CWN_B200_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// 128-bit read-only load that does not pollute L1 for rows that are consumed once per gather
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_max(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// activations of the coboundary message MLP and their derivatives (as functions of the pre-activation)
template <int ACT>
__device__ __forceinline__ float act_fwd(float v) {
  if (ACT == CWN_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == CWN_ACT_ELU) return v > 0.f ? v : expm1f(v);
  if (ACT == CWN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (ACT == CWN_ACT_TANH) return tanhf(v);
  return v;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float v) {
  if (ACT == CWN_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (ACT == CWN_ACT_ELU) return v > 0.f ? 1.f : expf(v);
  if (ACT == CWN_ACT_SIGMOID) {
    float s = 1.f / (1.f + expf(-v));
    return s * (1.f - s);
  }
  if (ACT == CWN_ACT_TANH) {
    float t = tanhf(v);
    return 1.f - t * t;
  }
  return 1.f;
}

}  // namespace cwn
