// One-launch Adam over the flat parameter / gradient buffers of the data-parallel step.
//
// The reference trains with torch.optim.Adam (exp/run_exp.py:343), which at this model size (~170 small tensors) is a
// dozen multi-tensor launches per step plus a memset of the gradients; with every gradient already living in one flat
// bucket (cwn_b200/dist.py) and the parameters flattened the same way, the update is a single streaming kernel that
// also clears the gradient for the next step. The step counter lives on the device (CUDA-graph friendly) and is
// advanced by the last CTA to finish.
#include "common.cuh"

namespace cwn {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
            float b1, float b2, float eps, float wd, int32_t* step, int32_t* counter, int zero_grad) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int t = *reinterpret_cast<volatile int32_t*>(step) + 1;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
    if (zero_grad) g[i] = 0.f;
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {  // every CTA has read `step` before it bumped the counter
    *step = t;
    *counter = 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Data parallel: gradient all-reduce FUSED with Adam over NVLink peer memory (one launch; no NCCL on the step).
//
// The flat gradient bucket of every rank lives in symmetric memory (torch.distributed._symmetric_memory: the same
// allocation mapped into every process of the node), so a kernel can address a peer's bucket directly through
// NVLink / NVSwitch. 1.7 MB of gradients is a latency problem, not a bandwidth one: an NCCL all-reduce costs ~30 us on the
// critical path between the backward pass and the optimizer; here it is two cross-GPU barriers (~2-3 us each) around
// ~3 MB of peer traffic:
//   barrier 1   every rank has finished its backward pass (the kernel is stream-ordered after it)
//   phase 1     two-shot: rank r owns slice r of the vector; it sums slice r of ALL ranks' buckets in rank order
//               0..W-1 (peer loads; the same order on every rank -> one bit pattern everywhere), scales by 1/W and
//               writes the average back into slice r of EVERY rank's bucket (peer stores)
//   barrier 2   every slice of my bucket now holds the average
//   phase 2     Adam over the whole vector from the local bucket, clearing it for the next step
// Barriers are per CTA: CTA b of rank r exchanges flags with CTA b of every peer (signal pad in symmetric memory, slot
// [b][sender]; put = CAS 0->1 with release at system scope, wait = CAS 1->0 with acquire: the flags reset themselves
// and a fast rank cannot overtake a slow one by more than one barrier), and CTA b touches the same elements of every
// slice in both phases, so no grid-wide synchronisation is needed. The grid is at most one CTA per SM (all CTAs
// co-resident: a waiting CTA can never keep the CTA it waits for from being scheduled). A spin that lasts longer than
// ~2 s raises `*error` and gives up instead of hanging the GPU.
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* a, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* a, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_f4(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void peer_barrier(uint32_t* const* pads, int rank, int world, int32_t* error) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    const long long t0 = clock64();
    const long long limit = 4000000000LL;  // ~2 s at 1.9 GHz
    uint32_t* theirs = pads[peer] + (size_t)blockIdx.x * world + rank;
    while (cas_release_sys(theirs, 0u, 1u) != 0u)
      if (clock64() - t0 > limit) { atomicExch(error, 1); break; }
    uint32_t* mine = pads[rank] + (size_t)blockIdx.x * world + peer;
    while (cas_acquire_sys(mine, 1u, 0u) != 1u)
      if (clock64() - t0 > limit) { atomicExch(error, 2); break; }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
allreduce_adam_kernel(float* __restrict__ p, float* const* __restrict__ grads, uint32_t* const* __restrict__ pads, int rank,
                      int world, float* __restrict__ m, float* __restrict__ v, int64_t n4, float lr, float b1, float b2,
                      float eps, float wd, int32_t* step, int32_t* counter, int zero_grad, int32_t* error) {
  const int t = *reinterpret_cast<volatile int32_t*>(step) + 1;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step_size = lr / bc1;
  const int64_t per = (n4 + world - 1) / world;  // float4 elements per slice
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4* const mine = reinterpret_cast<float4*>(grads[rank]);

  peer_barrier(pads, rank, world, error);
  {  // phase 1: average my slice over the ranks, in rank order, and publish it to every rank
    const int64_t lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
    const float inv = 1.f / (float)world;
    for (int64_t j = lo + first; j < hi; j += stride) {
      float4 s = ld_sys_f4(reinterpret_cast<const float4*>(grads[0]) + j);
      for (int r = 1; r < world; ++r) {
        const float4 x = ld_sys_f4(reinterpret_cast<const float4*>(grads[r]) + j);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
      for (int r = 0; r < world; ++r) st_sys_f4(reinterpret_cast<float4*>(grads[r]) + j, s);
    }
  }
  peer_barrier(pads, rank, world, error);
  // phase 2: Adam from the local bucket; CTA b visits, in every slice, exactly the elements CTA b of the owner published
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int q = 0; q < world; ++q) {
    const int64_t lo = per * q, hi = (lo + per < n4) ? lo + per : n4;
    for (int64_t j = lo + first; j < hi; j += stride) {
      const float4 g4 = ld_sys_f4(mine + j);
      const float4 pj = p4[j], mj = m4[j], vj = v4[j];
      float gg[4] = {g4.x, g4.y, g4.z, g4.w}, pp[4] = {pj.x, pj.y, pj.z, pj.w};
      float mm[4] = {mj.x, mj.y, mj.z, mj.w}, vv[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float gi = gg[c];
        if (wd != 0.f) gi = fmaf(wd, pp[c], gi);
        const float mi = b1 * mm[c] + (1.f - b1) * gi;
        const float vi = b2 * vv[c] + (1.f - b2) * gi * gi;
        mm[c] = mi;
        vv[c] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pp[c] = pp[c] - step_size * (mi / denom);
      }
      p4[j] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      m4[j] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      v4[j] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (zero_grad) mine[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    *step = t;
    *counter = 0;
  }
}

}  // namespace cwn

using namespace cwn;

extern "C" int cwn_allreduce_adam_step_f32(float* param, float* const* grad_ptrs, uint32_t* const* signal_pads,
                                           int32_t rank, int32_t world, int32_t n_ctas, float* exp_avg,
                                           float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                                           float weight_decay, int32_t* step, int32_t* counter, int32_t zero_grad,
                                           int32_t* error, cwn_stream_t stream) {
  if (n <= 0 || n % 4 != 0) return fail(CWN_E_SHAPE, "cwn_allreduce_adam_step_f32: n must be a positive multiple of 4");
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return fail(CWN_E_SHAPE, "cwn_allreduce_adam_step_f32: bad rank / world");
  if (n_ctas < 1 || n_ctas > kNumSMs) return fail(CWN_E_SHAPE, "cwn_allreduce_adam_step_f32: 1 <= n_ctas <= 148 (CTAs must be co-resident)");
  if (!param || !grad_ptrs || !signal_pads || !exp_avg || !exp_avg_sq || !step || !counter || !error)
    return fail(CWN_E_NULL, "cwn_allreduce_adam_step_f32");
  if (!aligned16(param) || !aligned16(exp_avg) || !aligned16(exp_avg_sq)) return fail(CWN_E_ALIGN, "cwn_allreduce_adam_step_f32");
  allreduce_adam_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(param, grad_ptrs, signal_pads, rank, world, exp_avg,
                                                                   exp_avg_sq, n / 4, lr, beta1, beta2, eps, weight_decay,
                                                                   step, counter, zero_grad, error);
  return launched("allreduce_adam_kernel");
}

extern "C" int cwn_adam_step_f32(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t* step,
                                 int32_t* counter, int32_t zero_grad, cwn_stream_t stream) {
  if (n < 0) return fail(CWN_E_SHAPE, "cwn_adam_step_f32: negative size");
  if (n == 0) return CWN_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || !counter) return fail(CWN_E_NULL, "cwn_adam_step_f32");
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_pdl(adam_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                             weight_decay, step, counter, zero_grad);
  return launched("adam_kernel");
}
