// Hardware probe 3: latency of the tcgen05 building blocks at the tiny-tile sizes of the dense kernels (clock64 stamps
// of ONE CTA; operands are zeros, only timing matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cwn_b200/csrc -o tools/_bin/tc5_probe3 tools/tc5_probe3.cu
//   tools/_bin/tc5_probe3 <M> <N> <n_mma> <n_acc>
// n_mma instructions (K = 8 each) are issued round-robin over n_acc accumulators (n_acc = 1: one dependent chain).
#include <cstdio>
#include <cstdlib>
#include "tc5.cuh"
using namespace cwn::tc5;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__global__ void __launch_bounds__(256) probe3_kernel(int M, int N, int n_mma, int n_acc, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  long long t[8];
  for (uint32_t i = tid; i < 160 * 1024 / 16; i += 256) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  t[0] = clock64();
  if (warp == 0) tmem_alloc(&tmem_s, 512);
  t[1] = clock64();
  fence_async_smem();
  t[2] = clock64();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  t[3] = clock64();
  const uint32_t tmem = tmem_s;
  const Tiled ta(128), tb(N);
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(M, N, 0, 0);
    const uint32_t a = smem_u32(smem), b = a + 80 * 1024;
    // lean issue loop: descriptors advance by adding to the start-address field, accumulators rotate without a division
    const uint64_t da0 = smem_desc(a, ta.s_c, ta.s_r), db0 = smem_desc(b, tb.s_c, tb.s_r);
    const uint64_t sa = (2 * ta.s_c) >> 4, sb = (2 * tb.s_c) >> 4;
    int acc = 0, ko = 0;
    for (int i = 0; i < n_mma; ++i) {
      mma_tf32(tmem + (uint32_t)(acc * N), da0 + (uint64_t)ko * sa, db0 + (uint64_t)ko * sb, idesc, i >= n_acc);
      if (++acc == n_acc) acc = 0;
      ko = (ko + 1) & 7;
    }
    t[4] = clock64();
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  t[5] = clock64();
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16), v);
  tmem_ld_wait();
  t[6] = clock64();
  float s = 0.f;
  for (int j = 0; j < 16; ++j) s += v[j];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  t[7] = clock64();
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) out[i] = t[i];
    out[8] = (long long)s;
  }
}

int main(int argc, char** argv) {
  if (argc < 5) { printf("usage: M N n_mma n_acc\n"); return 1; }
  const int M = atoi(argv[1]), N = atoi(argv[2]), n_mma = atoi(argv[3]), n_acc = atoi(argv[4]);
  if (n_acc * N > 512) { printf("too many accumulators\n"); return 1; }
  long long* d;
  CK(cudaMalloc(&d, 16 * 8));
  CK(cudaFuncSetAttribute(probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long h[16];
  for (int rep = 0; rep < 3; ++rep) {  // the last repetition (warm instruction cache) is the one reported
    probe3_kernel<<<1, 256, 160 * 1024>>>(M, N, n_mma, n_acc, d);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost));
  }
  printf("M=%d N=%d n_mma=%d n_acc=%d | alloc %lld | fence.proxy.async %lld | fence+bar.sync %lld | issue loop %lld (%.1f/mma) | "
         "issue..complete %lld (%.1f/mma) | tmem ld16+wait %lld | sync+dealloc %lld cycles\n",
         M, N, n_mma, n_acc, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], (double)(h[4] - h[3]) / n_mma, h[5] - h[3],
         (double)(h[5] - h[3]) / n_mma, h[6] - h[5], h[7] - h[6]);
  return 0;
}
