// Hardware probe 2 for tcgen05 kind::tf32: MN-major operands (128-byte swizzle with 32-byte atoms — the only layout the
// hardware accepts for MN-major 32-bit data) and the TMEM lane <-> row mapping of M = 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cwn_b200/csrc -o tools/_bin/tc5_probe2 tools/tc5_probe2.cu
//   tools/_bin/tc5_probe2 <a_fmt> <b_fmt> <K> <N> <M> <group_major>
//   a_fmt / b_fmt: 0 = K-major tiled (tc5.cuh Tiled), buffer [M|N][K];  1 = MN-major swizzled, buffer [K][M|N]
//   group_major  : MN-major atom placement: 0 = LBO = 512 (32-column groups adjacent), SBO = 512 * (cols / 32)
//                                           1 = SBO = 512 (4-row groups adjacent),     LBO = 512 * (rows / 4)
// D is read back from ALL 128 TMEM lanes; the host reports which logical row every lane holds.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc5.cuh"

using namespace cwn::tc5;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Args { const float* a; const float* b; float* d; int K, N, M, a_fmt, b_fmt, gm; };

__global__ void __launch_bounds__(128) probe2_kernel(Args p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.K, N = p.N, M = p.M;
  const uint32_t region = 56 * 1024;  // hi / lo of A, then of B: 4 regions of 56 KB (zero-filled) keep every read in bounds
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + region;
  unsigned char* b_hi = smem + 2 * region;
  unsigned char* b_lo = smem + 3 * region;
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (uint32_t i = tid; i < (4 * region) / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  auto fill = [&](const float* src, int fmt, int mn, unsigned char* hi, unsigned char* lo, uint32_t& lbo, uint32_t& sbo, uint32_t& step) {
    if (fmt == 0) {  // buffer [mn][K]
      const Tiled t(mn);
      for (int i = tid; i < mn * (K / 4); i += 128) {
        const int r = i / (K / 4), c4 = i % (K / 4);
        float4 h, l;
        split_tf32x4(*reinterpret_cast<const float4*>(src + (size_t)r * K + c4 * 4), h, l);
        *reinterpret_cast<float4*>(hi + t.off(r, c4)) = h;
        *reinterpret_cast<float4*>(lo + t.off(r, c4)) = l;
      }
      lbo = t.s_c; sbo = t.s_r; step = 2 * t.s_c;
    } else {         // buffer [K][mn]
      lbo = p.gm ? 512u * (uint32_t)(K / 4) : 512u;
      sbo = p.gm ? 512u : 512u * (uint32_t)((mn + 31) / 32);
      for (int i = tid; i < K * (mn / 4); i += 128) {
        const int r = i / (mn / 4), q = i % (mn / 4);
        float4 h, l;
        split_tf32x4(*reinterpret_cast<const float4*>(src + (size_t)r * mn + q * 4), h, l);
        *reinterpret_cast<float4*>(hi + mn_off(r, q, lbo, sbo)) = h;
        *reinterpret_cast<float4*>(lo + mn_off(r, q, lbo, sbo)) = l;
      }
      step = 2 * sbo;
    }
  };
  uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
  fill(p.a, p.a_fmt, M, a_hi, a_lo, a_lbo, a_sbo, a_step);
  fill(p.b, p.b_fmt, N, b_hi, b_lo, b_lbo, b_sbo, b_step);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(M, N, p.a_fmt, p.b_fmt);
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t ao = ks * a_step, bo = ks * b_step;
      const uint64_t dah = smem_desc(ah + ao, a_lbo, a_sbo, p.a_fmt), dal = smem_desc(al + ao, a_lbo, a_sbo, p.a_fmt);
      const uint64_t dbh = smem_desc(bh + bo, b_lbo, b_sbo, p.b_fmt), dbl = smem_desc(bl + bo, b_lbo, b_sbo, p.b_fmt);
      mma_tf32(tmem, dal, dbh, idesc, ks > 0);
      mma_tf32(tmem, dah, dbl, idesc, 1);
      mma_tf32(tmem, dah, dbh, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float acc[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d[(size_t)tid * N + c0 + j] = acc[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage: a_fmt b_fmt K N M group_major\n"); return 1; }
  Args p{};
  p.a_fmt = atoi(argv[1]); p.b_fmt = atoi(argv[2]); p.K = atoi(argv[3]); p.N = atoi(argv[4]); p.M = atoi(argv[5]); p.gm = atoi(argv[6]);
  const int K = p.K, N = p.N, M = p.M;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(4321);
  auto rnd = [&]() { float s = 0.f; for (int i = 0; i < 12; ++i) s += (float)rand() / RAND_MAX; return s - 6.f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<float> Abuf(A.size()), Bbuf(B.size());
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Abuf[p.a_fmt ? (size_t)k * M + m : (size_t)m * K + k] = A[(size_t)m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bbuf[p.b_fmt ? (size_t)k * N + n : (size_t)n * K + k] = B[(size_t)n * K + k];
  float *da, *db, *dd;
  CK(cudaMalloc(&da, A.size() * 4)); CK(cudaMalloc(&db, B.size() * 4)); CK(cudaMalloc(&dd, (size_t)128 * N * 4));
  CK(cudaMemcpy(da, Abuf.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, Bbuf.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, (size_t)128 * N * 4));
  p.a = da; p.b = db; p.d = dd;
  const size_t smem = 4 * 56 * 1024;
  CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe2_kernel<<<1, 128, smem>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)128 * N);
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<double> ref((size_t)M * N);
  double sq = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      ref[(size_t)m * N + n] = r; sq += r * r;
    }
  const double rms = sqrt(sq / ((double)M * N));
  printf("a_fmt=%d b_fmt=%d K=%d N=%d M=%d gm=%d rms(D)=%.4g | lane->row (err/rms):", p.a_fmt, p.b_fmt, K, N, M, p.gm, rms);
  int matched = 0, identity = 1;
  double worst = 0;
  for (int l = 0; l < 128; ++l) {
    int best = -1; double best_err = 1e30;
    for (int m = 0; m < M; ++m) {
      double e = 0;
      for (int n = 0; n < N; ++n) e = fmax(e, fabs((double)D[(size_t)l * N + n] - ref[(size_t)m * N + n]));
      if (e < best_err) { best_err = e; best = m; }
    }
    const bool ok = best_err / rms < 1e-4;
    if (ok) { matched++; worst = fmax(worst, best_err / rms); if (best != l) identity = 0; }
    static int prev = -100;
    const int cur = ok ? best : -1;
    if (l == 0 || cur != (prev < 0 ? -1 : prev + 1)) printf(" [%d:%d]", l, cur);
    prev = cur;
  }
  printf(" | lanes matched %d of 128, identity=%d, worst matched err/rms %.2e\n", matched, identity, worst);
  return 0;
}
