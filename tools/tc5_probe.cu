// Hardware probe for the tcgen05 (kind::tf32) building blocks of cwn_b200/csrc/tc5.cuh — run on a B200 through gpurun:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cwn_b200/csrc -o tools/_bin/tc5_probe tools/tc5_probe.cu
//   gpurun_out/tc5_probe <a_mn> <b_mn> <K> <N> <mode> <nblk> <swap> <positive>
// One CTA forms D[128 x N] = A[128 x K] * B[N x K]^T from fp32 inputs with 3xTF32 split operands on the tensor cores
// and the result is compared with an fp64 reference and with a sequential fp32 FMA chain (what the FFMA kernels do).
//   a_mn / b_mn : 0 = the operand buffer is [M or N][K] (K-major descriptor), 1 = it is [K][M or N] (MN-major)
//   mode        : 0 = one accumulator, lo*hi, hi*lo, hi*hi interleaved per k-step
//                 1 = small terms and hi*hi in separate accumulators, added (RN) in the epilogue
//                 2 = hi*hi split over `nblk` accumulators by k-range (+ one for the small terms), RN-summed in the epilogue
//                 3 = plain TF32 (hi*hi only)
//   swap        : 1 = exchange LBO / SBO in the descriptors (layout-semantics cross-check; must be WRONG)
// Answers (a) are the descriptor semantics of tc5.cuh right, (b) how far is the tensor-core fp32 accumulation from RN.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc5.cuh"

using namespace cwn::tc5;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Args { const float* a; const float* b; float* d; int K, N, a_mn, b_mn, mode, nblk, swap; };

__global__ void __launch_bounds__(128) probe_kernel(Args p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.K, N = p.N;
  // A buffer: rows x cols = a_mn ? [K][128] : [128][K];  B buffer: b_mn ? [K][N] : [N][K]
  const int a_rows = p.a_mn ? K : 128, a_cols = p.a_mn ? 128 : K;
  const int b_rows = p.b_mn ? K : N, b_cols = p.b_mn ? N : K;
  const Tiled ta(a_rows), tb(b_rows);
  const uint32_t a_bytes = Tiled::bytes(a_rows, a_cols), b_bytes = Tiled::bytes(b_rows, b_cols);
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < a_rows * (a_cols / 4); i += 128) {
    const int r = i / (a_cols / 4), c4 = i % (a_cols / 4);
    const float4 v = *reinterpret_cast<const float4*>(p.a + (size_t)r * a_cols + c4 * 4);
    float4 h, l;
    split_tf32x4(v, h, l);
    *reinterpret_cast<float4*>(a_hi + ta.off(r, c4)) = h;
    *reinterpret_cast<float4*>(a_lo + ta.off(r, c4)) = l;
  }
  for (int i = tid; i < b_rows * (b_cols / 4); i += 128) {
    const int r = i / (b_cols / 4), c4 = i % (b_cols / 4);
    const float4 v = *reinterpret_cast<const float4*>(p.b + (size_t)r * b_cols + c4 * 4);
    float4 h, l;
    split_tf32x4(v, h, l);
    *reinterpret_cast<float4*>(b_hi + tb.off(r, c4)) = h;
    *reinterpret_cast<float4*>(b_lo + tb.off(r, c4)) = l;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int nacc = p.mode == 0 || p.mode == 3 ? 1 : (p.mode == 1 ? 2 : p.nblk + 1);
  if (tid == 0) {
    uint32_t a_lbo = p.a_mn ? ta.s_r : ta.s_c, a_sbo = p.a_mn ? ta.s_c : ta.s_r, a_step = p.a_mn ? ta.s_r : 2 * ta.s_c;
    uint32_t b_lbo = p.b_mn ? tb.s_r : tb.s_c, b_sbo = p.b_mn ? tb.s_c : tb.s_r, b_step = p.b_mn ? tb.s_r : 2 * tb.s_c;
    if (p.swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t idesc = idesc_tf32(128, N, p.a_mn, p.b_mn);
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    const int ksteps = K / 8;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t ao = ks * a_step, bo = ks * b_step;
      const uint64_t dah = smem_desc(ah + ao, a_lbo, a_sbo), dal = smem_desc(al + ao, a_lbo, a_sbo);
      const uint64_t dbh = smem_desc(bh + bo, b_lbo, b_sbo), dbl = smem_desc(bl + bo, b_lbo, b_sbo);
      if (p.mode == 0) {
        mma_tf32(tmem, dal, dbh, idesc, ks > 0);
        mma_tf32(tmem, dah, dbl, idesc, 1);
        mma_tf32(tmem, dah, dbh, idesc, 1);
      } else if (p.mode == 3) {
        mma_tf32(tmem, dah, dbh, idesc, ks > 0);
      } else {
        const int per = (ksteps + (nacc - 1) - 1) / (nacc - 1);
        const int blk = ks / per;
        mma_tf32(tmem, dal, dbh, idesc, ks > 0);                                // accumulator 0: small terms
        mma_tf32(tmem, dah, dbl, idesc, 1);
        mma_tf32(tmem + (uint32_t)(1 + blk) * N, dah, dbh, idesc, (ks % per) > 0);  // accumulator 1 + blk: hi*hi
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const int row = tid;  // TMEM lane == D row (M = 128)
  for (int c0 = 0; c0 < N; c0 += 16) {
    float acc[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
    if (nacc > 1) {
      // big blocks summed in order first (RN), the small-term accumulator last
      float small[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) small[j] = acc[j];
      float big[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + N + c0, big);
      tmem_ld_wait();
      for (int b = 2; b < nacc; ++b) {
        float more[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + b * N + c0, more);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) big[j] += more[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = big[j] + small[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d[(size_t)row * N + c0 + j] = acc[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  if (argc < 9) { printf("usage: a_mn b_mn K N mode nblk swap positive\n"); return 1; }
  Args p{};
  p.a_mn = atoi(argv[1]); p.b_mn = atoi(argv[2]); p.K = atoi(argv[3]); p.N = atoi(argv[4]);
  p.mode = atoi(argv[5]); p.nblk = atoi(argv[6]); p.swap = atoi(argv[7]);
  const int positive = atoi(argv[8]);
  const int K = p.K, N = p.N, M = 128;
  if ((p.mode == 2 ? p.nblk + 1 : 2) * N > 512) { printf("too many accumulators\n"); return 1; }
  std::vector<float> A((size_t)M * K), B((size_t)N * K);  // logical A[m][k], B[n][k]
  srand(1234);
  auto rnd = [&]() {  // ~N(0,1) (sum of uniforms), or |.| + 0.5 for the bias test
    float s = 0.f;
    for (int i = 0; i < 12; ++i) s += (float)rand() / RAND_MAX;
    s -= 6.f;
    return positive ? fabsf(s) + 0.5f : s;
  };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<float> Abuf(A.size()), Bbuf(B.size());
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Abuf[p.a_mn ? (size_t)k * M + m : (size_t)m * K + k] = A[(size_t)m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bbuf[p.b_mn ? (size_t)k * N + n : (size_t)n * K + k] = B[(size_t)n * K + k];
  float *da, *db, *dd;
  CK(cudaMalloc(&da, A.size() * 4)); CK(cudaMalloc(&db, B.size() * 4)); CK(cudaMalloc(&dd, (size_t)M * N * 4));
  CK(cudaMemcpy(da, Abuf.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, Bbuf.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, (size_t)M * N * 4));
  p.a = da; p.b = db; p.d = dd;
  const int a_rows = p.a_mn ? K : 128, a_cols = p.a_mn ? 128 : K, b_rows = p.b_mn ? K : N, b_cols = p.b_mn ? N : K;
  const size_t smem = 2 * (size_t)Tiled::bytes(a_rows, a_cols) + 2 * (size_t)Tiled::bytes(b_rows, b_cols);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_tc = 0, sq_tc = 0, max_ff = 0, sq_ff = 0, sq_ref = 0, bias_tc = 0, bias_ff = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      float ff = 0.f;
      for (int k = 0; k < K; ++k) {
        ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        ff = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], ff);
      }
      const double e1 = (double)D[(size_t)m * N + n] - ref, e2 = (double)ff - ref;
      max_tc = fmax(max_tc, fabs(e1)); sq_tc += e1 * e1; bias_tc += e1 * (ref >= 0 ? 1 : -1);
      max_ff = fmax(max_ff, fabs(e2)); sq_ff += e2 * e2; bias_ff += e2 * (ref >= 0 ? 1 : -1);
      sq_ref += ref * ref;
    }
  const double cnt = (double)M * N, rms_ref = sqrt(sq_ref / cnt);
  printf("a_mn=%d b_mn=%d K=%d N=%d mode=%d nblk=%d swap=%d pos=%d | rms(D)=%.4g | tensor: max %.3e rms %.3e signed-mean %.3e | "
         "fp32 FMA chain: max %.3e rms %.3e signed-mean %.3e | (relative to rms(D)) tensor max %.3e rms %.3e, FMA max %.3e rms %.3e\n",
         p.a_mn, p.b_mn, K, N, p.mode, p.nblk, p.swap, positive, rms_ref, max_tc, sqrt(sq_tc / cnt), bias_tc / cnt, max_ff,
         sqrt(sq_ff / cnt), bias_ff / cnt, max_tc / rms_ref, sqrt(sq_tc / cnt) / rms_ref, max_ff / rms_ref, sqrt(sq_ff / cnt) / rms_ref);
  return 0;
}
