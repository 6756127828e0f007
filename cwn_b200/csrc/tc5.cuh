// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a (inline PTX), used by the dense kernels (dense_tc5.cuh).
//
// Operand layout ("core-matrix tiled", no swizzle). A row-major fp32 matrix [R][C] (R % 8 == 0, C % 4 == 0) is kept in
// shared memory as 8-row x 16-byte core matrices:
//     byte(r, c) = (r / 8) * S_r + (r % 8) * 16 + (c / 4) * S_c + (c % 4) * 4
// with S_r, S_c free multiples of 16 bytes (S_r >= 128): the "K-major" operand form (rows = M/N index, columns = inner
// index), descriptor SBO = S_r, LBO = S_c, k-step of 8: += 2 S_c. S_c = (R/8)*128 + 16 skews consecutive 16-byte column
// chunks by 4 banks: the transform-and-store loops (a lane per column chunk, i.e. coalesced global reads) are
// shared-memory conflict-free.
//
// Operands whose INNER index is the row of the row-major source (the transposed products of the backward pass,
// g_z^T X and g_z W) are "MN-major". For 32-bit data the hardware accepts exactly one MN-major layout, the 128-byte
// swizzle with 32-byte atoms (an un-swizzled MN-major tf32 descriptor silently yields zeros — profiles/
// r2_tc5_probe_accumulation.txt): rows of 32 consecutive M/N elements (128 B), 4 inner rows per 512-byte atom, the
// 32-byte chunks of a row XOR-ed with (inner row % 4); see mn_off(). Both forms are written with 16-byte stores from
// the same registers, no transposition anywhere. Verified on hardware by tools/tc5_probe*.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cwn {
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor or protocol bug must end in a trap (a reported launch failure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) asm volatile("trap;");
}

// ---- bulk async copy global -> shared (TMA engine, 1-D; completes on an mbarrier). bytes % 16 == 0, 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {  // one full warp; cols: power of two in [32, 512]
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) { tmem_alloc(smem_result, (uint32_t)COLS); }
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) { tmem_dealloc(taddr, (uint32_t)COLS); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors
// shared-memory matrix descriptor, no swizzle, Blackwell version bits; lbo / sbo in bytes (see the header comment)
// layout: 0 = no swizzle (K-major operands here), 1 = 128-byte swizzle with 32-byte atoms (the only layout the
// hardware takes for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor of kind::tf32, fp32 accumulate: M in {64, 128}, N % 8 == 0 (M = 128: N % 16 == 0), majors 0 = K, 1 = MN
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: lane l of warp w reads TMEM lane 32 (w % 4) + l, N consecutive 32-bit columns from `taddr`
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 split: v = hi + lo (+ O(2^-22 |v|)), both exactly representable in TF32 (10 explicit mantissa bits)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}
__device__ __forceinline__ void split_tf32x4(const float4 v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x);
  split_tf32(v.y, hi.y, lo.y);
  split_tf32(v.z, hi.z, lo.z);
  split_tf32(v.w, hi.w, lo.w);
}

// ---- the tiled operand layout
struct Tiled {
  uint32_t s_r, s_c;  // byte strides between 8-row groups / 4-column chunks
  __host__ __device__ static constexpr uint32_t sc_for_rows(int rows) { return (uint32_t)(rows / 8) * 128u + 16u; }
  __host__ __device__ static constexpr uint32_t bytes(int rows, int cols) { return (uint32_t)(cols / 4) * sc_for_rows(rows); }
  __host__ __device__ Tiled(int rows) : s_r(128u), s_c(sc_for_rows(rows)) {}
  __host__ __device__ __forceinline__ uint32_t off(int r, int c4) const {  // byte offset of the 16-byte chunk (row r, columns 4 c4 ..)
    return (uint32_t)(r >> 3) * s_r + (uint32_t)(r & 7) * 16u + (uint32_t)c4 * s_c;
  }
};

// Issue the 3xTF32 product D[128 x N] (+)= A * B over `inner` (multiple of 8) from split operand buffers.
// a_hi/a_lo/b_hi/b_lo: shared-memory byte addresses of the tiled buffers; (a_lbo, a_sbo, a_step) per the operand major.
// Small terms first, then hi*hi. ONE thread.
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                             uint32_t a_step, uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo,
                                             uint32_t b_step, int inner, uint32_t idesc, bool accumulate_first) {
  uint32_t acc = accumulate_first ? 1u : 0u;
  for (int k = 0; k < inner; k += 8) {
    const uint32_t ao = (uint32_t)(k >> 3) * a_step, bo = (uint32_t)(k >> 3) * b_step;
    const uint64_t dah = smem_desc(a_hi + ao, a_lbo, a_sbo), dal = smem_desc(a_lo + ao, a_lbo, a_sbo);
    const uint64_t dbh = smem_desc(b_hi + bo, b_lbo, b_sbo), dbl = smem_desc(b_lo + bo, b_lbo, b_sbo);
    mma_tf32(d_tmem, dal, dbh, idesc, acc);
    mma_tf32(d_tmem, dah, dbl, idesc, 1u);
    mma_tf32(d_tmem, dah, dbh, idesc, 1u);
    acc = 1u;
  }
}

// MN-major swizzled buffer of a row-major [R (inner)][C (M or N)] matrix: byte offset of the 16-byte chunk (r, c = 4 q).
// Atoms of 4 inner rows x 32 columns (512 B): sbo = stride between 4-row groups, lbo = stride between 32-column groups.
// The buffer must be 1024-byte aligned; descriptor layout type 1; k-step of 8 inner rows: += 2 sbo.
__host__ __device__ __forceinline__ uint32_t mn_off(int r, int q, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(q >> 3) * lbo + (uint32_t)(r >> 2) * sbo + (uint32_t)(r & 3) * 128u +
         ((((uint32_t)(q & 7) >> 1) ^ (uint32_t)(r & 3)) << 5) + ((uint32_t)(q & 1) << 4);
}

}  // namespace tc5
}  // namespace cwn
