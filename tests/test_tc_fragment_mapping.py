"""Index arithmetic of the opt-in tensor-core path of `linear_fwd_fast_kernel` (csrc/dense.cu, `TC = true`), emulated
on the CPU: which shared-memory element every lane feeds into `mma.sync.m16n8k8` (PTX fragment layout: lane = 4 gq + tq
holds A[gq | gq+8][tq | tq+4], B[k = tq | tq+4][n = gq], C[gq | gq+8][2 tq | 2 tq+1]), which output element it ends up
owning, and which warps own a column when the BatchNorm partials are combined. The emulation performs the MMA from the
fragments exactly as the hardware defines it, so a wrong row / column / k offset in the kernel's formulas shows up as a
wrong product. (The kernel itself can only run on the GPU box; this pins its address arithmetic in the CPU tier.)"""
import numpy as np
import pytest


def emulate_tile(X, W, R):
    """X [16R, K], W [64, K] -> (out [16R, 64] assembled from the lanes' fragments, coverage count per output)."""
    TR, K = X.shape
    NT = R
    out = np.zeros((TR, 64))
    cover = np.zeros((TR, 64), dtype=int)
    owners = {}
    for warp in range(8):
        mb, nb0 = warp % R, (warp // R) * NT
        for j in range(NT):
            C = np.zeros((16, 8))
            for k in range(0, K, 8):
                A = np.zeros((16, 8))
                B = np.zeros((8, 8))
                for lane in range(32):
                    gq, tq = lane >> 2, lane & 3
                    a_lo_row, a_hi_row = mb * 16 + gq, mb * 16 + gq + 8          # kernel: a_lo_row / a_hi_row
                    A[gq, tq], A[gq + 8, tq] = X[a_lo_row, k + tq], X[a_hi_row, k + tq]
                    A[gq, tq + 4], A[gq + 8, tq + 4] = X[a_lo_row, k + tq + 4], X[a_hi_row, k + tq + 4]
                    b_row = (nb0 + j) * 8 + gq                                      # kernel: b_row0 + j * 8 * ld
                    B[tq, gq], B[tq + 4, gq] = W[b_row, k + tq], W[b_row, k + tq + 4]
                C += A @ B
            for lane in range(32):
                gq, tq = lane >> 2, lane & 3
                r_lo, r_hi = mb * 16 + gq, mb * 16 + gq + 8
                cn = (nb0 + j) * 8 + 2 * tq
                for (r, c, v) in ((r_lo, cn, C[gq, 2 * tq]), (r_lo, cn + 1, C[gq, 2 * tq + 1]),
                                  (r_hi, cn, C[gq + 8, 2 * tq]), (r_hi, cn + 1, C[gq + 8, 2 * tq + 1])):
                    out[r, c] += v
                    cover[r, c] += 1
                    owners.setdefault(c, set()).add(warp)
    return out, cover, owners


@pytest.mark.parametrize('R', [2, 4])
def test_tensor_core_fragment_mapping_reproduces_the_tile_product(R):
    rng = np.random.default_rng(R)
    K = 24
    X, W = rng.standard_normal((16 * R, K)), rng.standard_normal((64, K))
    out, cover, owners = emulate_tile(X, W, R)
    assert (cover == 1).all()                      # every output element is owned by exactly one lane
    np.testing.assert_allclose(out, X @ W.T, rtol=1e-12, atol=1e-12)
    # BatchNorm partials: column `tid` is combined from red[owner0 .. owner0 + R - 1] with owner0 = ((tid >> 3) / NT) * R
    NT = R
    for tid in range(64):
        owner0 = ((tid >> 3) // NT) * R
        assert owners[tid] == set(range(owner0, owner0 + R))


def test_shared_memory_pitch_makes_fragment_loads_conflict_free():
    """Rows of Xs / Ws are K + 4 floats apart; with K a multiple of 32 (the shapes of the models) the 8 x 4 addresses
    gq * (K + 4) + tq of one fragment load fall into 32 different banks."""
    for K in (32, 64, 128):
        banks = {(gq * (K + 4) + tq) % 32 for gq in range(8) for tq in range(4)}
        assert len(banks) == 32


def _mma_from_lanes(a_of_lane, b_of_lane):
    """One m16n8k8 MMA from per-lane fragment values: a_of_lane(lane) -> (a0, a1, a2, a3), b_of_lane(lane) -> (b0, b1)."""
    A, B = np.zeros((16, 8)), np.zeros((8, 8))
    for lane in range(32):
        gq, tq = lane >> 2, lane & 3
        a0, a1, a2, a3 = a_of_lane(lane)
        A[gq, tq], A[gq + 8, tq], A[gq, tq + 4], A[gq + 8, tq + 4] = a0, a1, a2, a3
        b0, b1 = b_of_lane(lane)
        B[tq, gq], B[tq + 4, gq] = b0, b1
    return A @ B


@pytest.mark.parametrize('TR', [32, 64])
def test_unit_bwd_tensor_core_address_arithmetic(TR):
    """The two products of `unit_bwd_fast_kernel<..., TC = true>` with the kernel's own flat shared-memory index
    expressions: g_in chunk = Gz [TR x h] * Ws [h x 64 columns from kc]; weight-gradient chunk = GzT [64 x TR] * Xs
    [TR x 64 columns from kc]. Shared arrays are emulated as flat vectors with the kernel's pitches."""
    rng = np.random.default_rng(TR)
    R, h, K, kc = TR // 16, 64, 128, 64
    ldh, ldk, LDR = h + 4, K + 4, TR + 4
    Gz2, Ws2, Xs2 = rng.standard_normal((TR, h)), rng.standard_normal((h, K)), rng.standard_normal((TR, K))
    Gz, Ws, Xs = np.zeros(TR * ldh), np.zeros(h * ldk + 64), np.zeros(TR * ldk + 64)
    for r in range(TR):
        Gz[r * ldh:r * ldh + h] = Gz2[r]
        Xs[r * ldk:r * ldk + K] = Xs2[r]
    for c in range(h):
        Ws[c * ldk:c * ldk + K] = Ws2[c]
    ZT = np.zeros(64 * LDR)                       # transpose of g_z: ZT[c * LDR + r]
    for c in range(h):
        for r in range(TR):
            ZT[c * LDR + r] = Gz2[r, c]
    # ---- product 1
    out1, cover1 = np.zeros((TR, 64)), np.zeros((TR, 64), dtype=int)
    NT = R
    for warp in range(8):
        mb, nb0 = warp % R, (warp // R) * NT
        for j in range(NT):
            C = np.zeros((16, 8))
            for c0 in range(0, h, 8):
                def a_of(lane):
                    gq, tq = lane >> 2, lane & 3
                    lo = (mb * 16 + gq) * ldh + tq
                    hi = lo + 8 * ldh
                    return Gz[lo + c0], Gz[hi + c0], Gz[lo + c0 + 4], Gz[hi + c0 + 4]

                def b_of(lane):
                    gq, tq = lane >> 2, lane & 3
                    bp = tq * ldk + kc + nb0 * 8 + gq + c0 * ldk + j * 8
                    return Ws[bp], Ws[bp + 4 * ldk]
                C += _mma_from_lanes(a_of, b_of)
            for lane in range(32):
                gq, tq = lane >> 2, lane & 3
                r_lo, cn = mb * 16 + gq, (nb0 + j) * 8 + 2 * tq
                for (r, c, v) in ((r_lo, cn, C[gq, 2 * tq]), (r_lo, cn + 1, C[gq, 2 * tq + 1]),
                                  (r_lo + 8, cn, C[gq + 8, 2 * tq]), (r_lo + 8, cn + 1, C[gq + 8, 2 * tq + 1])):
                    out1[r, c] += v
                    cover1[r, c] += 1
    assert (cover1 == 1).all()
    np.testing.assert_allclose(out1, Gz2 @ Ws2[:, kc:kc + 64], rtol=1e-12, atol=1e-12)
    # ---- product 2 (mt = 0)
    out2, cover2 = np.zeros((64, 64)), np.zeros((64, 64), dtype=int)
    for warp in range(8):
        mb, nb0 = warp & 3, (warp >> 2) * 4
        for j in range(4):
            C = np.zeros((16, 8))
            for r0 in range(0, TR, 8):
                def a_of(lane):
                    gq, tq = lane >> 2, lane & 3
                    lo = (0 * 64 + mb * 16 + gq) * LDR + tq
                    hi = lo + 8 * LDR
                    return ZT[lo + r0], ZT[hi + r0], ZT[lo + r0 + 4], ZT[hi + r0 + 4]

                def b_of(lane):
                    gq, tq = lane >> 2, lane & 3
                    bp = tq * ldk + kc + nb0 * 8 + gq + r0 * ldk + j * 8
                    return Xs[bp], Xs[bp + 4 * ldk]
                C += _mma_from_lanes(a_of, b_of)
            for lane in range(32):
                gq, tq = lane >> 2, lane & 3
                c_lo, k = mb * 16 + gq, (nb0 + j) * 8 + 2 * tq
                for (c, kk, v) in ((c_lo, k, C[gq, 2 * tq]), (c_lo, k + 1, C[gq, 2 * tq + 1]),
                                   (c_lo + 8, k, C[gq + 8, 2 * tq]), (c_lo + 8, k + 1, C[gq + 8, 2 * tq + 1])):
                    out2[c, kk] += v
                    cover2[c, kk] += 1
    assert (cover2 == 1).all()
    np.testing.assert_allclose(out2, Gz2.T @ Xs2[:, kc:kc + 64], rtol=1e-12, atol=1e-12)
