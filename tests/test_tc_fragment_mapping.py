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
