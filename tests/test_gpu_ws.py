"""Warp-specialised TMA kernels of the HBM-bound regime (csrc/gsa_ws.cu) against the register-level kernels, bit for bit.

The register-level entry points are themselves pinned against the sequential CPU definition (test_gpu_parity.py), and
both families add the messages of a row in plan order with IEEE round-to-nearest adds, so equality must be exact on
arbitrary float data — staged tiles (shared memory), tiles that do not fit their buffers (generic pointers to global
memory), the ragged last tile, an unaligned payload tail, strided operand matrices, rows without messages."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _setup():
    from cwn_b200 import _lib, ops
    from cwn_b200.data import synthetic
    return _lib, ops, synthetic


def _windows(lib, plan, tile_rows):
    n_tiles = (plan.n_rows + tile_rows - 1) // tile_rows
    w = torch.zeros(8 * n_tiles + 4, dtype=torch.int32, device=DEV)
    lib.check(lib.load().cwn_csr_tile_windows(plan.rowptr.data_ptr(), plan.pay0.data_ptr(),
                                              plan.pay1.data_ptr() if plan.pay1 is not None else None, plan.n_rows,
                                              tile_rows, w.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return w, w[-4:].tolist()


def _strided(n, F, gen, strided):
    """[n, F] matrix, optionally a column slice of a wider one (pitch != row bytes: per-row bulk copies)."""
    if not strided:
        return torch.randn(n, F, generator=gen).to(DEV)
    return torch.randn(n, 2 * F + 8, generator=gen).to(DEV)[:, 4:4 + F]


def _adjacency(synthetic, ops, kind, units, drop_tail=0):
    index, cob, n_src, n_dst, n_cob = synthetic.tiled_adjacency(kind, units)
    if drop_tail:  # E % 4 != 0 and trailing rows without messages
        index, cob = index[:, :-drop_tail].contiguous(), (cob[:-drop_tail].contiguous() if cob is not None else None)
    index = index.to(DEV)
    cob = cob.to(DEV) if cob is not None else None
    return ops.Adjacency.of(index, n_src, n_dst, cob, n_cob if cob is not None else None)


def test_tile_windows_match_a_host_scan():
    lib, ops, synthetic = _setup()
    adj = _adjacency(synthetic, ops, 'edge_up', 1500, drop_tail=3)
    plan = adj.by_dst
    for tile_rows in (16, 64, 256):
        w, stats = _windows(lib, plan, tile_rows)
        rowptr, p0, p1 = plan.rowptr.cpu(), plan.pay0.cpu(), plan.pay1.cpu()
        n_tiles = (plan.n_rows + tile_rows - 1) // tile_rows
        w = w[:-4].view(n_tiles, 8).cpu()
        for t in list(range(0, n_tiles, 37)) + [n_tiles - 1]:
            r0, r1 = t * tile_rows, min((t + 1) * tile_rows, plan.n_rows)
            m0, m1 = int(rowptr[r0]), int(rowptr[r1])
            exp = [m0, m1, 0, 0, 0, 0, 0, 0]
            if m1 > m0:
                for a, p in enumerate((p0, p1)):
                    lo, hi = int(p[m0:m1].min()), int(p[m0:m1].max())
                    exp[2 + 2 * a], exp[3 + 2 * a] = lo, hi - lo + 1
            assert w[t].tolist() == exp, (tile_rows, t)
        cnt0 = w[:, 3].max().item()
        cnt1 = w[:, 5].max().item()
        hull = (((w[:, 1] + 3) & ~3) - (w[:, 0] & ~3))[w[:, 1] > w[:, 0]].max().item()
        assert stats[:3] == [cnt0, cnt1, hull]


@pytest.mark.parametrize('kind,units,F,tile_rows', [('edge_up', 1500, 64, 64), ('edge_up', 1400, 16, 256),
                                                    ('edge_boundary', 1400, 128, 32), ('ring_boundary', 12000, 20, 64),
                                                    ('vertex_up', 1500, 32, 128)])
@pytest.mark.parametrize('variant', ['staged', 'tight_caps', 'strided'])
def test_ws_gather_equals_row_kernels(kind, units, F, tile_rows, variant):
    lib, ops, synthetic = _setup()
    L = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    adj = _adjacency(synthetic, ops, kind, units, drop_tail=3)
    gen = torch.Generator().manual_seed(F + units)
    for plan, n_src, n_rows in ((adj.by_dst, adj.n_src, adj.n_dst), (adj.by_src, adj.n_dst, adj.n_src)):
        w, (cap0, _, capm, _) = _windows(lib, plan, tile_rows)
        if variant == 'tight_caps':  # most tiles exceed the buffers -> generic path, mixed with staged ones
            cap0, capm = max(cap0 // 2, 1), max((capm // 2) & ~3, 4)
        x = _strided(n_src, F, gen, variant == 'strided')
        res = _strided(n_rows, F, gen, variant == 'strided')
        eps = torch.tensor([0.375], device=DEV)
        for reduce, with_res in ((0, False), (0, True), (1, False)):
            ref = torch.full((n_rows, F), float('nan'), device=DEV)
            out = torch.full((n_rows, F), float('nan'), device=DEV)
            r = res if with_res else None
            lib.check(L.cwn_csr_gather_reduce_f32(x.data_ptr(), x.stride(0), plan.rowptr.data_ptr(), plan.pay0.data_ptr(),
                                                  n_rows, F, r.data_ptr() if with_res else None,
                                                  r.stride(0) if with_res else F, eps.data_ptr(), ref.data_ptr(), F, reduce,
                                                  st))
            lib.check(L.cwn_csr_gather_reduce_ws_f32(x.data_ptr(), x.stride(0), plan.rowptr.data_ptr(),
                                                     plan.pay0.data_ptr(), plan.E, w.data_ptr(), tile_rows, cap0, capm,
                                                     n_rows, F, r.data_ptr() if with_res else None,
                                                     r.stride(0) if with_res else F, eps.data_ptr(), out.data_ptr(), F,
                                                     reduce, st))
            torch.cuda.synchronize()
            assert torch.equal(out, ref), (kind, F, variant, reduce, with_res, int((out != ref).sum()))


@pytest.mark.parametrize('kind,units,F,tile_rows,act', [('edge_up', 1500, 64, 64, 1), ('edge_up', 1400, 16, 128, 2),
                                                        ('vertex_up', 1500, 32, 64, 4), ('edge_up', 1400, 128, 32, 3),
                                                        ('vertex_up', 1600, 64, 64, 0)])
@pytest.mark.parametrize('variant', ['staged', 'tight_caps', 'strided'])
def test_ws_coboundary_passes_equal_row_kernels(kind, units, F, tile_rows, act, variant):
    lib, ops, synthetic = _setup()
    L = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    adj = _adjacency(synthetic, ops, kind, units, drop_tail=1)
    gen = torch.Generator().manual_seed(F + units + act)
    strided = variant == 'strided'
    P, Q = _strided(adj.n_src, F, gen, strided), _strided(adj.n_cob, F, gen, strided)
    res, G = _strided(adj.n_dst, F, gen, strided), _strided(adj.n_dst, F, gen, strided)
    eps = torch.tensor([-0.125], device=DEV)

    def caps(plan):
        tr = tile_rows
        while True:  # the by-coboundary plan has few, heavy rows: its windows need shorter tiles
            w, (c0, c1, cm, _) = _windows(lib, plan, tr)
            if L.cwn_csr_ws_stages(F, tr, c0, c1, cm, 2, 1) >= 2 or tr == 4:
                break
            tr //= 2
        if variant == 'tight_caps':
            c0, c1, cm = max(c0 // 2, 1), max(c1 - 1, 1), max((cm // 2) & ~3, 4)
        return w, tr, c0, c1, cm

    plan = adj.by_dst
    w, tr, c0, c1, cm = caps(plan)
    for with_res in (False, True):
        ref = torch.full((adj.n_dst, F), float('nan'), device=DEV)
        out = torch.full((adj.n_dst, F), float('nan'), device=DEV)
        rp = res.data_ptr() if with_res else None
        rl = res.stride(0) if with_res else F
        lib.check(L.cwn_csr_cob_fwd_f32(P.data_ptr(), P.stride(0), Q.data_ptr(), Q.stride(0), plan.rowptr.data_ptr(),
                                        plan.pay0.data_ptr(), plan.pay1.data_ptr(), adj.n_dst, F, act, rp, rl,
                                        eps.data_ptr(), ref.data_ptr(), F, st))
        lib.check(L.cwn_csr_cob_fwd_ws_f32(P.data_ptr(), P.stride(0), Q.data_ptr(), Q.stride(0), plan.rowptr.data_ptr(),
                                           plan.pay0.data_ptr(), plan.pay1.data_ptr(), plan.E, w.data_ptr(), tr,
                                           c0, c1, cm, adj.n_dst, F, act, rp, rl, eps.data_ptr(), out.data_ptr(), F, st))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), ('fwd', kind, F, variant, with_res, int((out != ref).sum()))
    for plan, A, B, n_rows in ((adj.by_src, P, Q, adj.n_src), (adj.by_cob, Q, P, adj.n_cob)):
        w, tr, c0, c1, cm = caps(plan)
        ref = torch.full((n_rows, F), float('nan'), device=DEV)
        out = torch.full((n_rows, F), float('nan'), device=DEV)
        lib.check(L.cwn_csr_cob_bwd_f32(G.data_ptr(), G.stride(0), A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                        plan.rowptr.data_ptr(), plan.pay0.data_ptr(), plan.pay1.data_ptr(), n_rows, F,
                                        act, ref.data_ptr(), F, st))
        lib.check(L.cwn_csr_cob_bwd_ws_f32(G.data_ptr(), G.stride(0), A.data_ptr(), A.stride(0), B.data_ptr(),
                                           B.stride(0), plan.rowptr.data_ptr(), plan.pay0.data_ptr(),
                                           plan.pay1.data_ptr(), plan.E, w.data_ptr(), tr, c0, c1, cm, n_rows, F,
                                           act, out.data_ptr(), F, st))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), ('bwd', kind, F, variant, n_rows, int((out != ref).sum()))


def test_ops_select_the_ws_kernels_for_large_block_diagonal_batches():
    """Through the public operators: a 40k-row block-diagonal batch takes the warp-specialised kernels (the plans get
    their tile windows), a uniform-random one of the same size does not (windows span the matrix), and the results equal
    those of the register-level kernels (CWN_B200_WS=0 behaviour) exactly, forward and backward."""
    lib, ops, synthetic = _setup()
    index, cob, n_src, n_dst, n_cob = synthetic.tiled_adjacency('edge_up', 1600)
    index, cob = index.to(DEV), cob.to(DEV)
    gen = torch.Generator().manual_seed(5)
    F = 64
    P0, Q0 = torch.randn(n_src, F, generator=gen).to(DEV), torch.randn(n_cob, F, generator=gen).to(DEV)
    g = torch.randn(n_dst, F, generator=gen).to(DEV)
    results = []
    for enabled in (True, False):
        ops._ws_enabled = enabled
        try:
            ops.clear_plan_cache(index)
            x = P0.clone().requires_grad_(True)
            out = ops.gather_scatter(x, index, n_dst, x_res=x, eps=torch.tensor([0.5], device=DEV))
            out.backward(g)
            P, Q = P0.clone().requires_grad_(True), Q0.clone().requires_grad_(True)
            o2 = ops.cob_pass(P, Q, index, cob, n_dst, act='relu')
            o2.backward(g)
            adj = ops.Adjacency.of(index, n_src, n_dst, cob, n_cob)
            used = any(p.ws for p in adj._plans.values())
            assert used == enabled
            results.append([out.detach(), x.grad, o2.detach(), P.grad, Q.grad])
        finally:
            ops._ws_enabled = True
    for a, b in zip(*results):
        assert torch.equal(a, b)
    ridx, _, rs, rd, _ = synthetic.random_adjacency(40_000, 40_000, 120_000)
    ridx = ridx.to(DEV)
    ops.gather_scatter(torch.randn(rs, F, device=DEV), ridx, rd)
    radj = ops.Adjacency.of(ridx, rs, rd)
    assert ops._ws_config(radj.by_dst, F, 1, False) is None
