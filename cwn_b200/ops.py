"""Autograd operators of the hot path, all thin wrappers over the C ABI in include/cwn_b200.h.

PyTorch is used for device memory (caching allocator), the current stream and autograd bookkeeping only; the
arithmetic of gather / message / aggregate runs in the hand-written sm_100a kernels. Every function here
requires CUDA fp32 tensors and raises otherwise (no CPU path, no torch_scatter, no eager fallback).

Replaces in the reference: `__lift__` (mp/cell_mp.py:195-198), `aggregate_*` (mp/cell_mp.py:423-479),
`InitReduceConv` (mp/layers.py:484-487), `global_add_pool/global_mean_pool` (mp/nn.py:59), the `up_attr` gather of
`data/complex.py:579-580`, and the autograd of all of them.
"""
import os

import torch
from torch import Tensor
from torch.autograd import Function

from cwn_b200 import _lib
from cwn_b200 import streams as _streams

REDUCE_CODES = {'add': 0, 'sum': 0, 'mean': 1, 'max': 2}
ACT_CODES = {'id': 0, 'relu': 1, 'elu': 2, 'sigmoid': 3, 'tanh': 4}


# --------------------------------------------------------------------------------------------- helpers
def _require_cuda_f32(t: Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'cwn_b200: `{name}` is on {t.device}; the message-passing path is CUDA-only '
                           f'(sm_100a kernels, no CPU fallback)')
    if t.dtype != torch.float32:
        raise TypeError(f'cwn_b200: `{name}` must be float32, got {t.dtype}')


def _require_cuda_float(t: Tensor, name: str, like: Tensor = None):
    """float32 (the tuned kernels) or float64 (`*_f64` entry points: the reference's SR experiments run in double,
    exp/run_exp.py:41-43); all floating operands of one call share the dtype."""
    if not t.is_cuda:
        raise RuntimeError(f'cwn_b200: `{name}` is on {t.device}; the message-passing path is CUDA-only '
                           f'(sm_100a kernels, no CPU fallback)')
    if t.dtype not in (torch.float32, torch.float64):
        raise TypeError(f'cwn_b200: `{name}` must be float32 or float64, got {t.dtype}')
    if like is not None and t.dtype != like.dtype:
        raise TypeError(f'cwn_b200: `{name}` is {t.dtype} but the other operands are {like.dtype}')


def _f64(t: Tensor) -> bool:
    return t.dtype == torch.float64


def _rows(t: Tensor) -> Tensor:
    """2-D view with unit inner stride (the C ABI takes an explicit leading dimension)."""
    if t.dim() == 1:
        t = t.unsqueeze(-1)
    if t.dim() != 2:
        raise ValueError(f'cwn_b200: expected a [rows, features] matrix, got shape {tuple(t.shape)}')
    if t.size(1) > 0 and t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _ld(t):
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)


def _stream():
    return torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------------------------- per-kernel timing
class KernelProfile(object):
    """CUDA-event timing of every C-ABI launch made while active (bench.py's roofline leg). Events are recorded on
    the launching stream right around the launch; `summary()` synchronises once at the end.

    `pad_cycles` > 0 enqueues a spin kernel of that many SM clocks on the launching stream before the start event. An
    eager step is host-bound (a launch through ctypes takes longer than a 3-15 us kernel runs), so without it the
    stream is idle when the start event is reached and the event pair measures the HOST gap until the kernel arrives
    (measured: 16 us for a 3.3 us kernel). With the pad the start event, the kernel and the stop event are all queued
    while the GPU spins, and the pair brackets device time only."""

    def __init__(self, pad_cycles: int = 0):
        self.records = []  # (kernel name, algorithmic bytes, start event, stop event, algorithmic flops)
        self.pad_cycles = int(pad_cycles)

    def __enter__(self):
        global _profile
        _profile = self
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None

    def event_overhead_ms(self, n=20):
        """Median elapsed time of an EMPTY event pair queued behind a pad (what a pair costs with nothing between)."""
        if not self.pad_cycles:
            return 0.0
        pairs = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(self.pad_cycles)
            e0.record()
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in pairs)[n // 2]

    def summary(self):
        torch.cuda.synchronize()
        self.overhead_ms = self.event_overhead_ms()
        out = {}
        for name, nbytes, e0, e1, flops in self.records:
            rec = out.setdefault(name, {'launches': 0, 'ms': 0.0, 'bytes': 0, 'flops': 0})
            rec['launches'] += 1
            rec['ms'] += max(e0.elapsed_time(e1) - self.overhead_ms, 0.0)
            rec['bytes'] += nbytes
            rec['flops'] += flops
        return out


_profile = None


def _call(name, algo_bytes, fn, *args, flops=0):
    """Invoke one C-ABI entry point, timing it when a KernelProfile is active."""
    if _profile is None:
        _lib.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if _profile.pad_cycles:
        torch.cuda._sleep(_profile.pad_cycles)
    e0.record()
    _lib.check(fn(*args), name)
    e1.record()
    _profile.records.append((name, algo_bytes, e0, e1, flops))


# --------------------------------------------------------------------------------------------- CSR plans
class Plan(object):
    """Messages grouped (stably) by one index column: `rowptr[n_rows+1]`, `perm[E]`, up to two payload columns."""
    __slots__ = ('rowptr', 'perm', 'pay0', 'pay1', 'n_rows', 'E', 'ws')

    def __init__(self, rowptr, perm, pay0, pay1, n_rows, E):
        self.rowptr, self.perm, self.pay0, self.pay1, self.n_rows, self.E = rowptr, perm, pay0, pay1, n_rows, E
        self.ws = {}  # tile height -> (windows, cap_rows0, cap_rows1, cap_msgs): see _ws_config


# Warp-specialised TMA kernels of the HBM-bound regime (csrc/gsa_ws.cu). CWN_B200_WS=0 keeps the register-level kernels.
WS_MIN_ROWS = 32768
_ws_enabled = os.environ.get('CWN_B200_WS', '1') != '0'
_WS_MIN_STAGES = int(os.environ.get('CWN_B200_WS_MIN_STAGES', '4'))
_WS_FORCED_TILE = int(os.environ.get('CWN_B200_WS_TILE', '0'))  # A/B switch: one tile height instead of the search


def _ws_ok(*mats):
    return all(m is None or (m.data_ptr() % 16 == 0 and _ld(m) % 4 == 0) for m in mats)


def _ws_config(plan: Plan, F: int, narr: int, rowop: bool):
    """(windows, tile_rows, cap_rows0, cap_rows1, cap_msgs) when the plan is large enough for the warp-specialised
    kernels and the operand windows of its tiles fit their shared-memory stages, else None. The tile windows are
    computed once per plan and tile height (cwn_csr_tile_windows) and their maxima read back (one synchronisation
    per plan, never during stream capture: a plan first seen inside a capture keeps the register-level kernels)."""
    if (not _ws_enabled or plan.n_rows < WS_MIN_ROWS or plan.E == 0 or F % 4 or F > 128 or plan.pay0 is None
            or (narr == 2 and plan.pay1 is None)):
        return None
    if plan.rowptr.data_ptr() % 16 or plan.pay0.data_ptr() % 16 or (plan.pay1 is not None and plan.pay1.data_ptr() % 16):
        return None
    lib = _lib.load()
    # the tallest tile whose stage still leaves a 4-deep pipeline: taller tiles re-read less of the neighbouring
    # tiles' windows (a window is the tile's rows plus a margin of about one complex on either side)
    forced = _WS_FORCED_TILE
    # ... and no shorter than the number of lane groups that share its rows (a plan with few, heavy rows — the
    # by-coboundary plan: ~26 messages per ring — would leave most consumer groups idle: it keeps the row kernels)
    groups = lib.cwn_csr_ws_consumer_threads() // lib.cwn_csr_ws_lanes_per_row(F)
    for tile_rows in ((forced,) if forced else (256, 128, 64, 32, 16)):
        if tile_rows < groups and not forced:
            break
        entry = plan.ws.get(tile_rows)
        if entry is None:
            if torch.cuda.is_current_stream_capturing():
                return None
            n_tiles = (plan.n_rows + tile_rows - 1) // tile_rows
            with torch.cuda.device(plan.rowptr.device):
                windows = torch.zeros(8 * n_tiles + 4, dtype=torch.int32, device=plan.rowptr.device)
                _call('csr_tile_windows', 4 * (plan.n_rows + 1) + 4 * plan.E * (1 + (plan.pay1 is not None)) + 32 * n_tiles,
                      lib.cwn_csr_tile_windows, plan.rowptr.data_ptr(), _ptr(plan.pay0), _ptr(plan.pay1), plan.n_rows,
                      tile_rows, windows.data_ptr(), _stream())
            cap0, cap1, capm, _ = windows[-4:].tolist()
            entry = plan.ws[tile_rows] = (windows, cap0, cap1, capm)
        windows, cap0, cap1, capm = entry
        if lib.cwn_csr_ws_stages(F, tile_rows, cap0, cap1 if narr == 2 else 0, capm, narr, int(rowop)) >= _WS_MIN_STAGES:
            return windows, tile_rows, cap0, cap1, capm
    return None


def build_plan(key: Tensor, n_rows: int, pay0: Tensor = None, pay1: Tensor = None) -> Plan:
    """cwn_csr_plan_build: stable CSR grouping of `key` (int64 [E], values in [0, n_rows))."""
    if not key.is_cuda:
        raise RuntimeError('cwn_b200: index tensors must live on the GPU (CUDA-only path)')
    if key.dtype != torch.long:
        raise TypeError('cwn_b200: index tensors must be torch.long')
    lib = _lib.load()
    E = key.numel()
    dev = key.device
    key = key.contiguous()
    pay0 = pay0.contiguous() if pay0 is not None else None
    pay1 = pay1.contiguous() if pay1 is not None else None
    with torch.cuda.device(dev):
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        p0 = torch.empty(E, dtype=torch.int32, device=dev) if pay0 is not None else None
        p1 = torch.empty(E, dtype=torch.int32, device=dev) if pay1 is not None else None
        ws_bytes = lib.cwn_csr_plan_workspace_bytes(E, n_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        npay = (pay0 is not None) + (pay1 is not None)
        _call('csr_plan_build', 8 * E * (1 + npay) + 4 * E * (1 + npay) + 4 * (n_rows + 1),
              lib.cwn_csr_plan_build, _ptr(key), _ptr(pay0), _ptr(pay1), E, n_rows, rowptr.data_ptr(), _ptr(perm),
              _ptr(p0), _ptr(p1), None, ws.data_ptr(), ws_bytes, _stream())
    return Plan(rowptr, perm, p0, p1, n_rows, E)


def build_plans(requests):
    """Build many plans at once. `requests`: list of (key, n_rows, pay0, pay1). Every plan small enough for
    cwn_csr_plan_build_small goes into ONE kernel launch (one CTA per plan); larger ones use the CUB path."""
    if not requests:
        return []
    lib = _lib.load()
    cap = lib.cwn_csr_plan_small_capacity()
    plans = [None] * len(requests)
    small, keep = [], []
    for i, (key, n_rows, pay0, pay1) in enumerate(requests):
        if key.numel() > cap:
            plans[i] = build_plan(key, n_rows, pay0, pay1)
            continue
        if not key.is_cuda or key.dtype != torch.long:
            raise RuntimeError('cwn_b200: index tensors must be CUDA torch.long tensors (CUDA-only path)')
        dev, E = key.device, key.numel()
        key = key.contiguous()
        pay0 = pay0.contiguous() if pay0 is not None else None
        pay1 = pay1.contiguous() if pay1 is not None else None
        with torch.cuda.device(dev):
            # one allocation per plan: [rowptr | perm | pay0 | pay1]
            npay = (pay0 is not None) + (pay1 is not None)
            buf = torch.empty(n_rows + 1 + E * (1 + npay), dtype=torch.int32, device=dev)
        rowptr, perm = buf[:n_rows + 1], buf[n_rows + 1:n_rows + 1 + E]
        off = n_rows + 1 + E
        p0 = p1 = None
        if pay0 is not None:
            p0, off = buf[off:off + E], off + E
        if pay1 is not None:
            p1 = buf[off:off + E]
        plans[i] = Plan(rowptr, perm, p0, p1, n_rows, E)
        small.append(_lib.PlanDesc(_ptr(key), _ptr(pay0), _ptr(pay1), E, n_rows, rowptr.data_ptr(), _ptr(perm),
                                   _ptr(p0), _ptr(p1)))
        keep += [key, pay0, pay1]
    if small:
        arr = (_lib.PlanDesc * len(small))(*small)
        algo = sum(12 * d.E * (1 + (d.pay0 is not None) + (d.pay1 is not None)) + 4 * (d.n_rows + 1) for d in small)
        with torch.cuda.device(requests[0][0].device):
            _call('csr_plan_build_small', algo, lib.cwn_csr_plan_build_small, arr, len(small), None, _stream())
    return plans


class Adjacency(object):
    """One adjacency of a cochain (`index` int64 [2, E]: row 0 = source, row 1 = destination; optional per-message
    coboundary/boundary id column `cob`) with its lazily built, cached CSR plans:
        by_dst : grouped by destination, payload (source, cob)   -> forward passes
        by_src : grouped by source,      payload (destination, cob) -> gradient w.r.t. the gathered operand
        by_cob : grouped by cob,         payload (destination, source) -> gradient w.r.t. the coboundary operand
    Index tensors are constant across layers and between forward and backward, so each plan is built once per
    batch. The cache lives on the index tensor object itself and is keyed by its version counter."""

    def __init__(self, index: Tensor, n_src: int, n_dst: int, cob: Tensor = None, n_cob: int = None):
        if index.dim() != 2 or index.size(0) != 2:
            raise ValueError('cwn_b200: adjacency index must have shape [2, num_messages]')
        self.index, self.n_src, self.n_dst, self.cob, self.n_cob = index, int(n_src), int(n_dst), cob, n_cob
        self.E = index.size(1)

    @classmethod
    def of(cls, index: Tensor, n_src: int, n_dst: int, cob: Tensor = None, n_cob: int = None):
        cache = index.__dict__.setdefault('_cwn_adj', {})
        gen = (index._version, index.data_ptr())
        if cache.get('gen') != gen:  # first use, or the index was modified in place: drop stale plans
            cache.clear()
            cache['gen'] = gen
        k = (int(n_src), int(n_dst), None if cob is None else (cob.data_ptr(), cob._version), n_cob)
        adj = cache.get(k)
        if adj is None:
            adj = cache[k] = cls(index, n_src, n_dst, cob, n_cob)
            adj._plans = {}
        return adj

    def request(self, kind):
        """(key, n_rows, pay0, pay1) of plan `kind`."""
        src, dst = self.index[0], self.index[1]
        if kind == 'by_dst':
            return dst, self.n_dst, src, self.cob
        if kind == 'by_src':
            return src, self.n_src, dst, self.cob
        return self.cob, self.n_cob, dst, src

    def _plan(self, kind):
        plan = self._plans.get(kind)
        if plan is None:
            plan = self._plans[kind] = build_plans([self.request(kind)])[0]
        return plan

    by_dst = property(lambda self: self._plan('by_dst'))
    by_src = property(lambda self: self._plan('by_src'))
    by_cob = property(lambda self: self._plan('by_cob'))


def _row_plan(idx: Tensor, n_rows: int, install: Plan = None) -> Plan:
    """Plan grouping the entries of a 1-D index (readout `batch` vector, gather indices) by value; cached on it."""
    cache = idx.__dict__.setdefault('_cwn_rowplan', {})
    gen = (idx._version, idx.data_ptr())
    if cache.get('gen') != gen:
        cache.clear()
        cache['gen'] = gen
    if install is not None:
        cache[n_rows] = install
    plan = cache.get(n_rows)
    if plan is None:
        plan = cache[n_rows] = build_plans([(idx, n_rows, None, None)])[0]
    return plan


def prepare_plans(data, max_dim: int = 2, use_coboundaries: bool = False, backward: bool = True,
                  readout: bool = True):
    """Build, in ONE kernel launch, every CSR plan a SparseCIN-family forward (+ backward) pass over `data` will
    ask for: per dimension the upper adjacency (by destination; by source and by coboundary for the gradients),
    the boundary adjacency (by destination / by source) and the readout grouping of `batch`. Plans that are
    already cached are skipped; anything not prepared here is still built lazily on first use."""
    wanted = []  # (installer, request)
    top = min(max_dim, data.dimension)
    for d in range(top + 1):
        c = data.cochains[d]
        n = c.num_cells
        if c.upper_index is not None and (d + 1) in data.cochains and c.upper_index.is_cuda:
            cob, n_cob = None, None
            if use_coboundaries and c.shared_coboundaries is not None:
                cob, n_cob = c.shared_coboundaries, data.cochains[d + 1].num_cells
            adj = Adjacency.of(c.upper_index, n, n, cob, n_cob)
            kinds = ['by_dst'] + (['by_src'] + (['by_cob'] if cob is not None else []) if backward else [])
            wanted += [(adj, k) for k in kinds if k not in adj._plans]
        if d > 0 and c.boundary_index is not None and c.boundary_index.is_cuda:
            adj = Adjacency.of(c.boundary_index, data.cochains[d - 1].num_cells, n)
            wanted += [(adj, k) for k in (['by_dst', 'by_src'] if backward else ['by_dst']) if k not in adj._plans]
    reqs = [adj.request(k) for adj, k in wanted]
    pools = []
    if readout:
        size = getattr(data, 'num_complexes', None)
        for d in range(top + 1):
            b = data.cochains[d].batch
            if size is not None and b is not None and b.is_cuda:
                cache = b.__dict__.get('_cwn_rowplan', {})
                if cache.get('gen') != (b._version, b.data_ptr()) or int(size) not in cache:
                    pools.append(b)
                    reqs.append((b, int(size), None, None))
    plans = build_plans(reqs)
    for (adj, k), plan in zip(wanted, plans):
        adj._plans[k] = plan
    for b, plan in zip(pools, plans[len(wanted):]):
        _row_plan(b, plan.n_rows, install=plan)
    return len(plans)


def clear_plan_cache(*indices):
    for idx in indices:
        if idx is not None:
            idx.__dict__.pop('_cwn_adj', None)
            idx.__dict__.pop('_cwn_rowplan', None)
            idx.__dict__.pop('_cwn_splitkey', None)


# --------------------------------------------------------------------------------------------- raw launches
def _launch_gather_reduce(x_src, plan_rowptr, idx, n_rows, F, x_res, eps, reduce_code, plan: Plan = None):
    """Algorithmic bytes: 16 B of int64 index per message + one read of every source row + one write
    of every destination row (+ one read of the residual rows when fused)."""
    lib = _lib.load()
    dev = plan_rowptr.device
    E = idx.numel() if idx is not None else (x_src.size(0) if x_src is not None else 0)
    n_src = x_src.size(0) if x_src is not None else 0
    if _profile is not None and idx is not None and x_src is not None:
        # only the source rows some message reads are compulsory traffic (a ring-boundary pass touches the 17 of 25
        # edges that lie on a ring; charging all rows made that pass read 1.4x the HBM peak). Counted once per plan,
        # under profiling only (it synchronises).
        cached = idx.__dict__.get('_cwn_touched')
        if cached is None:
            cached = idx.__dict__['_cwn_touched'] = int(torch.unique(idx).numel())
        n_src = min(n_src, cached)
    algo = 16 * E + 4 * F * (n_src + n_rows + (n_rows if x_res is not None else 0))
    ref = x_src if x_src is not None else x_res
    if ref is not None and _f64(ref):
        with torch.cuda.device(dev):
            out = torch.empty(n_rows, F, dtype=torch.float64, device=dev)
            _call('csr_gather_reduce', 2 * algo - 16 * E, lib.cwn_csr_gather_reduce_f64,
                  _ptr(x_src), _ld(x_src) if x_src is not None else F, plan_rowptr.data_ptr(), _ptr(idx), n_rows, F,
                  _ptr(x_res), _ld(x_res) if x_res is not None else F, _ptr(eps), _ptr(out), F, reduce_code, _stream())
        return out
    ws = None
    if plan is not None and idx is plan.pay0 and reduce_code in (0, 1) and x_src is not None and _ws_ok(x_src, x_res):
        ws = _ws_config(plan, F, 1, x_res is not None)
    with torch.cuda.device(dev):
        out = torch.empty(n_rows, F, dtype=torch.float32, device=dev)
        if ws is not None:
            windows, tile_rows, cap0, _, capm = ws
            _call('csr_gather_reduce', algo, lib.cwn_csr_gather_reduce_ws_f32,
                  _ptr(x_src), _ld(x_src), plan_rowptr.data_ptr(), _ptr(idx), E, windows.data_ptr(), tile_rows, cap0,
                  capm, n_rows, F, _ptr(x_res), _ld(x_res) if x_res is not None else F, _ptr(eps), _ptr(out), F,
                  reduce_code, _stream())
            return out
        _call('csr_gather_reduce', algo, lib.cwn_csr_gather_reduce_f32,
              _ptr(x_src), _ld(x_src) if x_src is not None else F, plan_rowptr.data_ptr(), _ptr(idx), n_rows, F,
              _ptr(x_res), _ld(x_res) if x_res is not None else F, _ptr(eps), _ptr(out), F, reduce_code, _stream())
    return out


def _launch_gather_rows(x, idx, scale):
    lib = _lib.load()
    E, F = idx.numel(), x.size(1)
    with torch.cuda.device(x.device):
        out = torch.empty(E, F, dtype=x.dtype, device=x.device)
        if _f64(x):
            _call('gather_rows', 8 * E + 8 * F * (x.size(0) + E), lib.cwn_gather_rows_f64,
                  _ptr(x), _ld(x), _ptr(idx), E, F, float(scale), _ptr(out), F, _stream())
        else:
            _call('gather_rows', 8 * E + 4 * F * (x.size(0) + E), lib.cwn_gather_rows_f32,
                  _ptr(x), _ld(x), _ptr(idx), E, F, float(scale), _ptr(out), F, _stream())
    return out


def _residual_grads(needs_res, needs_eps, g, x_res_saved, eps):
    """(grad x_res, grad eps) of out = (1 + eps) * x_res + ... ; `x_res_saved` is only kept when eps needs a grad."""
    g_res = g_eps = None
    if needs_res:
        g_res = g if eps is None else torch.addcmul(g, g, eps)  # (1 + eps) * g in one launch
    if needs_eps and eps is not None:
        g_eps = (g * x_res_saved).sum().reshape(eps.shape)
    return g_res, g_eps


def _check_eps(eps, like=None):
    if eps is not None:
        _require_cuda_float(eps, 'eps', like)
        if eps.numel() != 1:
            raise ValueError('cwn_b200: eps must hold a single value')


# --------------------------------------------------------------------------------------------- autograd ops
class _GatherReduce(Function):
    """out[r] = (1+eps)*x_res[r] + REDUCE_{e: dst_e = r} x_src[src_e]  — the fused identity-message pass."""

    @staticmethod
    def forward(ctx, x_src, x_res, eps, adj: Adjacency, reduce: str):
        code = REDUCE_CODES[reduce]
        plan = adj.by_dst
        F = x_src.size(1)
        arg = None
        if reduce == 'max' and ctx.needs_input_grad[0] and adj.E > 0 and _f64(x_src):
            raise NotImplementedError("cwn_b200: the backward of reduce='max' is float32-only (no shipped model uses it)")
        if reduce == 'max' and ctx.needs_input_grad[0] and adj.E > 0:
            # the backward needs to know WHICH message won every (row, feature): a scalar kernel that tracks it
            with torch.cuda.device(x_src.device):
                out = torch.empty(adj.n_dst, F, dtype=torch.float32, device=x_src.device)
                arg = torch.empty(adj.n_dst, F, dtype=torch.int32, device=x_src.device)
                _call('csr_gather_max_arg', 16 * adj.E + 4 * F * (adj.n_src + 2 * adj.n_dst),
                      _lib.load().cwn_csr_gather_max_arg_f32, _ptr(x_src), _ld(x_src), plan.rowptr.data_ptr(),
                      _ptr(plan.pay0), _ptr(plan.perm), adj.n_dst, F, _ptr(out), F, _ptr(arg), _stream())
        else:
            out = _launch_gather_reduce(x_src, plan.rowptr, plan.pay0, adj.n_dst, F, x_res, eps, code, plan=plan)
        ctx.adj, ctx.reduce = adj, reduce
        ctx.save_for_backward(x_res if (eps is not None and eps.requires_grad) else None, eps, arg)
        ctx.has_res = x_res is not None
        return out

    @staticmethod
    def backward(ctx, g):
        adj = ctx.adj
        x_res, eps, arg = ctx.saved_tensors
        g = _rows(g)
        g_src = None
        if ctx.needs_input_grad[0] and ctx.reduce == 'max':
            plan = adj.by_src
            with torch.cuda.device(g.device):
                g_src = torch.zeros(adj.n_src, g.size(1), dtype=torch.float32, device=g.device)
                if arg is not None:
                    _call('csr_max_bwd', 16 * adj.E + 4 * g.size(1) * (2 * adj.n_dst + adj.n_src),
                          _lib.load().cwn_csr_max_bwd_f32, _ptr(g), _ld(g), _ptr(arg), plan.rowptr.data_ptr(),
                          _ptr(plan.pay0), _ptr(plan.perm), adj.n_src, g.size(1), _ptr(g_src), g.size(1), _stream())
        elif ctx.needs_input_grad[0]:
            gm = g
            if ctx.reduce == 'mean':
                deg = (adj.by_dst.rowptr[1:] - adj.by_dst.rowptr[:-1]).clamp_(min=1).to(g.dtype)
                gm = g / deg.unsqueeze(-1)
            plan = adj.by_src  # transposed pass: gX[s] = sum_{e: src_e = s} G[dst_e]
            g_src = _launch_gather_reduce(gm, plan.rowptr, plan.pay0, adj.n_src, g.size(1), None, None, 0, plan=plan)
        g_res = g_eps = None
        if ctx.has_res:
            g_res, g_eps = _residual_grads(ctx.needs_input_grad[1], ctx.needs_input_grad[2], g, x_res, eps)
        return g_src, g_res, g_eps, None, None


def _launch_cob_bwd(algo, g, A, B, plan, E, n_rows, F, act, gA):
    lib = _lib.load()
    if _f64(A):
        g = g if _f64(g) else g.double()
        _call('csr_cob_bwd', 2 * algo - 24 * E, lib.cwn_csr_cob_bwd_f64,
              _ptr(g), _ld(g), _ptr(A), _ld(A), _ptr(B), _ld(B), plan.rowptr.data_ptr(), _ptr(plan.pay0),
              _ptr(plan.pay1), n_rows, F, act, _ptr(gA), F, _stream())
        return
    ws = _ws_config(plan, F, 2, True) if _ws_ok(g, A, B) else None
    if ws is not None:
        windows, tile_rows, cap0, cap1, capm = ws
        _call('csr_cob_bwd', algo, lib.cwn_csr_cob_bwd_ws_f32,
              _ptr(g), _ld(g), _ptr(A), _ld(A), _ptr(B), _ld(B), plan.rowptr.data_ptr(), _ptr(plan.pay0),
              _ptr(plan.pay1), E, windows.data_ptr(), tile_rows, cap0, cap1, capm, n_rows, F, act, _ptr(gA), F,
              _stream())
    else:
        _call('csr_cob_bwd', algo, lib.cwn_csr_cob_bwd_f32,
              _ptr(g), _ld(g), _ptr(A), _ld(A), _ptr(B), _ld(B), plan.rowptr.data_ptr(), _ptr(plan.pay0),
              _ptr(plan.pay1), n_rows, F, act, _ptr(gA), F, _stream())


class _CobPass(Function):
    """out[r] = (1+eps)*x_res[r] + SUM_{e: dst_e = r} act(P[src_e] + Q[cob_e]) — upper pass with coboundaries."""

    @staticmethod
    def forward(ctx, P, Q, x_res, eps, adj: Adjacency, act: str):
        lib = _lib.load()
        plan = adj.by_dst
        F = P.size(1)
        with torch.cuda.device(P.device):
            out = torch.empty(adj.n_dst, F, dtype=P.dtype, device=P.device)
            algo = 24 * adj.E + 4 * F * (P.size(0) + Q.size(0) + adj.n_dst * (2 if x_res is not None else 1))
            ws = _ws_config(plan, F, 2, x_res is not None) if (not _f64(P) and _ws_ok(P, Q, x_res)) else None
            if _f64(P):
                _call('csr_cob_fwd', 2 * algo - 24 * adj.E, lib.cwn_csr_cob_fwd_f64,
                      _ptr(P), _ld(P), _ptr(Q), _ld(Q), plan.rowptr.data_ptr(), _ptr(plan.pay0), _ptr(plan.pay1),
                      adj.n_dst, F, ACT_CODES[act], _ptr(x_res), _ld(x_res) if x_res is not None else F, _ptr(eps),
                      _ptr(out), F, _stream())
            elif ws is not None:
                windows, tile_rows, cap0, cap1, capm = ws
                _call('csr_cob_fwd', algo, lib.cwn_csr_cob_fwd_ws_f32,
                      _ptr(P), _ld(P), _ptr(Q), _ld(Q), plan.rowptr.data_ptr(), _ptr(plan.pay0), _ptr(plan.pay1), adj.E,
                      windows.data_ptr(), tile_rows, cap0, cap1, capm, adj.n_dst, F, ACT_CODES[act], _ptr(x_res),
                      _ld(x_res) if x_res is not None else F, _ptr(eps), _ptr(out), F, _stream())
            else:
                _call('csr_cob_fwd', algo, lib.cwn_csr_cob_fwd_f32,
                      _ptr(P), _ld(P), _ptr(Q), _ld(Q), plan.rowptr.data_ptr(), _ptr(plan.pay0), _ptr(plan.pay1),
                      adj.n_dst, F, ACT_CODES[act], _ptr(x_res), _ld(x_res) if x_res is not None else F, _ptr(eps),
                      _ptr(out), F, _stream())
        ctx.adj, ctx.act = adj, act
        ctx.has_res = x_res is not None
        ctx.save_for_backward(P, Q, x_res if (eps is not None and eps.requires_grad) else None, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        adj, act = ctx.adj, ACT_CODES[ctx.act]
        P, Q, x_res, eps = ctx.saved_tensors
        g = _rows(g)
        F = P.size(1)
        gP = gQ = None
        with torch.cuda.device(P.device):
            if ctx.needs_input_grad[0]:
                plan = adj.by_src
                gP = torch.empty_like(P, memory_format=torch.contiguous_format)
                algo = 24 * adj.E + 4 * F * (g.size(0) + 2 * P.size(0) + Q.size(0))
                _launch_cob_bwd(algo, g, P, Q, plan, adj.E, adj.n_src, F, act, gP)
            if ctx.needs_input_grad[1]:
                plan = adj.by_cob
                gQ = torch.empty_like(Q, memory_format=torch.contiguous_format)
                algo = 24 * adj.E + 4 * F * (g.size(0) + 2 * Q.size(0) + P.size(0))
                _launch_cob_bwd(algo, g, Q, P, plan, adj.E, adj.n_cob, F, act, gQ)
        g_res = g_eps = None
        if ctx.has_res:
            g_res, g_eps = _residual_grads(ctx.needs_input_grad[2], ctx.needs_input_grad[3], g, x_res, eps)
        return gP, gQ, g_res, g_eps, None, None


class _CinMsgPass(Function):
    """out[t] = SUM_{e -> t} BN(act(P[src_e] + Q[att_e])) with BatchNorm over the E messages (K5; reference
    mp/layers.py:94-103 with the nets of mp/models.py:40-47). See csrc/cin_msg.cu for the algebra. `stats` = None for
    batch statistics (training) or (running_mean, running_var) for eval. Returns (out, mean, var) — mean / var (detached)
    are what the caller feeds the module's running statistics with."""

    @staticmethod
    def forward(ctx, P, Q, gamma, beta, adj: Adjacency, act: str, bn_eps: float, stats):
        lib = _lib.load()
        plan = adj.by_dst
        F, E, dev = P.size(1), adj.E, P.device
        code = ACT_CODES[act]
        with torch.cuda.device(dev):
            S = torch.empty(adj.n_dst, F, dtype=torch.float32, device=dev)
            algo = 24 * E + 4 * F * (P.size(0) + Q.size(0) + adj.n_dst)
            _call('csr_cob_fwd', algo, lib.cwn_csr_cob_fwd_f32, _ptr(P), _ld(P), _ptr(Q), _ld(Q), plan.rowptr.data_ptr(),
                  _ptr(plan.pay0), _ptr(plan.pay1), adj.n_dst, F, code, None, F, None, _ptr(S), F, _stream())
            deg = (plan.rowptr[1:] - plan.rowptr[:-1]).to(torch.float32).unsqueeze(1)
            if stats is None:
                mean = S.sum(0) / E
                S2 = torch.empty_like(S)
                _call('cin_msg_sq', algo, lib.cwn_cin_msg_sq_f32, _ptr(P), _ld(P), _ptr(Q), _ld(Q), plan.rowptr.data_ptr(),
                      _ptr(plan.pay0), _ptr(plan.pay1), adj.n_dst, F, code, _ptr(mean), _ptr(S2), F, _stream())
                var = S2.sum(0) / E
            else:
                mean, var = stats[0].to(torch.float32), stats[1].to(torch.float32)
            rstd = torch.rsqrt(var + bn_eps)
            scale = rstd * gamma if gamma is not None else rstd
            shift = (beta if beta is not None else 0) - scale * mean
            out = S * scale + deg * shift
        ctx.adj, ctx.act, ctx.training, ctx.E = adj, code, stats is None, E
        ctx.has_affine = gamma is not None
        ctx.save_for_backward(P, Q, S, deg, mean.contiguous(), rstd.contiguous(), scale.contiguous())
        ctx.mark_non_differentiable(mean, var)
        return out, mean, var

    @staticmethod
    def backward(ctx, g, _gm, _gv):
        lib = _lib.load()
        adj, code, E = ctx.adj, ctx.act, ctx.E
        P, Q, S, deg, mean, rstd, scale = ctx.saved_tensors
        g = _rows(g)
        F, dev = P.size(1), P.device
        gP = gQ = g_gamma = g_beta = None
        with torch.cuda.device(dev):
            sum_g = (deg * g).sum(0)                       # SUM_e G_e
            sum_ga = (g * (S - deg * mean)).sum(0) * rstd  # SUM_e G_e ahat_e
            if ctx.training:
                c1, c2 = (sum_g / E).contiguous(), (sum_ga / E).contiguous()
            else:
                c1 = c2 = torch.zeros(F, dtype=torch.float32, device=dev)
            for want, plan, A, B, n_rows in ((ctx.needs_input_grad[0], adj.by_src, P, Q, adj.n_src),
                                             (ctx.needs_input_grad[1], adj.by_cob, Q, P, adj.n_cob)):
                if not want:
                    continue
                gA = torch.empty_like(A, memory_format=torch.contiguous_format)
                algo = 24 * E + 4 * F * (g.size(0) + 2 * A.size(0) + B.size(0))
                _call('cin_msg_bwd', algo, lib.cwn_cin_msg_bwd_f32, _ptr(g), _ld(g), _ptr(A), _ld(A), _ptr(B), _ld(B),
                      plan.rowptr.data_ptr(), _ptr(plan.pay0), _ptr(plan.pay1), n_rows, F, code, _ptr(scale), _ptr(mean),
                      _ptr(rstd), _ptr(c1), _ptr(c2), _ptr(gA), F, _stream())
                if A is P:
                    gP = gA
                else:
                    gQ = gA
            if ctx.has_affine:
                g_gamma = sum_ga if ctx.needs_input_grad[2] else None
                g_beta = sum_g if ctx.needs_input_grad[3] else None
        return gP, gQ, g_gamma, g_beta, None, None, None, None


def cin_message_pass(P: Tensor, Q: Tensor, index: Tensor, att: Tensor, n_dst: int, act: str, bn=None) -> Tensor:
    """SUM over the messages into each destination of BN(act(P[src] + Q[att])), BatchNorm over the MESSAGE population
    (`bn`: torch BatchNorm1d or None; its running statistics are updated in training mode as torch does)."""
    if act not in ACT_CODES:
        raise ValueError(f'cwn_b200: unknown activation {act!r}')
    P, Q = _rows(P), _rows(Q)
    _require_cuda_float(P, 'P')
    _require_cuda_float(Q, 'Q', P)
    adj = Adjacency.of(index, P.size(0), n_dst, att, Q.size(0))
    if bn is None:
        return _CobPass.apply(P, Q, None, None, adj, act)
    use_batch = bn.training or not bn.track_running_stats
    stats = None if use_batch else (bn.running_mean, bn.running_var)
    out, mean, var = _CinMsgPass.apply(P, Q, bn.weight, bn.bias, adj, act, float(bn.eps), stats)
    if bn.training and bn.track_running_stats:
        with torch.no_grad():
            E = adj.E
            bn.num_batches_tracked += 1
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var * (E / max(E - 1, 1)), alpha=m)
    return out


def _rows_grad_plan(idx, n):
    """(plan, split) for the gradient of `x[idx]` w.r.t. x: rows of x grouped by `idx`. Few, very long rows (an embedding
    table): one thread group per row would serialise, so every row is split into `split` interleaved sub-rows (a
    deterministic two-level segmented sum). Cached on `idx`."""
    E = idx.numel()
    split = 1
    if E > 32 * max(n, 1):
        split = 2
        while split < 256 and split * 8 * n < E:
            split *= 2
    if split == 1:
        return _row_plan(idx, n), 1
    cache = idx.__dict__.setdefault('_cwn_splitkey', {})
    gen = (idx._version, idx.data_ptr(), split)
    if cache.get('gen') != gen:
        cache.clear()
        cache['gen'] = gen
        lanes = torch.arange(E, device=idx.device).bitwise_and_(split - 1)
        cache['key'] = torch.add(lanes, idx, alpha=split)
    return _row_plan(cache['key'], n * split), split


class _GatherRows(Function):
    """out[e] = scale * x[idx[e]]"""

    @staticmethod
    def forward(ctx, x, idx, scale):
        ctx.idx, ctx.n, ctx.scale = idx, x.size(0), scale
        ctx.plan_ready = None
        if ctx.needs_input_grad[0] and idx.is_cuda and idx.numel() > 0 and _streams.ENABLED:
            # The backward needs a CSR plan of `idx` (and, for embedding tables, of its split key): 3 tiny torch
            # launches + a plan build that would sit at the very END of the step's critical path. Build it now on a
            # dedicated side stream — it overlaps the forward pass — and let the backward wait on the event.
            cur = torch.cuda.current_stream(idx.device)
            side = _streams.aux_stream(cur)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                _rows_grad_plan(idx, ctx.n)
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
        return _launch_gather_rows(x, idx, scale)

    @staticmethod
    def backward(ctx, g):
        g = _rows(g)
        idx, n = ctx.idx, ctx.n
        if ctx.plan_ready is not None:
            torch.cuda.current_stream(g.device).wait_event(ctx.plan_ready)
        plan, split = _rows_grad_plan(idx, n)
        if split == 1:
            gx = _launch_gather_reduce(g, plan.rowptr, plan.perm, n, g.size(1), None, None, 0)
        else:
            part = _launch_gather_reduce(g, plan.rowptr, plan.perm, n * split, g.size(1), None, None, 0)
            gx = part.view(n, split, g.size(1)).sum(dim=1)
        if ctx.scale != 1.0:
            gx = gx * ctx.scale
        return gx, None, None


class _ScatterRows(Function):
    """out[r] = REDUCE_{e: dst[e] = r} msg[e]  (aggregation of messages materialised by a user hook, or readout)."""

    @staticmethod
    def forward(ctx, msg, dst, n_dst, reduce):
        plan = _row_plan(dst, n_dst)
        ctx.dst, ctx.plan, ctx.reduce = dst, plan, reduce
        ctx.arg = None
        if reduce == 'max' and ctx.needs_input_grad[0] and msg.size(0) > 0 and _f64(msg):
            raise NotImplementedError("cwn_b200: the backward of reduce='max' is float32-only (no shipped model uses it)")
        if reduce == 'max' and ctx.needs_input_grad[0] and msg.size(0) > 0:
            F = msg.size(1)
            with torch.cuda.device(msg.device):
                out = torch.empty(n_dst, F, dtype=torch.float32, device=msg.device)
                ctx.arg = torch.empty(n_dst, F, dtype=torch.int32, device=msg.device)
                _call('csr_gather_max_arg', 16 * msg.size(0) + 4 * F * (msg.size(0) + 2 * n_dst),
                      _lib.load().cwn_csr_gather_max_arg_f32, _ptr(msg), _ld(msg), plan.rowptr.data_ptr(),
                      _ptr(plan.perm), _ptr(plan.perm), n_dst, F, _ptr(out), F, _ptr(ctx.arg), _stream())
            return out
        return _launch_gather_reduce(msg, plan.rowptr, plan.perm, n_dst, msg.size(1), None, None,
                                     REDUCE_CODES[reduce])

    @staticmethod
    def backward(ctx, g):
        g = _rows(g)
        if ctx.reduce == 'max':
            E, F = ctx.dst.numel(), g.size(1)
            with torch.cuda.device(g.device):
                g_msg = torch.zeros(E, F, dtype=torch.float32, device=g.device)
                if ctx.arg is not None:
                    # every message is its own row: g_msg[e] = (arg[dst[e]] == e) ? g[dst[e]] : 0
                    rowptr = torch.arange(E + 1, dtype=torch.int32, device=g.device)
                    dst32 = ctx.dst.to(torch.int32)
                    _call('csr_max_bwd', 8 * E + 4 * F * (2 * g.size(0) + E), _lib.load().cwn_csr_max_bwd_f32,
                          _ptr(g), _ld(g), _ptr(ctx.arg), rowptr.data_ptr(), _ptr(dst32), None, E, F, _ptr(g_msg), F,
                          _stream())
            return g_msg, None, None, None
        if ctx.reduce == 'mean':
            deg = (ctx.plan.rowptr[1:] - ctx.plan.rowptr[:-1]).clamp_(min=1).to(g.dtype)
            g = g / deg.unsqueeze(-1)
        return _launch_gather_rows(g, ctx.dst, 1.0), None, None, None


# --------------------------------------------------------------------------------------------- public API
def gather_scatter(x_src: Tensor, index: Tensor, n_dst: int, reduce: str = 'add', x_res: Tensor = None,
                   eps: Tensor = None) -> Tensor:
    """Fused `scatter(x_src.index_select(0, index[0]), index[1], dim_size=n_dst, reduce)` [+ (1+eps) * x_res]."""
    if reduce not in REDUCE_CODES:
        raise ValueError(f'cwn_b200: unknown aggregation {reduce!r}')
    x_src = _rows(x_src)
    _require_cuda_float(x_src, 'x_src')
    if x_res is not None:
        if reduce not in ('add', 'sum'):
            # (the kernel fuses the residual into an additive pass only; refusing here keeps the behaviour the same with
            # and without autograd — the max path with gradients used to drop the residual silently)
            raise ValueError("cwn_b200: a fused residual (x_res) requires reduce='add'")
        x_res = _rows(x_res)
        _require_cuda_float(x_res, 'x_res', x_src)
        if x_res.size(0) != n_dst or x_res.size(1) != x_src.size(1):
            raise ValueError('cwn_b200: residual operand must have shape [n_dst, F]')
    _check_eps(eps, x_src)
    adj = Adjacency.of(index, x_src.size(0), n_dst)
    return _GatherReduce.apply(x_src, x_res, eps, adj, reduce)


def cob_pass(P: Tensor, Q: Tensor, index: Tensor, cob: Tensor, n_dst: int, act: str = 'relu',
             x_res: Tensor = None, eps: Tensor = None) -> Tensor:
    """Upper-adjacency pass with coboundary features in split-weight form (see include/cwn_b200.h)."""
    if act not in ACT_CODES:
        raise ValueError(f'cwn_b200: unknown activation {act!r}')
    P, Q = _rows(P), _rows(Q)
    _require_cuda_float(P, 'P')
    _require_cuda_float(Q, 'Q', P)
    if P.size(1) != Q.size(1):
        raise ValueError('cwn_b200: P and Q must have the same width')
    if x_res is not None:
        x_res = _rows(x_res)
        _require_cuda_float(x_res, 'x_res', P)
    _check_eps(eps, P)
    if cob.numel() != index.size(1):
        raise ValueError('cwn_b200: one coboundary id per message is required')
    adj = Adjacency.of(index, P.size(0), n_dst, cob, Q.size(0))
    return _CobPass.apply(P, Q, x_res, eps, adj, act)


def gather_rows(x: Tensor, idx: Tensor, scale: float = 1.0) -> Tensor:
    """`scale * x.index_select(0, idx)` (reference `__lift__`, mp/cell_mp.py:195-198)."""
    x = _rows(x)
    _require_cuda_float(x, 'x')
    if idx.dtype != torch.long or not idx.is_cuda:
        raise TypeError('cwn_b200: gather index must be a CUDA torch.long tensor')
    return _GatherRows.apply(x, idx.contiguous(), float(scale))


def scatter_rows(msg: Tensor, dst: Tensor, n_dst: int, reduce: str = 'add') -> Tensor:
    """`torch_scatter.scatter(msg, dst, dim=0, dim_size=n_dst, reduce)` (mp/cell_mp.py:439-440)."""
    if reduce not in REDUCE_CODES:
        raise ValueError(f'cwn_b200: unknown aggregation {reduce!r}')
    msg = _rows(msg)
    _require_cuda_float(msg, 'messages')
    if dst.dtype != torch.long or not dst.is_cuda:
        raise TypeError('cwn_b200: scatter index must be a CUDA torch.long tensor')
    if dst.numel() != msg.size(0):
        raise ValueError('cwn_b200: one destination per message is required')
    return _ScatterRows.apply(msg, dst.contiguous(), int(n_dst), reduce)


def segment_pool(x: Tensor, batch: Tensor, size: int, mean: bool = False) -> Tensor:
    """Per-complex readout `global_add_pool / global_mean_pool(x, batch, size)` (mp/nn.py:59)."""
    return scatter_rows(x, batch, size, 'mean' if mean else 'add')
