"""The dense update / combine nets of one `SparseCINConv` layer — all cochain dimensions, both branches — as ONE
autograd node over the grouped kernels of `csrc/dense.cu`.

Reference semantics (mp/layers.py:191-199 with the default nets of :303-325), per dimension d:
    a_up  = act(BN(Linear(act(BN(Linear(u_d))))))          update_up_nn
    a_bnd = act(BN(Linear(act(BN(Linear(b_d))))))          update_boundaries_nn
    out_d = act(BN(Linear(cat[a_up, a_bnd])))              combine_nn
where u_d / b_d are the aggregated upper / boundary messages with the GIN residual already added. PyTorch runs this
as ~40 library kernels per dimension per layer (plus as many again in backward); here a layer is 7 launches forward
(3 grouped GEMM+statistics, 3 grouped BatchNorm finalisations, 1 grouped BatchNorm+activation) and 11 backward,
BatchNorm in training mode with exact batch statistics, running statistics updated as torch does.

Supported closed form: `graph_norm` BatchNorm1d (affine, tracking running statistics) or Identity, activation in
{relu, elu, id, sigmoid, tanh}. Anything else (LayerNorm, user-supplied nets, eval-mode backward, a cochain with a
single cell — where torch's BatchNorm raises) makes `recognise()` / `applicable()` return None / False and the
layer runs through its torch modules instead.
"""
import ctypes
import os

import torch
from torch.autograd import Function
from torch.nn import BatchNorm1d, Identity, Linear, Sequential

from cwn_b200 import _lib, ops
from cwn_b200.mp.nn import activation_name

TM = 64  # default row tile of csrc/dense.cu


def _tile_rows(row_counts):
    """Rows per CTA tile for one grouped launch: 32 only when 64-row tiles would leave most of the 148 SMs without a
    CTA. (At the real-data shape — 6 problems of ~3 000 rows — 64-row tiles measured 5 % faster per training step
    than 32-row ones: half the CTAs re-reading the weight tile from L2, half the per-tile BatchNorm records for the
    last CTA to merge, and the backward kernel fits one wave.)"""
    if _FORCED_TILE_ROWS:
        return _FORCED_TILE_ROWS
    if _TC5:
        return 64  # the tensor-core kernels (csrc/dense_tc5.cuh) own 64-row tiles; other tile heights take the FFMA path
    return 32 if sum((n + 63) // 64 for n in row_counts) < _TILE32_BELOW else 64


_TILE32_BELOW = int(os.environ.get('CWN_B200_TILE32_BELOW', '96'))  # A/B switch for profiling
_FORCED_TILE_ROWS = int(os.environ.get('CWN_B200_TILE_ROWS', '0'))  # A/B switch for profiling (32 or 64)
_TC5 = os.environ.get('CWN_B200_DENSE_TC5', '1') != '0'  # tcgen05 dense kernels (default); 0 = FFMA kernels
# BatchNorm-backward reductions of a unit taken from the g_in tiles of the unit downstream (no launch of their own); 0 = A/B
_FUSE_REDUCE = os.environ.get('CWN_B200_FUSE_REDUCE', '1') != '0'


# Live row counts of a fixed-capacity (padded) batch: one int32 DEVICE scalar per cochain dimension, or None.
# cwn_b200.bucketed installs them around the model call (`with live_rows(...)`); the fused dense node hands them to
# the kernels so that BatchNorm statistics and its backward see the real cells only (cwn_linear_desc::n_rows_live).
_LIVE = None


class live_rows(object):
    def __init__(self, per_dimension):
        self.per_dimension = per_dimension

    def __enter__(self):
        global _LIVE
        self._saved, _LIVE = _LIVE, self.per_dimension
        return self

    def __exit__(self, *exc):
        global _LIVE
        _LIVE = self._saved
        return False


def padded_batch_needs_fused_path(what):
    """Called where a layer leaves the fused dense node: with live row counts installed (a padded batch) the torch
    modules would put the padding rows into the BatchNorm statistics, silently."""
    if _LIVE is not None:
        raise RuntimeError(f'cwn_b200: padded batches (cwn_b200.bucketed) need the fused dense path; {what} is outside it')


def _live(d):
    if _LIVE is None or d >= len(_LIVE):
        return None
    return _LIVE[d]


class _Unit(object):
    """Linear (+ optional BatchNorm) (+ activation) as the kernels see it."""
    __slots__ = ('lin', 'bn', 'act')

    def __init__(self, lin, bn, act):
        self.lin, self.bn, self.act = lin, bn, act


def _parse(seq, n_units):
    """Sequential(Linear, norm, act[, Linear, norm, act]) -> list of _Unit, or None if it is something else."""
    if not isinstance(seq, Sequential) or len(seq) != 3 * n_units:
        return None
    units = []
    for i in range(n_units):
        lin, norm, act = seq[3 * i], seq[3 * i + 1], seq[3 * i + 2]
        if not isinstance(lin, Linear) or lin.bias is None:
            return None
        if isinstance(norm, BatchNorm1d):
            if not (norm.affine and norm.track_running_stats and norm.momentum is not None):
                return None
        elif not isinstance(norm, Identity):
            return None
        name = activation_name(act)
        if name is None:
            return None
        units.append(_Unit(lin, norm if isinstance(norm, BatchNorm1d) else None, name))
    return units


def recognise(level):
    """(up units, boundary units, combine unit, down units or None) of a SparseCINCochainConv / CINppCochainConv built
    with the default nets, else None. CIN++ (reference mp/layers.py:216-260) has a third update branch, `update_down_nn`,
    and a combine over three blocks [up | down | boundaries]."""
    up = _parse(level.update_up_nn, 2)
    bnd = _parse(level.update_boundaries_nn, 2)
    comb = _parse(level.combine_nn, 1)
    down = _parse(level.update_down_nn, 2) if hasattr(level, 'update_down_nn') else None
    if up is None or bnd is None or comb is None or (hasattr(level, 'update_down_nn') and down is None):
        return None
    branches = [up, bnd] + ([down] if down is not None else [])
    if len({u.act for br in branches for u in br}) != 1:
        return None
    h = up[1].lin.out_features
    if any(br[1].lin.out_features != h or br[0].lin.out_features != br[1].lin.in_features for br in branches):
        return None
    if comb[0].lin.in_features != len(branches) * h:
        return None
    if max(u.lin.out_features for br in branches + [comb] for u in br) > 128:
        return None  # shared-memory budget of unit_bwd_kernel
    return up, bnd, comb[0], down


def applicable(forms, us, bs, training, ds=None):
    if any(f is None for f in forms):
        return False
    if any((f[3] is not None) != (ds is not None) for f in forms):
        return False
    params = _parameters(forms)
    if len({id(p) for p in params}) != len(params):
        return False  # nets shared between dimensions (passed_update_*_nn): gradients would collide in one launch
    for d, (f, u, b) in enumerate(zip(forms, us, bs)):
        if not (u.is_cuda and u.dtype == torch.float32 and u.dim() == 2 and b.shape == u.shape):
            return False
        if u.size(1) != f[0][0].lin.in_features or b.size(1) != f[1][0].lin.in_features:
            return False
        if ds is not None and (ds[d].shape != u.shape or ds[d].dtype != torch.float32 or not ds[d].is_cuda
                               or ds[d].size(1) != f[3][0].lin.in_features):
            return False
        has_bn = any(unit.bn is not None for unit in f[0] + f[1] + [f[2]] + (f[3] or []))
        if has_bn and training and u.size(0) < 2:
            return False  # torch raises "Expected more than 1 value per channel"; keep that behaviour
        if has_bn and not training and torch.is_grad_enabled():
            return False  # eval-mode BatchNorm backward: rare, leave it to torch
    return True


def _p(t):
    return None if t is None else t.data_ptr()


_counter_rings = {}


def _counters(device, n):
    """`n` zero int32 counters for the last-CTA hand-offs of one grouped launch. They come from a per-device ring
    and are reset to zero by the kernel that uses them, so no memset is ever launched."""
    ring = _counter_rings.get(device)
    if ring is None:
        ring = _counter_rings[device] = [torch.zeros(8192, dtype=torch.int32, device=device), 0]
    buf, pos = ring
    if pos + n > buf.numel():
        pos = 0
    ring[1] = pos + n
    return buf[pos:pos + n]


def _direct_grad(p):
    """Gradient buffer to accumulate INTO (the parameter's pre-allocated `.grad`, e.g. a view of
    cwn_b200.dist.FlatGradBucket) or None. Writing there from the kernel replaces one AccumulateGrad elementwise
    launch per parameter per step (~150 of them); autograd then receives no gradient for that input — so
    `torch.autograd.grad()` over such parameters returns None for them (and mutates `.grad`): set
    `fused.DIRECT_PARAM_GRADS = False` (or keep `.grad` None) when a functional gradient is wanted."""
    g = p.grad
    if DIRECT_PARAM_GRADS and p.requires_grad and g is not None and g.dtype == torch.float32 and g.is_contiguous() \
            and g.is_cuda and g.shape == p.shape and not p._backward_hooks:
        return g  # (frozen parameters and parameters with gradient hooks go through autograd as usual)
    return None


DIRECT_PARAM_GRADS = True


def _algorithmic_bytes(fn_name, d):
    """Compulsory fp32 traffic of one problem of a grouped launch (inputs read once, outputs written once)."""
    n, h = d.n_rows, d.h
    if fn_name == 'cwn_linear_fwd_grouped':
        k = d.k0 + d.k1
        return 4 * (n * k + n * h + h * k + h)
    if fn_name == 'cwn_bn_act_grouped':
        return 4 * (2 * n * h + 3 * h)
    if fn_name == 'cwn_unit_bwd_reduce_grouped':
        return 4 * (2 * n * h) if d.has_bn else 0
    if fn_name == 'cwn_unit_bwd_grouped':
        k = d.k0 + d.k1
        return 4 * (2 * n * h + n * k + (n * k if (d.g_in0 or d.g_in1) else 0) + h * k + d.n_ctas * (h * k + h))
    if fn_name == 'cwn_wgrad_finalize_grouped':
        k = d.k0 + d.k1
        return 4 * (d.n_ctas + 1) * (h * k + h)
    return 0  # the two finalize kernels move a few KB


def _algorithmic_flops(fn_name, d):
    """fp32-equivalent flops of the contractions of one problem (2 n K h per product)."""
    if fn_name == 'cwn_linear_fwd_grouped':
        return 2 * d.n_rows * (d.k0 + d.k1) * d.h
    if fn_name == 'cwn_unit_bwd_grouped':
        return 2 * d.n_rows * (d.k0 + d.k1) * d.h * (2 if (d.g_in0 or d.g_in1) else 1)
    return 0


def _finalize_wgrads(descs, keep, dev):
    """Ordered sums of the per-CTA weight-gradient partials -> dW, db. When every problem accumulates straight into a
    pre-allocated `.grad` (the FlatGradBucket mode) nothing later in the backward pass reads the result, so the launch
    goes to the deferred tail stream (joined when backward ends); gradients handed back to autograd stay in line."""
    from cwn_b200.streams import run_deferred
    if all(d.accumulate_w for d in descs):
        run_deferred(lambda: _launch('cwn_wgrad_finalize_grouped', _lib.UnitBwdDesc, descs), (descs, list(keep)), dev)
    else:
        _launch('cwn_wgrad_finalize_grouped', _lib.UnitBwdDesc, descs)


def _launch(fn_name, desc_type, descs):
    lib = _lib.load()
    fn = getattr(lib, fn_name)
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(0, len(descs), _lib.MAX_GROUP):
        chunk = descs[i:i + _lib.MAX_GROUP]
        arr = (desc_type * len(chunk))(*chunk)
        nbytes = sum(_algorithmic_bytes(fn_name, d) for d in chunk) if ops._profile is not None else 0
        flops = sum(_algorithmic_flops(fn_name, d) for d in chunk) if ops._profile is not None else 0
        ops._call(fn_name[4:], nbytes, fn, arr, len(chunk), stream, flops=flops)


def _n_ctas(tile_counts):
    """CTAs per problem of one grouped backward launch (each CTA strides over its problem's row tiles). The
    tensor-core kernel holds one CTA per SM (130-190 KB of shared memory), so more than 148 CTAs would mean a second
    wave that pays the descriptor + weight staging again: the tiles are spread over at most 148 CTAs instead."""
    total = sum(tile_counts)
    if not _TC5 or total <= 148:
        return [min(t, 2 * 148) for t in tile_counts]
    return [max(1, t * 148 // total) if t else 0 for t in tile_counts]


class _UnitState(object):
    """Everything one unit keeps between forward and backward."""
    __slots__ = ('unit', 'x0', 'x1', 'in0', 'in1', 'in_act', 'z', 'mean', 'scale', 'rstd', 'n', 'h', 'live')


def _in_vectors(prev):
    """(mean, scale, beta) of the BatchNorm feeding a unit's input, or None if that unit had no BatchNorm."""
    if prev is None or prev.unit.bn is None:
        return None
    return prev.mean, prev.scale, prev.unit.bn.bias


class FusedSparseCINDense(Function):
    """outs = f(us[0], bs[0], us[1], bs[1], ..., *parameters). `ctx_forms` carries the module structure."""

    @staticmethod
    def forward(ctx, forms, training, n_dims, nb, *tensors):
        us, bs = list(tensors[0:nb * n_dims:nb]), list(tensors[1:nb * n_dims:nb])
        ds = list(tensors[2:nb * n_dims:nb]) if nb == 3 else None
        dev = us[0].device
        with torch.cuda.device(dev):
            states = []  # per dim: dict name -> _UnitState

            def new_state(unit, x0, x1, prev0, prev1, d, z=None, vecs=None):
                """`z` / `vecs`: pre-placed output matrix [n, h] (a column range of a wider matrix) and statistics
                vectors [3, h] — how two branches of CIN++ write side by side, so that the combine unit sees them as ONE
                input block (the kernels take two blocks; CIN++ has three)."""
                st = _UnitState()
                st.unit, st.x0, st.x1 = unit, x0, x1
                st.live = _live(d)
                st.in0, st.in1 = _in_vectors(prev0), _in_vectors(prev1)
                st.in_act = prev0.unit.act if prev0 is not None else 'id'
                st.n, st.h = x0.size(0), unit.lin.out_features
                st.z = z if z is not None else torch.empty(st.n, st.h, dtype=torch.float32, device=dev)
                st.mean = st.scale = st.rstd = None
                if vecs is not None and unit.bn is not None:
                    st.mean, st.scale, st.rstd = vecs[0], vecs[1], vecs[2]
                return st

            def run_units(sts):
                lin, alive = [], []  # `alive`: descriptors hold raw pointers only — keep the scratch tensors referenced
                counters = _counters(dev, len(sts))
                tr = _tile_rows([st.n for st in sts])
                for i, st in enumerate(sts):
                    unit = st.unit
                    n_tiles = (st.n + tr - 1) // tr
                    stats = None
                    bn_fields = (None, 0.0, 0.0, 0, None, None, None, None, None, None, None)
                    if unit.bn is not None:
                        m = unit.bn
                        if st.mean is None:
                            vecs = torch.empty(3, st.h, dtype=torch.float32, device=dev)
                            st.mean, st.scale, st.rstd = vecs[0], vecs[1], vecs[2]
                        if training:
                            stats = torch.empty(max(n_tiles, 1) * 2 * st.h, dtype=torch.float32, device=dev)
                            alive.append(stats)
                        bn_fields = (_p(m.weight), float(m.eps), float(m.momentum), int(training), _p(m.running_mean),
                                     _p(m.running_var), _p(m.num_batches_tracked), _p(st.mean), _p(st.scale),
                                     _p(st.rstd), counters[i:i + 1].data_ptr())
                    i0 = st.in0 or (None, None, None)
                    i1 = st.in1 or (None, None, None)
                    lin.append(_lib.LinearDesc(
                        _p(st.x0), st.x0.stride(0), st.x0.size(1), _p(st.x1), st.x1.stride(0) if st.x1 is not None else 0,
                        st.x1.size(1) if st.x1 is not None else 0, _p(i0[0]), _p(i0[1]), _p(i0[2]), _p(i1[0]), _p(i1[1]),
                        _p(i1[2]), ops.ACT_CODES[st.in_act], _p(unit.lin.weight), unit.lin.weight.stride(0),
                        _p(unit.lin.bias), _p(st.z), st.z.stride(0), _p(stats), st.n, st.h, *bn_fields, tr, _p(st.live)))
                _launch('cwn_linear_fwd_grouped', _lib.LinearDesc, lin)

            l1, l2, l3 = [], [], []
            for d in range(n_dims):
                up, bnd, comb, down = forms[d]
                s = {}
                s['u1'] = new_state(up[0], us[d].contiguous(), None, None, None, d)
                s['b1'] = new_state(bnd[0], bs[d].contiguous(), None, None, None, d)
                states.append(s)
                l1 += [s['u1'], s['b1']]
                if down is not None:
                    s['d1'] = new_state(down[0], ds[d].contiguous(), None, None, None, d)
                    l1.append(s['d1'])
            run_units(l1)
            for d in range(n_dims):
                up, bnd, comb, down = forms[d]
                s = states[d]
                if down is None:
                    s['u2'] = new_state(up[1], s['u1'].z, None, s['u1'], None, d)
                else:  # up and down second units write the two halves of ONE [n, 2h] matrix (and of one [3, 2h] vector block)
                    h = up[1].lin.out_features
                    zz = torch.empty(s['u1'].n, 2 * h, dtype=torch.float32, device=dev)
                    vv = torch.empty(3, 2 * h, dtype=torch.float32, device=dev)
                    s['zz'], s['vv'] = zz, vv
                    s['u2'] = new_state(up[1], s['u1'].z, None, s['u1'], None, d, z=zz[:, :h], vecs=vv[:, :h])
                    s['d2'] = new_state(down[1], s['d1'].z, None, s['d1'], None, d, z=zz[:, h:], vecs=vv[:, h:])
                    l2.append(s['d2'])
                s['b2'] = new_state(bnd[1], s['b1'].z, None, s['b1'], None, d)
                l2 += [s['u2'], s['b2']]
            run_units(l2)
            for d in range(n_dims):
                s = states[d]
                if forms[d][3] is None:
                    s['c'] = new_state(forms[d][2], s['u2'].z, s['b2'].z, s['u2'], s['b2'], d)
                else:  # block 0 = [up | down] (2h columns), block 1 = boundaries
                    st = new_state(forms[d][2], s['zz'], s['b2'].z, None, s['b2'], d)
                    u2, d2 = s['u2'], s['d2']
                    if u2.unit.bn is not None:
                        st.in0 = (s['vv'][0], s['vv'][1], torch.cat([u2.unit.bn.bias.detach(), d2.unit.bn.bias.detach()]))
                    st.in_act = u2.unit.act
                    s['c'] = st
                l3.append(s['c'])
            run_units(l3)
            outs, descs = [], []
            for d in range(n_dims):
                st = states[d]['c']
                out = torch.empty(st.n, st.h, dtype=torch.float32, device=dev)
                beta = st.unit.bn.bias if st.unit.bn is not None else None
                descs.append(_lib.BNActDesc(_p(st.z), st.h, _p(st.mean), _p(st.scale), _p(beta),
                                            ops.ACT_CODES[st.unit.act], _p(out), st.h, st.n, st.h))
                outs.append(out)
            _launch('cwn_bn_act_grouped', _lib.BNActDesc, descs)
        ctx.states, ctx.n_dims, ctx.forms, ctx.nb = states, n_dims, forms, nb
        ctx.n_inputs = len(tensors)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        states, n_dims = ctx.states, ctx.n_dims
        dev = states[0]['c'].z.device
        grads = {}  # id(parameter) -> gradient tensor
        keep = []

        def new(*shape):
            t = torch.empty(*shape, dtype=torch.float32, device=dev)
            keep.append(t)
            return t

        bn_bufs = {}  # id(state) -> BatchNorm-backward buffers of that unit (shared by its own descriptor and, when the
        #               reduction is fused, by the `next_*` fields of the unit downstream)

        def bn_backward_buffers(st, tr):
            b = bn_bufs.get(id(st))
            if b is None:
                unit = st.unit
                n_tiles = (st.n + tr - 1) // tr
                red, c1, c2 = new(max(n_tiles, 1) * 2 * st.h), new(st.h), new(st.h)
                gg, gbeta = _direct_grad(unit.bn.weight), _direct_grad(unit.bn.bias)
                acc_a = int(gg is not None and gbeta is not None)
                if not acc_a:
                    gg, gbeta = new(st.h), new(st.h)
                    grads[id(unit.bn.weight)], grads[id(unit.bn.bias)] = gg, gbeta
                b = bn_bufs[id(st)] = dict(red=red, c1=c1, c2=c2, gg=gg, gbeta=gbeta, acc_a=acc_a, tr=tr, reduced=False)
            return b

        def unit_descs(items):
            """items: list of (state, g_out, want_g_in0, want_g_in1[, upstream0, upstream1]) -> (descs, list of (g_in0,
            g_in1)). `upstream_i`: the unit whose output is input block i (its BatchNorm-backward reduction is then
            taken from this unit's g_in tiles, cwn_unit_bwd_desc::next_*), or None."""
            descs, gins, ups = [], [], []
            counters = _counters(dev, 2 * len(items))
            tr = _tile_rows([it[0].n for it in items])
            ctas = _n_ctas([(it[0].n + tr - 1) // tr for it in items])
            for i, it in enumerate(items):
                st, g, want0, want1 = it[:4]
                up0, up1 = (it[4], it[5]) if len(it) > 4 else (None, None)
                unit = st.unit
                k0 = st.x0.size(1)
                k1 = st.x1.size(1) if st.x1 is not None else 0
                n_ctas = ctas[i]
                has_bn = unit.bn is not None
                g = g if (g.stride(1) == 1 and g.stride(0) % 4 == 0) else g.contiguous()  # (column ranges stay views)
                gi0 = new(st.n, k0) if want0 else None
                gi1 = new(st.n, k1) if (want1 and k1) else None
                # parameter gradients: straight into `.grad` when it is pre-allocated (accumulate), else fresh tensors
                gw, gb = _direct_grad(unit.lin.weight), _direct_grad(unit.lin.bias)
                acc_w = int(gw is not None and gb is not None)
                if not acc_w:
                    gw, gb = new(st.h, k0 + k1), new(st.h)
                    grads[id(unit.lin.weight)], grads[id(unit.lin.bias)] = gw, gb
                red = c1 = c2 = gg = gbeta = None
                acc_a = 0
                if has_bn:
                    b = bn_backward_buffers(st, tr)
                    red, c1, c2, gg, gbeta, acc_a = b['red'], b['c1'], b['c2'], b['gg'], b['gbeta'], b['acc_a']
                def upstream(up, gi):
                    """(buffers, descriptor fields) of the fused reduction of upstream unit `up` through input gradient `gi`"""
                    if not (_FUSE_REDUCE and up is not None and up.unit.bn is not None and gi is not None and tr == 64
                            and up.rstd is not None):
                        return None, (None,) * 6
                    b_ = bn_backward_buffers(up, tr)
                    return b_, (_p(up.rstd), _p(b_['red']), _p(b_['c1']), _p(b_['c2']), _p(b_['gg']), _p(b_['gbeta']))
                nb0, f0 = upstream(up0, gi0)
                nb1, f1 = upstream(up1, gi1)
                any_next = nb0 is not None or nb1 is not None
                i0 = st.in0 or (None, None, None)
                i1 = st.in1 or (None, None, None)
                descs.append(_lib.UnitBwdDesc(
                    _p(st.x0), st.x0.stride(0), k0, _p(st.x1), st.x1.stride(0) if st.x1 is not None else 0, k1,
                    _p(i0[0]), _p(i0[1]), _p(i0[2]), _p(i1[0]), _p(i1[1]), _p(i1[2]), ops.ACT_CODES[st.in_act],
                    _p(unit.lin.weight), unit.lin.weight.stride(0), _p(st.z), st.z.stride(0), int(has_bn),
                    ops.ACT_CODES[unit.act], _p(st.mean), _p(st.scale), _p(st.rstd),
                    _p(unit.bn.bias) if has_bn else None, _p(g), g.stride(0), _p(red), _p(c1), _p(c2), _p(gg),
                    _p(gbeta), acc_a, _p(gi0), k0, _p(gi1), k1, _p(new(max(n_ctas, 1) * st.h * (k0 + k1))),
                    _p(new(max(n_ctas, 1) * st.h)), n_ctas, _p(gw), gw.stride(0), _p(gb), acc_w, st.n, st.h,
                    counters[i:i + 1].data_ptr() if has_bn else None, tr, 0, _p(st.live),
                    *f0, *f1, nb0['acc_a'] if nb0 else 0, nb1['acc_a'] if nb1 else 0,
                    counters[len(items) + i:len(items) + i + 1].data_ptr() if any_next else None))
                keep.append(g)
                gins.append((gi0, gi1))
                ups.append((nb0, nb1))
            return descs, gins, ups

        _NEXT = ('next_rstd0', 'next_red0', 'next_c1_0', 'next_c2_0', 'next_g_gamma0', 'next_g_beta0', 'next_rstd1',
                 'next_red1', 'next_c1_1', 'next_c2_1', 'next_g_gamma1', 'next_g_beta1', 'next_counter')

        def run(items):
            descs, gins, ups = unit_descs(items)
            if any(nb_ is not None for pair in ups for nb_ in pair):
                arr = (_lib.UnitBwdDesc * len(descs))(*descs)
                if len(descs) <= _lib.MAX_GROUP and _lib.load().cwn_unit_bwd_fuses_reduce(arr, len(descs)):
                    for pair in ups:
                        for nb_ in pair:
                            if nb_ is not None:
                                nb_['reduced'] = True  # this launch produces c1 / c2 / g_gamma / g_beta of the unit upstream
                else:  # (FFMA fallback shapes, or more than one launch group: the upstream units reduce for themselves)
                    for d_ in descs:
                        for name in _NEXT:
                            setattr(d_, name, None)
            # column reductions of the units whose g_out did not come out of a fusing launch; the last CTA finalises them
            todo = [d_ for d_, it in zip(descs, items) if d_.has_bn and not bn_bufs[id(it[0])]['reduced']]
            if todo:
                _launch('cwn_unit_bwd_reduce_grouped', _lib.UnitBwdDesc, todo)
            _launch('cwn_unit_bwd_grouped', _lib.UnitBwdDesc, descs)
            return descs, gins

        with torch.cuda.device(dev):
            all_descs = []
            g_outs = [g if g is not None else torch.zeros_like(states[d]['c'].z) for d, g in enumerate(g_outs)]
            descs, gins = run([(states[d]['c'], g_outs[d], True, True) + ((states[d]['u2'], states[d]['b2']) if ctx.nb == 2 else ())
                                for d in range(n_dims)])
            all_descs += descs
            nb = ctx.nb
            items, slot2 = [], []
            for d in range(n_dims):
                s = states[d]
                if nb == 2:
                    items += [(s['u2'], gins[d][0], True, False, s['u1'], None), (s['b2'], gins[d][1], True, False, s['b1'], None)]
                    slot2 += [(d, 'u'), (d, 'b')]
                else:
                    h = s['u2'].h
                    items += [(s['u2'], gins[d][0][:, :h], True, False), (s['d2'], gins[d][0][:, h:], True, False),
                              (s['b2'], gins[d][1], True, False)]
                    slot2 += [(d, 'u'), (d, 'd'), (d, 'b')]
            descs, gins2 = run(items)
            all_descs += descs
            g2 = {key: gi[0] for key, gi in zip(slot2, gins2)}
            need = {(d, 'ubd'[j]): ctx.needs_input_grad[4 + nb * d + j] for d in range(n_dims) for j in range(nb)}
            items, slot1 = [], []
            for d in range(n_dims):
                for key in (['u', 'b'] if nb == 2 else ['u', 'b', 'd']):
                    items.append((states[d][key + '1'], g2[(d, key)], need[(d, key)], False))
                    slot1.append((d, key))
            descs, gins1 = run(items)
            all_descs += descs
            g1 = {key: gi[0] for key, gi in zip(slot1, gins1)}
            _finalize_wgrads(all_descs, keep, dev)

        out = [None, None, None, None]
        for d in range(n_dims):
            out += [g1[(d, key)] for key in (['u', 'b'] if ctx.nb == 2 else ['u', 'b', 'd'])]
        out += [grads.get(id(p)) for p in _parameters(ctx.forms)]
        ctx.states = None
        return tuple(out)


def _parameters(forms):
    params = []
    for up, bnd, comb, down in forms:
        for unit in up + bnd + [comb] + (down or []):
            params += [unit.lin.weight, unit.lin.bias]
            if unit.bn is not None:
                params += [unit.bn.weight, unit.bn.bias]
    return params


def sparse_cin_dense(forms, us, bs, training, ds=None):
    """Run the update/combine nets of every dimension of a layer. `forms[d] = recognise(level_d)`; `ds`: the inputs of
    the third (down) branch of CIN++."""
    n_dims = len(us)
    params = _parameters(forms)
    flat = []
    for d, (u, b) in enumerate(zip(us, bs)):
        flat += [u, b] + ([ds[d]] if ds is not None else [])
    return list(FusedSparseCINDense.apply(forms, training, n_dims, 3 if ds is not None else 2, *flat, *params))


# ------------------------------------------------------------------------------------------------ plain linears
class GroupedLinear(Function):
    """ys[i] = xs[i] @ W_i[:, off_i : off_i + k_i].T (+ b_i) for up to 8 independent problems in one launch forward
    and two backward (input gradients + per-CTA weight-gradient partials, then their ordered sum). A problem may use a
    COLUMN RANGE of a larger weight matrix — that is how the coboundary message Linear(2F -> F) of
    mp/layers.py:290-293 is applied as two per-cell products (x W[:, :F]^T and y W[:, F:]^T + b) without slicing
    tensors, and how both halves of its gradient land in one [F, 2F] buffer.

    spec: tuple of (weight slot, column offset, bias slot or -1) per problem; tensors = xs, unique weights, biases."""

    @staticmethod
    def forward(ctx, spec, n_weights, *tensors):
        n = len(spec)
        xs = [x if x.stride(1) == 1 else x.contiguous() for x in tensors[:n]]
        weights = tensors[n:n + n_weights]
        biases = tensors[n + n_weights:]
        dev = xs[0].device
        ys, descs = [], []
        with torch.cuda.device(dev):
            tr = _tile_rows([x.size(0) for x in xs])
            for x, (wi, off, bi) in zip(xs, spec):
                w = weights[wi]
                b = biases[bi] if bi >= 0 else None
                y = torch.empty(x.size(0), w.size(0), dtype=torch.float32, device=dev)
                descs.append(_lib.LinearDesc(_p(x), x.stride(0), x.size(1), None, 0, 0, None, None, None, None, None,
                                             None, 0, w.data_ptr() + 4 * off, w.stride(0), _p(b), _p(y), y.size(1), None,
                                             x.size(0), w.size(0), None, 0.0, 0.0, 0, None, None, None, None, None, None,
                                             None, tr))
                ys.append(y)
            _launch('cwn_linear_fwd_grouped', _lib.LinearDesc, descs)
        ctx.spec, ctx.n_weights, ctx.n_biases = spec, n_weights, len(biases)
        ctx.params = list(weights) + list(biases)  # for direct gradient accumulation (identity of the Parameters)
        ctx.save_for_backward(*xs, *weights)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gs):
        spec, n = ctx.spec, len(ctx.spec)
        xs, weights = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        w_params, b_params = ctx.params[:ctx.n_weights], ctx.params[ctx.n_weights:]
        dev = xs[0].device
        with torch.cuda.device(dev):
            # gradient buffers: the parameter's own `.grad` (accumulate) when pre-allocated, else fresh tensors
            gw_bufs, gw_out, acc = [], [], []
            for slot, (w, prm) in enumerate(zip(weights, w_params)):
                direct = _direct_grad(prm)
                covered = sum(xs[i].size(1) for i, (wi, _, _) in enumerate(spec) if wi == slot)
                if direct is not None:
                    gw_bufs.append(direct), gw_out.append(None), acc.append(1)
                else:
                    buf = torch.empty_like(w, memory_format=torch.contiguous_format) if covered == w.size(1) \
                        else torch.zeros_like(w, memory_format=torch.contiguous_format)
                    gw_bufs.append(buf), gw_out.append(buf), acc.append(0)
            gb_bufs, gb_out, acc_b = [], [], []
            for prm in b_params:
                direct = _direct_grad(prm)
                if direct is not None:
                    gb_bufs.append(direct), gb_out.append(None), acc_b.append(1)
                else:
                    buf = torch.empty_like(prm)
                    gb_bufs.append(buf), gb_out.append(buf), acc_b.append(0)
            gxs, descs, keep = [], [], []
            tr = _tile_rows([x.size(0) for x in xs])
            ctas = _n_ctas([(x.size(0) + tr - 1) // tr for x in xs])
            for i, (x, g, (wi, off, bi)) in enumerate(zip(xs, gs, spec)):
                w = weights[wi]
                nr, k, h = x.size(0), x.size(1), w.size(0)
                g = (g if g is not None else torch.zeros(nr, h, device=dev)).contiguous()
                n_tiles = (nr + tr - 1) // tr
                n_ctas = ctas[i]
                gx = torch.empty(nr, k, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2 + i] else None
                wp = torch.empty(max(n_ctas, 1) * h * k, dtype=torch.float32, device=dev)
                bp = torch.empty(max(n_ctas, 1) * h, dtype=torch.float32, device=dev)
                gb = gb_bufs[bi] if bi >= 0 else None
                # a weight and its bias accumulate together (one flag per problem): both or neither are direct
                accumulate = acc[wi] if bi < 0 else int(acc[wi] and acc_b[bi])
                if bi >= 0 and acc[wi] != acc_b[bi]:
                    raise RuntimeError('cwn_b200: weight and bias of a Linear must both have (or both lack) a .grad')
                keep += [g, wp, bp]
                # the "z" operand is only read through act'(z) with act = id, so any valid [n, h] matrix will do
                descs.append(_lib.UnitBwdDesc(
                    _p(x), x.stride(0), k, None, 0, 0, None, None, None, None, None, None, 0,
                    w.data_ptr() + 4 * off, w.stride(0), _p(g), g.stride(0), 0, 0, None, None, None, None, _p(g),
                    g.stride(0), None, None, None, None, None, 0, _p(gx), k, None, 0, _p(wp), _p(bp), n_ctas,
                    gw_bufs[wi].data_ptr() + 4 * off, gw_bufs[wi].stride(0), _p(gb), accumulate, nr, h, None, tr))
                gxs.append(gx)
            _launch('cwn_unit_bwd_grouped', _lib.UnitBwdDesc, descs)
            _finalize_wgrads(descs, keep + gw_bufs + gb_bufs, dev)
        return (None, None, *gxs, *gw_out, *gb_out)


def grouped_linear(problems):
    """problems: list of (x, weight, column offset, bias or None); y = x @ weight[:, off:off + x.size(1)].T + bias.
    One launch for all of them; fp32 CUDA matrices."""
    weights, biases, spec = [], [], []
    for x, w, off, b in problems:
        wi = next((i for i, t in enumerate(weights) if t is w), None)
        if wi is None:
            weights.append(w)
            wi = len(weights) - 1
        bi = -1
        if b is not None:
            biases.append(b)
            bi = len(biases) - 1
        spec.append((wi, int(off), bi))
    return list(GroupedLinear.apply(tuple(spec), len(weights), *[p[0] for p in problems], *weights, *biases))


# ----------------------------------------------------------------------------------------------- readout head
class _ReadoutHead(Function):
    """pool_complex + per-dimension lin1 + act + sum/mean over dimensions + lin2 (reference mp/nn.py:50-60,
    mp/models.py:230-254, mp/molec_models.py:137-161) as one launch forward and two backward (`csrc/head.cu`)."""

    @staticmethod
    def forward(ctx, cfg, w2, b2, *tensors):
        n = cfg['n_dims']
        xs = [None if x is None else (x if x.stride(1) == 1 else x.contiguous()) for x in tensors[:n]]
        w1s = [w.contiguous() for w in tensors[n:2 * n]]
        b1s = tensors[2 * n:3 * n]
        B, K, H2, out_size = cfg['B'], cfg['K'], w1s[0].size(0), w2.size(0)
        dev = w2.device
        w2c = w2.contiguous()
        with torch.cuda.device(dev):
            scratch = torch.empty(n * B * (K + H2) + B * H2, dtype=torch.float32, device=dev)
            pooled = [scratch[d * B * K:(d + 1) * B * K] for d in range(n)]
            zs = [scratch[n * B * K + d * B * H2:n * B * K + (d + 1) * B * H2] for d in range(n)]
            h = scratch[n * B * (K + H2):]
            out = torch.empty(B, out_size, dtype=torch.float32, device=dev)
            descs = (_lib.HeadDim * n)()
            for d in range(n):
                plan = cfg['plans'][d] if xs[d] is not None else None
                descs[d] = _lib.HeadDim(_p(xs[d]), xs[d].stride(0) if xs[d] is not None else 0,
                                        plan.rowptr.data_ptr() if plan is not None else None,
                                        _p(plan.perm) if plan is not None else None, _p(w1s[d]), _p(b1s[d]),
                                        _p(pooled[d]), _p(zs[d]), None, None, 0, None, None, 0)
            ops._call('readout_head_fwd', 4 * (sum(x.numel() for x in xs if x is not None) + n * H2 * K + B * out_size),
                      _lib.load().cwn_readout_head_fwd, descs, n, B, K, H2, out_size, cfg['act'], cfg['pool_mean'],
                      cfg['final_mean'], _p(w2c), _p(b2), _p(h), _p(out), ops._stream())
        ctx.cfg, ctx.n = cfg, n
        ctx.shapes = [None if x is None else tuple(x.shape) for x in xs]
        ctx.save_for_backward(scratch, w2c, *w1s)
        return out

    @staticmethod
    def backward(ctx, g_out):
        cfg, n = ctx.cfg, ctx.n
        scratch, w2 = ctx.saved_tensors[:2]
        w1s = ctx.saved_tensors[2:]
        B, K, H2, out_size = cfg['B'], cfg['K'], w1s[0].size(0), w2.size(0)
        dev = w2.device
        g_out = g_out.contiguous()
        lin1s, lin2 = cfg['lin1s'], cfg['lin2']
        with torch.cuda.device(dev):
            pooled = [scratch[d * B * K:(d + 1) * B * K] for d in range(n)]
            zs = [scratch[n * B * K + d * B * H2:n * B * K + (d + 1) * B * H2] for d in range(n)]
            h = scratch[n * B * (K + H2):]
            g_z = torch.empty(n * B * H2, dtype=torch.float32, device=dev)
            gxs, gw1, gb1 = [], [], []
            descs = (_lib.HeadDim * n)()
            for d in range(n):
                shape = ctx.shapes[d]
                gx = torch.empty(shape, dtype=torch.float32, device=dev) \
                    if shape is not None and ctx.needs_input_grad[3 + d] else None
                plan = cfg['plans'][d] if shape is not None else None
                lin = lin1s[d]
                dw, db = _direct_grad(lin.weight), (_direct_grad(lin.bias) if lin.bias is not None else None)
                direct = dw is not None and (lin.bias is None or db is not None)
                if direct:
                    gw, gb = dw, db
                    gw1.append(None), gb1.append(None)
                else:
                    gw = torch.empty_like(lin.weight, memory_format=torch.contiguous_format)
                    gb = torch.empty_like(lin.bias) if lin.bias is not None else None
                    gw1.append(gw), gb1.append(gb)
                descs[d] = _lib.HeadDim(None, K, plan.rowptr.data_ptr() if plan is not None else None,
                                        _p(plan.perm) if plan is not None else None, _p(w1s[d]), None, _p(pooled[d]),
                                        _p(zs[d]), g_z.data_ptr() + 4 * d * B * H2, _p(gx), K, _p(gw), _p(gb),
                                        1 if direct else 0)
                gxs.append(gx)
            dw2, db2 = _direct_grad(lin2.weight), (_direct_grad(lin2.bias) if lin2.bias is not None else None)
            direct2 = dw2 is not None and (lin2.bias is None or db2 is not None)
            if direct2:
                gw2, gb2, gw2_out, gb2_out = dw2, db2, None, None
            else:
                gw2 = gw2_out = torch.empty_like(w2)
                gb2 = gb2_out = torch.empty_like(lin2.bias) if lin2.bias is not None else None
            algo = 4 * (sum(g.numel() for g in gxs if g is not None) + 2 * n * H2 * K)

            def launch(parts):
                ops._call('readout_head_bwd', algo if parts & 1 else 0, _lib.load().cwn_readout_head_bwd_parts, descs, n, B, K,
                          H2, out_size, cfg['act'], cfg['pool_mean'], cfg['final_mean'], _p(w2), _p(h), _p(g_out), _p(gw2),
                          _p(gb2), 1 if direct2 else 0, parts, ops._stream())
            if direct2 and all(g is None for g in gw1) and all(g is None for g in gb1):
                # every parameter gradient goes straight into `.grad`: the ordered sums over the complexes (~40 us at
                # B = 128) leave the critical path for the tail stream that joins when backward ends
                from cwn_b200.streams import run_deferred
                launch(1)
                run_deferred(lambda: launch(2), (descs, g_z, g_out, h, pooled, zs, gw2, gb2), dev)
            else:
                launch(3)
        return (None, gw2_out, gb2_out, *gxs, *gw1, *gb1)


def readout_head(model, xs, data, act_name):
    """The fused readout of a SparseCIN-family model, or NotImplemented when its configuration is outside the closed
    form (active dropout, partial outputs requested by the caller, exotic readouts, non-CUDA tensors)."""
    from cwn_b200.mp.nn import num_complexes_of
    if act_name is None or model.readout not in ('sum', 'mean') or model.final_readout not in ('sum', 'mean'):
        return NotImplemented
    if model.training and model.dropout_rate > 0:
        return NotImplemented
    dims = list(model.readout_dims)
    n = len(dims)
    if n < 1 or n > 4 or not xs or xs[0] is None:
        return NotImplemented
    K = xs[0].size(-1)
    sel = [xs[d] if d < len(xs) else None for d in dims]  # dimensions absent from the batch pool to zeros
    for x in sel:
        if x is not None and not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.size(1) == K):
            return NotImplemented
    lin1s = [model.lin1s[d] for d in dims]
    lin2 = model.lin2
    H2 = lin1s[0].out_features
    if any(l.in_features != K or l.out_features != H2 for l in lin1s) or lin2.in_features != H2:
        return NotImplemented
    if any((l.bias is None) != (lin1s[0].bias is None) for l in lin1s) or (n * K + H2) * 4 > 160 * 1024:
        return NotImplemented
    B = num_complexes_of(data)
    plans = [ops._row_plan(data.cochains[d].batch, B) if x is not None else None for d, x in zip(dims, sel)]
    cfg = {'n_dims': n, 'B': B, 'K': K, 'plans': plans, 'act': ops.ACT_CODES[act_name],
           'pool_mean': int(model.readout == 'mean'), 'final_mean': int(model.final_readout == 'mean'),
           'lin1s': lin1s, 'lin2': lin2}
    return _ReadoutHead.apply(cfg, lin2.weight, lin2.bias, *sel, *[l.weight for l in lin1s], *[l.bias for l in lin1s])


# ----------------------------------------------------------------------------------------------- layer aggregation
def _scale_rows(x, eps, x2=None, eps2=None, src=None, plan=None):
    """(1+eps) x [+ (1+eps2) x2] [+ SUM over `plan` rows of src[idx]] in one launch (cwn_csr_gather_reduce2_f32)."""
    n, F = x.size(0), x.size(1)
    out = torch.empty(n, F, dtype=torch.float32, device=x.device)
    E = plan.pay0.numel() if plan is not None else 0
    algo = 16 * E + 4 * F * (n * (2 + (x2 is not None)) + (src.size(0) if src is not None else 0))
    ops._call('csr_gather_reduce', algo, _lib.load().cwn_csr_gather_reduce2_f32,
              _p(src) if E else None, src.stride(0) if src is not None else F,
              plan.rowptr.data_ptr() if E else None, _p(plan.pay0) if E else None, n, F,
              _p(x), x.stride(0), _p(eps), _p(x2), x2.stride(0) if x2 is not None else F, _p(eps2), _p(out), F,
              ops._stream())
    return out


class _LayerAggregate(Function):
    """Every aggregation pass of one SparseCINConv layer — all cochain dimensions, upper and boundary branches, the
    split-weight products of the coboundary message nets — as ONE autograd node.

        u_d = SUM_up act(P_d[src] + Q_d[cob]) + (1+eps1_d) x_d       P_d = x_d W1_d^T,  Q_d = x_{d+1} W2_d^T + b_d
        b_d = SUM_bnd x_{d-1}[src]            + (1+eps2_d) x_d       (reference mp/layers.py:184-199, 210-214, 290-299)

    Why one node: x_d feeds five consumers (two residuals, P_d, Q_{d-1}, the boundary pass of d+1). As separate
    autograd nodes their gradients meet in the engine as 2 scaled copies + 4 elementwise additions per dimension per
    layer (~70 launches per step). Here the backward computes
        gx_d  = (1+eps1_d) gU_d + (1+eps2_d) gB_d + transposed boundary pass of gB_{d+1}      (one launch per d)
        gx_d += gP_d W1_d ;  gx_{d+1} += gQ_d W2_d                                            (accumulating GEMM epilogues)
    so nothing is left for the engine to add."""

    @staticmethod
    def forward(ctx, cfg, *tensors):
        from cwn_b200.streams import run_concurrently
        n = cfg['n_dims']
        xs = [t if t.stride(1) == 1 else t.contiguous() for t in tensors[:n]]
        eps = list(tensors[n:3 * n])                       # eps1_0, eps2_0, eps1_1, ...
        wb = list(tensors[3 * n:])                         # (weight, bias) per dimension with a coboundary up pass
        dev = xs[0].device
        ups, bnds = cfg['ups'], cfg['bnds']
        cob_dims = [d for d in range(n) if ups[d] is not None]
        with torch.cuda.device(dev):
            prods = {}
            if cob_dims:
                problems = []
                for i, d in enumerate(cob_dims):
                    w, b = wb[2 * i], wb[2 * i + 1]
                    fx = xs[d].size(1)
                    problems += [(xs[d], w, 0, None), (xs[d + 1], w, fx, b)]
                outs = grouped_linear(problems)            # autograd is off inside Function.forward: forward only
                for i, d in enumerate(cob_dims):
                    prods[d] = (outs[2 * i], outs[2 * i + 1])
            thunks = []
            for d in range(n):
                x, e1, e2 = xs[d], eps[2 * d], eps[2 * d + 1]
                if ups[d] is not None:
                    adj, act = ups[d]
                    thunks.append(lambda x=x, e1=e1, adj=adj, act=act, pq=prods[d]:
                                  ops._CobPass.apply(pq[0], pq[1], x, e1, adj, act))
                else:
                    thunks.append(lambda x=x, e1=e1: _scale_rows(x, e1))
                if bnds[d] is not None:
                    thunks.append(lambda x=x, e2=e2, adj=bnds[d], src=xs[d - 1]:
                                  ops._GatherReduce.apply(src, x, e2, adj, 'add'))
                else:
                    thunks.append(lambda x=x, e2=e2: _scale_rows(x, e2))
            outs = run_concurrently(thunks, dev)
        ctx.cfg, ctx.n, ctx.cob_dims = cfg, n, cob_dims
        ctx.save_for_backward(*xs, *eps, *wb, *[t for d in cob_dims for t in prods[d]])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        from cwn_b200.streams import run_concurrently
        cfg, n, cob_dims = ctx.cfg, ctx.n, ctx.cob_dims
        sv = ctx.saved_tensors
        xs, eps = list(sv[:n]), list(sv[n:3 * n])
        wb = list(sv[3 * n:3 * n + 2 * len(cob_dims)])
        pq = list(sv[3 * n + 2 * len(cob_dims):])
        ups, bnds, lins = cfg['ups'], cfg['bnds'], cfg['lins']
        dev = xs[0].device
        lib = _lib.load()
        gU = [gs[2 * d] for d in range(n)]
        gB = [gs[2 * d + 1] for d in range(n)]
        with torch.cuda.device(dev):
            for d in range(n):
                if gU[d] is None:
                    gU[d] = torch.zeros_like(xs[d])
                if gB[d] is None:
                    gB[d] = torch.zeros_like(xs[d])
                gU[d], gB[d] = gU[d].contiguous(), gB[d].contiguous()
            need_x = [ctx.needs_input_grad[1 + d] for d in range(n)]
            # a dimension's gradient buffer is also needed as the accumulation target of the linear backward
            thunks, slots = [], []
            for d in range(n):
                if not need_x[d]:
                    continue
                up_adj = bnds[d + 1] if d + 1 < n else None    # boundary pass of d+1 reads x_d: transposed plan
                thunks.append(lambda d=d, up_adj=up_adj: _scale_rows(
                    gU[d], eps[2 * d], gB[d], eps[2 * d + 1],
                    gB[d + 1] if up_adj is not None else None, up_adj.by_src if up_adj is not None else None))
                slots.append(('x', d))
            for i, d in enumerate(cob_dims):
                adj, act = ups[d]
                P, Q = pq[2 * i], pq[2 * i + 1]
                F = P.size(1)
                code = ops.ACT_CODES[act]

                def cob_bwd(A, B_, plan, n_rows, g=gU[d], F=F, code=code):
                    out = torch.empty(n_rows, F, dtype=torch.float32, device=dev)
                    algo = 24 * plan.pay0.numel() + 4 * F * (g.size(0) + 2 * A.size(0) + B_.size(0))
                    ops._call('csr_cob_bwd', algo, lib.cwn_csr_cob_bwd_f32, _p(g), g.stride(0), _p(A), A.stride(0),
                              _p(B_), B_.stride(0), plan.rowptr.data_ptr(), _p(plan.pay0), _p(plan.pay1), n_rows, F,
                              code, _p(out), F, ops._stream())
                    return out
                thunks.append(lambda P=P, Q=Q, adj=adj, f=cob_bwd: f(P, Q, adj.by_src, adj.n_src))
                slots.append(('p', d))
                thunks.append(lambda P=P, Q=Q, adj=adj, f=cob_bwd: f(Q, P, adj.by_cob, adj.n_cob))
                slots.append(('q', d))
            res = run_concurrently(thunks, dev)
            gx = [None] * n
            gP, gQ = {}, {}
            for (kind, d), t in zip(slots, res):
                if kind == 'x':
                    gx[d] = t
                elif kind == 'p':
                    gP[d] = t
                else:
                    gQ[d] = t
            # split-weight products backward: two grouped launches (the P problems, then the Q problems) so that no two
            # problems of one launch accumulate into the same gx buffer
            gw_out, gb_out = [], []
            if cob_dims:
                tr = _tile_rows([xs[d].size(0) for d in cob_dims])
                keep, all_descs = [], []
                bufs = []
                for i, d in enumerate(cob_dims):
                    lin = lins[d]
                    dw, db = _direct_grad(lin.weight), _direct_grad(lin.bias)
                    direct = dw is not None and db is not None
                    if direct:
                        bufs.append((dw, db, 1))
                        gw_out.append(None), gb_out.append(None)
                    else:
                        gw, gb = torch.empty_like(wb[2 * i], memory_format=torch.contiguous_format), torch.empty_like(wb[2 * i + 1])
                        bufs.append((gw, gb, 0))
                        gw_out.append(gw), gb_out.append(gb)
                for which in ('p', 'q'):
                    descs = []
                    ctas = _n_ctas([((xs[d] if which == 'p' else xs[d + 1]).size(0) + tr - 1) // tr for d in cob_dims])
                    for i, d in enumerate(cob_dims):
                        w = wb[2 * i]
                        gw, gb, acc = bufs[i]
                        fx = xs[d].size(1)
                        x, g, off, tgt = (xs[d], gP[d], 0, d) if which == 'p' else (xs[d + 1], gQ[d], fx, d + 1)
                        nr, k, h = x.size(0), x.size(1), w.size(0)
                        n_ctas = ctas[i]
                        wp = torch.empty(max(n_ctas, 1) * h * k, dtype=torch.float32, device=dev)
                        bp = torch.empty(max(n_ctas, 1) * h, dtype=torch.float32, device=dev)
                        keep += [wp, bp]
                        descs.append(_lib.UnitBwdDesc(
                            _p(x), x.stride(0), k, None, 0, 0, None, None, None, None, None, None, 0,
                            w.data_ptr() + 4 * off, w.stride(0), _p(g), g.stride(0), 0, 0, None, None, None, None, _p(g),
                            g.stride(0), None, None, None, None, None, 0, _p(gx[tgt]), k, None, 0, _p(wp), _p(bp),
                            n_ctas, gw.data_ptr() + 4 * off, gw.stride(0), _p(gb) if which == 'q' else None, acc, nr, h,
                            None, tr, 1))
                    _launch('cwn_unit_bwd_grouped', _lib.UnitBwdDesc, descs)
                    all_descs += descs
                _finalize_wgrads(all_descs, keep + [b for pair in bufs for b in pair[:2]], dev)
            # trainable epsilons (train_eps=True): scalar reductions through torch
            g_eps = []
            for d in range(n):
                for j, g in ((0, gU[d]), (1, gB[d])):
                    e = eps[2 * d + j]
                    g_eps.append((g * xs[d]).sum().reshape(e.shape) if ctx.needs_input_grad[1 + n + 2 * d + j] else None)
        flat_wb = [t for pair in zip(gw_out, gb_out) for t in pair]
        return (None, *gx, *g_eps, *flat_wb)


def layer_aggregate(levels, cochain_params):
    """(us, bs) of a SparseCINConv layer through `_LayerAggregate`, or NotImplemented when some level is outside its
    closed form: upper pass = coboundary message net `act(Linear([x_j ; y_cob]))` or no upper adjacency at all,
    boundary pass = identity messages from `x_{d-1}`, fp32 CUDA features of one width, default hooks."""
    from cwn_b200.mp.layers import identity
    from cwn_b200.mp.params import LazyRows
    n = len(cochain_params)
    xs = [p.x for p in cochain_params]
    if not all(isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 for x in xs):
        return NotImplemented
    F = xs[0].size(1)
    ups, bnds, lins, eps, wb = [], [], {}, [], []
    for d, (level, p) in enumerate(zip(levels, cochain_params)):
        x = xs[d]
        if x.size(1) != F or x.size(0) < 1 or level.msg_boundaries_nn is not identity or not level._hooks_untouched():
            return NotImplemented
        up_index, up_attr = p.up_index, p.kwargs['up_attr']
        if up_index is None:
            if level.up_msg_size != F:
                return NotImplemented
            ups.append(None)
        else:
            form = level._up_message_form()
            if form is None or form[0] != 'cob' or d + 1 >= n:
                return NotImplemented
            lin, act = form[1], form[2]
            if not (isinstance(up_attr, LazyRows) and up_attr.source is xs[d + 1] and lin.in_features == 2 * F
                    and lin.out_features == F and lin.bias is not None):
                return NotImplemented
            adj = ops.Adjacency.of(up_index, x.size(0), x.size(0), up_attr.index, xs[d + 1].size(0))
            ups.append((adj, act))
            lins[d] = lin
            wb += [lin.weight, lin.bias]
        b_index, b_attr = p.boundary_index, p.kwargs['boundary_attr']
        if b_attr is None:
            if level.boundary_msg_size != F:
                return NotImplemented
            bnds.append(None)
        else:
            if b_index is None or d == 0 or b_attr is not xs[d - 1]:
                return NotImplemented
            bnds.append(ops.Adjacency.of(b_index, xs[d - 1].size(0), x.size(0)))
        eps += [level.eps1, level._boundary_eps()]
    cfg = {'n_dims': n, 'ups': ups, 'bnds': bnds, 'lins': lins}
    outs = _LayerAggregate.apply(cfg, *xs, *eps, *wb)
    return list(outs[0::2]), list(outs[1::2])
