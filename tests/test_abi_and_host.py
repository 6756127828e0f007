"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/cwn_b200.h declares; the
operator API refuses non-CUDA tensors (no CPU fallback); host-side hook bookkeeping mirrors the reference."""
import ctypes
import os
import re

import pytest
import torch

from cwn_b200 import _lib, ops
from cwn_b200.mp.cell_mp import CochainMessagePassing
from cwn_b200.mp.layers import CINConv, DummyCochainMessagePassing, SparseCINConv, SparseCINCochainConv
from cwn_b200.mp.models import CIN0, CINpp, SparseCIN
from cwn_b200.mp.molec_models import EmbedCINpp, EmbedSparseCIN, OGBEmbedSparseCIN
from helpers import fixture, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'cwn_b200.h')).read()
    declared = set(re.findall(r'\b(cwn_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(_lib.load()._name)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/cwn_b200.h but not exported'
    assert declared == set(_lib.exported_symbols())
    assert b'sm_100a' in _lib.load().cwn_version()


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.cwn_csr_gather_reduce_f32(None, 4, None, None, -1, 4, None, 4, None, None, 4, 0, None) == -2
    assert lib.cwn_csr_gather_reduce_f32(None, 4, None, None, 3, 4, None, 4, None, None, 4, 9, None) == -3
    assert lib.cwn_csr_gather_reduce_f32(None, 4, None, None, 3, 4, None, 4, None, None, 4, 0, None) == -1
    assert b'rowptr' in lib.cwn_last_error_string()
    assert lib.cwn_csr_cob_fwd_f32(None, 4, None, 4, None, None, None, 3, 4, 77, None, 4, None, None, 4, None) == -3
    assert lib.cwn_csr_plan_workspace_bytes(1000, 10) >= 3 * 4000
    with pytest.raises(ValueError):
        _lib.check(-2, 'x')
    with pytest.raises(RuntimeError):
        _lib.check(700, 'x')


def test_no_cpu_fallback():
    house = fixture('house')
    p = house.get_cochain_params(dim=1)
    cmp = CochainMessagePassing(up_msg_size=1, down_msg_size=1)
    with pytest.raises(RuntimeError, match='CUDA-only'):
        cmp.propagate(p.up_index, p.down_index, p.boundary_index, x=p.x, up_attr=p.kwargs['up_attr'],
                      down_attr=p.kwargs['down_attr'], boundary_attr=p.kwargs['boundary_attr'])
    with pytest.raises(RuntimeError, match='CUDA-only'):
        ops.gather_scatter(p.x, p.up_index, 6)
    with pytest.raises(RuntimeError):
        ops.scatter_rows(torch.ones(3, 2), torch.tensor([0, 1, 0]), 2)
    with pytest.raises(RuntimeError):
        ops.build_plan(torch.tensor([0, 1]), 2)
    model = SparseCIN(1, 3, 2, 5)
    from cwn_b200.data.complex import ComplexBatch
    with pytest.raises(RuntimeError, match='CUDA-only'):
        model.eval()(ComplexBatch.from_complex_list([fixture('house'), fixture('kite')]))


def test_propagate_input_checks_match_reference():  # mp/cell_mp.py:153-193
    cmp = CochainMessagePassing(1, 1)
    x = torch.ones(3, 1)
    with pytest.raises(AssertionError):
        cmp.propagate(torch.zeros(2, 2, dtype=torch.int32), None, None, x=x, up_attr=None)
    with pytest.raises(AssertionError):
        cmp.propagate(torch.zeros(3, 2, dtype=torch.long), None, None, x=x, up_attr=None)
    with pytest.raises(ValueError):
        cmp.propagate('nope', None, None, x=x, up_attr=None)
    with pytest.raises(ValueError, match='expected size'):
        cmp.propagate(torch.zeros(2, 2, dtype=torch.long), None, None, up_size=(7, 7), x=x, up_attr=None)
    with pytest.raises(TypeError, match='up_attr'):  # hook argument never passed (Inspector.distribute)
        cmp.propagate(torch.zeros(2, 2, dtype=torch.long), None, None, x=x)
    # no adjacency at all: update() fabricates zeros of the configured message sizes (mp/cell_mp.py:511-524)
    cmp = CochainMessagePassing(up_msg_size=4, down_msg_size=2, boundary_msg_size=3)
    up, down, bnd = cmp.propagate(None, None, None, x=x)
    assert (up.shape, down.shape, bnd.shape) == ((3, 4), (3, 2), (3, 3)) and float(up.abs().sum()) == 0


def test_hook_bookkeeping():
    base = CochainMessagePassing(1, 1)
    assert base._default_hooks == {'up': True, 'down': True, 'boundary': True}
    assert base.__user_args__ == {'up_x_j', 'up_attr', 'down_x_j', 'down_attr', 'boundary_x_j'}
    assert base.__update_user_args__ == {'x'}
    assert not (base.fuse_up or base.fuse_down or base.fuse_boundary)
    dummy = DummyCochainMessagePassing(1, 1)
    assert dummy._default_hooks == {'up': False, 'down': False, 'boundary': True}
    assert base.boundary_msg_size == 1 and CochainMessagePassing(2, 5).boundary_msg_size == 5
    with pytest.raises(AssertionError):
        CochainMessagePassing(1, 1, aggr_up='median')


def test_module_structure_loads_reference_state_dicts():
    """Parameter names/shapes equal the reference's (golden state_dicts were saved by reference models)."""
    from cwn_b200.mp.nn import get_graph_norm
    for name, klass in [('sparse_cin_eval', SparseCIN), ('sparse_cin_eval_dim1', SparseCIN),
                        ('embed_sparse_cin_eval', EmbedSparseCIN), ('cin0_eval', CIN0),
                        ('sparse_cin_train', SparseCIN), ('embed_sparse_cin_train', EmbedSparseCIN),
                        ('embed_sparse_cin_train_nocob', EmbedSparseCIN),
                        ('ogb_embed_sparse_cin_train', OGBEmbedSparseCIN), ('cin0_train', CIN0),
                        ('cinpp_train', CINpp), ('embed_cinpp_train', EmbedCINpp)]:
        m = golden()['models'][name]
        model = klass(**m['cfg'])
        missing, unexpected = model.load_state_dict(m['state_dict'], strict=True)
        assert not missing and not unexpected
    conv = SparseCINConv(4, 4, 4, None, None, None, None, layer_dim=4, hidden=8, act_module=torch.nn.ReLU,
                         use_coboundaries=True, train_eps=True)
    assert isinstance(conv.mp_levels[0], SparseCINCochainConv) and len(conv.mp_levels) == 3
    assert conv.mp_levels[1]._up_message_form()[0] == 'cob' and isinstance(conv.mp_levels[0].eps1, torch.nn.Parameter)
    shared = CIN0(1, 3, 2, 5).convs[0]
    assert isinstance(shared, CINConv)
    assert shared.mp_levels[0].msg_up_nn is shared.mp_levels[2].msg_up_nn  # weights shared across dimensions


def test_bench_clock_sampler_sources_and_windows():
    """bench.py's clock sampler on the CPU: a fake NVML module drives the in-process poller, rows in nvidia-smi's CSV
    format are injected by hand; `stop()` must window the samples, report the median clock, the throttle reasons and how
    many samples fell inside the timed region — and never raise when a source is missing."""
    import sys
    import time
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    fake = types.SimpleNamespace(
        NVML_CLOCK_SM=1, nvmlClocksThrottleReasonHwSlowdown=8, nvmlClocksThrottleReasonHwThermalSlowdown=64,
        nvmlClocksThrottleReasonSwThermalSlowdown=32, nvmlClocksThrottleReasonSwPowerCap=4,
        nvmlInit=lambda: None, nvmlDeviceGetHandleByIndex=lambda i: ('gpu', i),
        nvmlDeviceGetMaxClockInfo=lambda h, c: 1965, nvmlDeviceGetClockInfo=lambda h, c: 1950,
        nvmlDeviceGetCurrentClocksEventReasons=lambda h: 4)
    s = bench.ClockSampler(0, nvml=fake)
    s.start()                      # nvidia-smi is absent in the build container: that source must just stay empty
    s.wait_first(2.0)
    t0 = time.perf_counter()
    time.sleep(0.05)
    t1 = time.perf_counter()
    s.samples.append((['0', '1312', '1965', '300', 'Active', 'Not Active', 'Active', 'Not Active', 'Not Active'],
                      time.perf_counter(), 'nvidia-smi'))
    info = s.stop(t0 - 1.0, time.perf_counter() + 1.0, (t0, t1))
    assert info['sm_max_mhz'] == 1965 and info['sm_mhz'] == 1950 and info['samples'] >= 10
    assert info['samples_in_timed_region'] >= 5
    assert set(info['reasons']) == {'sw_power_cap', 'hw_thermal_slowdown'}
    assert 'nvml' in info['sources'] and 'nvidia-smi' in info['sources']
    # a sampler with no source at all reports that instead of raising
    broken = types.SimpleNamespace(nvmlInit=lambda: (_ for _ in ()).throw(RuntimeError('no NVML')))
    s2 = bench.ClockSampler(0, nvml=broken)
    s2.start()
    s2.wait_first(0.05)
    out = s2.stop()
    assert out['sm_mhz'] is None or isinstance(out['sm_mhz'], int)


def test_ws_stage_geometry_queries_are_host_functions():
    """The host-side sizing of the warp-specialised CSR kernels (what `ops._ws_config` decides with): lanes per row,
    consumer threads and the pipeline depth a (F, tile height, window, message) configuration gets in 220 KB of shared
    memory — pure host arithmetic, no GPU."""
    lib = _lib.load()
    assert lib.cwn_csr_ws_consumer_threads() % 32 == 0 and lib.cwn_csr_ws_consumer_threads() >= 256
    for F, lpr_min in ((4, 2), (16, 2), (20, 4), (64, 8), (128, 16)):
        lpr = lib.cwn_csr_ws_lanes_per_row(F)
        assert lpr >= lpr_min and lpr & (lpr - 1) == 0 and lpr * 8 >= F // 4  # (one or two 128-bit vectors per lane)
    assert lib.cwn_csr_ws_lanes_per_row(6) == 0  # F % 4 != 0: no 128-bit path
    # edge-upper pass at F = 64: 128-row tiles, windows of ~180 rows, ~450 messages -> at least 4 stages
    assert lib.cwn_csr_ws_stages(64, 128, 180, 0, 448, 1, 0) >= 4
    deep = lib.cwn_csr_ws_stages(16, 256, 310, 0, 900, 1, 0)
    assert deep >= lib.cwn_csr_ws_stages(16, 256, 310, 40, 900, 2, 1) >= 2  # a second window and the row operand cost stages
    assert lib.cwn_csr_ws_stages(64, 128, 1_000_000, 0, 448, 1, 0) == 0     # a window that spans the matrix does not fit
    assert lib.cwn_csr_ws_stages(256, 128, 180, 0, 448, 1, 0) == 0          # F > 128: not served
    assert lib.cwn_csr_ws_stages(64, 130, 180, 0, 448, 1, 0) == 0           # tile height must be a multiple of 4
    # argument errors of the launch entry points are reported without touching the device
    assert lib.cwn_csr_gather_reduce_ws_f32(None, 64, None, None, 10, None, 128, 180, 448, 40000, 64, None, 64, None,
                                            None, 64, 0, None) == -1
    assert lib.cwn_csr_tile_windows(None, None, None, 40000, 128, None, None) == -1


def test_fused_reduction_query_follows_the_tensor_core_eligibility():
    """`cwn_unit_bwd_fuses_reduce` (host): 1 exactly when a group would run on the tcgen05 backward kernel — the only one
    that can take the upstream units' BatchNorm-backward sums from its g_in tiles."""
    lib = _lib.load()
    buf = torch.zeros(64 * 256, dtype=torch.float32)  # a 16-byte aligned host address to stand in for device pointers
    ptr = (buf.data_ptr() + 15) & ~15

    def desc(h, k0, k1=0, tile_rows=64, ptr_off=0):
        d = _lib.UnitBwdDesc()
        d.x0, d.ld_x0, d.k0 = ptr, k0, k0
        if k1:
            d.x1, d.ld_x1, d.k1 = ptr, k1, k1
        d.w, d.ld_w = ptr, k0 + k1
        d.z, d.ld_z, d.g_out, d.ld_g = ptr, h, ptr + ptr_off, h
        d.w_partials, d.b_partials = ptr, ptr
        d.n_rows, d.h, d.tile_rows, d.n_ctas = 640, h, tile_rows, 10
        return d

    def ask(*ds):
        arr = (_lib.UnitBwdDesc * len(ds))(*ds)
        return lib.cwn_unit_bwd_fuses_reduce(arr, len(ds))

    tc5 = os.environ.get('CWN_B200_DENSE_TC5', '1') != '0'
    assert ask(desc(64, 64), desc(64, 64, 64)) == int(tc5)
    assert ask(desc(64, 64), desc(64, 48)) == 0            # K = 48 is not a tensor-core shape: the whole group falls back
    assert ask(desc(32, 64)) == 0                          # h = 32
    assert ask(desc(64, 64, tile_rows=32)) == 0            # 32-row tiles
    assert ask(desc(64, 64, ptr_off=4)) == 0               # misaligned gradient
    assert lib.cwn_unit_bwd_fuses_reduce(None, 0) == 0
