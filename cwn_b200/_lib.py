"""ctypes binding of `libcwn_b200.so` (declared in include/cwn_b200.h). There is NO fallback: if the library
cannot be loaded, every message-passing op raises."""
import ctypes
import os

from cwn_b200 import build as _build

_c_f32p = ctypes.c_void_p
_c_i32p = ctypes.c_void_p
_c_i64p = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_f32 = ctypes.c_float
_vp = ctypes.c_void_p

class PlanDesc(ctypes.Structure):
    """cwn_plan_desc of include/cwn_b200.h."""
    _fields_ = [('key', ctypes.c_void_p), ('pay0', ctypes.c_void_p), ('pay1', ctypes.c_void_p),
                ('E', ctypes.c_int64), ('n_rows', ctypes.c_int64), ('rowptr', ctypes.c_void_p),
                ('perm', ctypes.c_void_p), ('pay0_sorted', ctypes.c_void_p), ('pay1_sorted', ctypes.c_void_p)]


_P, _I64, _I32, _F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float


class LinearDesc(ctypes.Structure):
    """cwn_linear_desc"""
    _fields_ = [('x0', _P), ('ld_x0', _I64), ('k0', _I32), ('x1', _P), ('ld_x1', _I64), ('k1', _I32),
                ('in_mean0', _P), ('in_scale0', _P), ('in_beta0', _P), ('in_mean1', _P), ('in_scale1', _P),
                ('in_beta1', _P), ('in_act', _I32), ('w', _P), ('ld_w', _I64), ('bias', _P), ('z', _P),
                ('ld_z', _I64), ('stats', _P), ('n_rows', _I64), ('h', _I32),
                ('bn_gamma', _P), ('bn_eps', _F), ('bn_momentum', _F), ('bn_training', _I32),
                ('bn_running_mean', _P), ('bn_running_var', _P), ('bn_num_batches_tracked', _P),
                ('bn_mean', _P), ('bn_scale', _P), ('bn_rstd', _P), ('counter', _P), ('tile_rows', _I32),
                ('n_rows_live', _P)]


class HeadDim(ctypes.Structure):
    """cwn_head_dim"""
    _fields_ = [('x', _P), ('ld_x', _I64), ('rowptr', _P), ('perm', _P), ('w1', _P), ('b1', _P), ('pooled', _P),
                ('z', _P), ('g_z', _P), ('g_x', _P), ('ld_gx', _I64), ('g_w1', _P), ('g_b1', _P),
                ('accumulate', _I32)]


class BNDesc(ctypes.Structure):
    """cwn_bn_desc"""
    _fields_ = [('stats', _P), ('n_tiles', _I32), ('n_rows', _I64), ('h', _I32), ('gamma', _P), ('beta', _P),
                ('eps', _F), ('momentum', _F), ('training', _I32), ('running_mean', _P), ('running_var', _P),
                ('num_batches_tracked', _P), ('mean', _P), ('scale', _P), ('rstd', _P), ('tile_rows', _I32)]


class BNActDesc(ctypes.Structure):
    """cwn_bn_act_desc"""
    _fields_ = [('z', _P), ('ld_z', _I64), ('mean', _P), ('scale', _P), ('beta', _P), ('act', _I32), ('out', _P),
                ('ld_out', _I64), ('n_rows', _I64), ('h', _I32)]


class UnitBwdDesc(ctypes.Structure):
    """cwn_unit_bwd_desc"""
    _fields_ = [('x0', _P), ('ld_x0', _I64), ('k0', _I32), ('x1', _P), ('ld_x1', _I64), ('k1', _I32),
                ('in_mean0', _P), ('in_scale0', _P), ('in_beta0', _P), ('in_mean1', _P), ('in_scale1', _P),
                ('in_beta1', _P), ('in_act', _I32), ('w', _P), ('ld_w', _I64), ('z', _P), ('ld_z', _I64),
                ('has_bn', _I32), ('act', _I32), ('mean', _P), ('scale', _P), ('rstd', _P), ('beta', _P),
                ('g_out', _P), ('ld_g', _I64), ('red_partials', _P), ('c1', _P), ('c2', _P), ('g_gamma', _P),
                ('g_beta', _P), ('accumulate_affine', _I32), ('g_in0', _P), ('ld_gi0', _I64), ('g_in1', _P),
                ('ld_gi1', _I64), ('w_partials', _P), ('b_partials', _P), ('n_ctas', _I32), ('g_w', _P),
                ('ld_gw', _I64), ('g_b', _P), ('accumulate_w', _I32), ('n_rows', _I64), ('h', _I32),
                ('counter', _P), ('tile_rows', _I32), ('accumulate_in', _I32), ('n_rows_live', _P),
                ('next_rstd0', _P), ('next_red0', _P), ('next_c1_0', _P), ('next_c2_0', _P), ('next_g_gamma0', _P),
                ('next_g_beta0', _P), ('next_rstd1', _P), ('next_red1', _P), ('next_c1_1', _P), ('next_c2_1', _P),
                ('next_g_gamma1', _P), ('next_g_beta1', _P), ('next_accumulate_affine0', _I32),
                ('next_accumulate_affine1', _I32), ('next_counter', _P)]


class CollateJob(ctypes.Structure):
    """cwn_collate_job"""
    _fields_ = [('src', _P), ('dst', _P), ('src_start', _P), ('dst_start', _P), ('add', _P), ('n_segments', _I32),
                ('kind', _I32), ('row_elems', _I32), ('n_out', _I64)]


MAX_GROUP = 8

_SIGNATURES = {
    'cwn_version': (ctypes.c_char_p, []),
    'cwn_last_error_string': (ctypes.c_char_p, []),
    'cwn_launch_count': (ctypes.c_ulonglong, []),
    'cwn_debug_force_generic_dense': (ctypes.c_int, [ctypes.c_int32]),
    'cwn_readout_head_fwd': (ctypes.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                            _vp]),
    'cwn_readout_head_bwd': (ctypes.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                            _vp, _i32, _vp]),
    'cwn_readout_head_bwd_parts': (ctypes.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                                  _vp, _i32, _i32, _vp]),
    'cwn_csr_plan_workspace_bytes': (ctypes.c_size_t, [_i64, _i64]),
    'cwn_csr_plan_build': (ctypes.c_int, [_c_i64p, _c_i64p, _c_i64p, _i64, _i64, _c_i32p, _c_i32p, _c_i32p,
                                          _c_i32p, _c_i32p, _vp, ctypes.c_size_t, _vp]),
    'cwn_csr_plan_small_capacity': (ctypes.c_int64, []),
    'cwn_csr_plan_build_small': (ctypes.c_int, [ctypes.POINTER(PlanDesc), _i32, _c_i32p, _vp]),
    'cwn_csr_gather_reduce_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i32p, _c_i32p, _i64, _i32, _c_f32p, _i64,
                                                 _c_f32p, _c_f32p, _i64, _i32, _vp]),
    'cwn_csr_gather_reduce2_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i32p, _c_i32p, _i64, _i32, _c_f32p, _i64,
                                                  _c_f32p, _c_f32p, _i64, _c_f32p, _c_f32p, _i64, _vp]),
    'cwn_csr_gather_max_arg_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _i32, _c_f32p, _i64,
                                                  _c_i32p, _vp]),
    'cwn_csr_max_bwd_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _c_i32p, _i64, _i32, _c_f32p, _i64,
                                           _vp]),
    'cwn_gather_rows_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i64p, _i64, _i32, _f32, _c_f32p, _i64, _vp]),
    'cwn_csr_cob_fwd_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _i32,
                                           _i32, _c_f32p, _i64, _c_f32p, _c_f32p, _i64, _vp]),
    'cwn_csr_cob_bwd_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p,
                                           _c_i32p, _i64, _i32, _i32, _c_f32p, _i64, _vp]),
    'cwn_csr_gather_reduce_f64': (ctypes.c_int, [_vp, _i64, _c_i32p, _c_i32p, _i64, _i32, _vp, _i64, _vp, _vp, _i64, _i32, _vp]),
    'cwn_gather_rows_f64': (ctypes.c_int, [_vp, _i64, _c_i64p, _i64, _i32, ctypes.c_double, _vp, _i64, _vp]),
    'cwn_csr_cob_fwd_f64': (ctypes.c_int, [_vp, _i64, _vp, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _i32, _i32, _vp, _i64,
                                           _vp, _vp, _i64, _vp]),
    'cwn_csr_cob_bwd_f64': (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _i32, _i32,
                                           _vp, _i64, _vp]),
    'cwn_csr_tile_windows': (ctypes.c_int, [_c_i32p, _c_i32p, _c_i32p, _i64, _i32, _c_i32p, _vp]),
    'cwn_csr_ws_consumer_threads': (ctypes.c_int, []),
    'cwn_csr_ws_lanes_per_row': (ctypes.c_int, [_i32]),
    'cwn_csr_ws_stages': (ctypes.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    'cwn_csr_gather_reduce_ws_f32': (ctypes.c_int, [_c_f32p, _i64, _c_i32p, _c_i32p, _i64, _c_i32p, _i32, _i32, _i32,
                                                    _i64, _i32, _c_f32p, _i64, _c_f32p, _c_f32p, _i64, _i32, _vp]),
    'cwn_csr_cob_fwd_ws_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _c_i32p,
                                              _i32, _i32, _i32, _i32, _i64, _i32, _i32, _c_f32p, _i64, _c_f32p,
                                              _c_f32p, _i64, _vp]),
    'cwn_csr_cob_bwd_ws_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p,
                                              _i64, _c_i32p, _i32, _i32, _i32, _i32, _i64, _i32, _i32, _c_f32p, _i64,
                                              _vp]),
    'cwn_cin_msg_sq_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _i64, _i32, _i32,
                                          _c_f32p, _c_f32p, _i64, _vp]),
    'cwn_cin_msg_bwd_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i64, _c_f32p, _i64, _c_i32p, _c_i32p, _c_i32p, _i64,
                                           _i32, _i32, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _i64, _vp]),
    'cwn_linear_fwd_grouped': (ctypes.c_int, [ctypes.POINTER(LinearDesc), _i32, _vp]),
    'cwn_bn_finalize_grouped': (ctypes.c_int, [ctypes.POINTER(BNDesc), _i32, _vp]),
    'cwn_bn_act_grouped': (ctypes.c_int, [ctypes.POINTER(BNActDesc), _i32, _vp]),
    'cwn_unit_bwd_reduce_grouped': (ctypes.c_int, [ctypes.POINTER(UnitBwdDesc), _i32, _vp]),
    'cwn_unit_bwd_finalize_grouped': (ctypes.c_int, [ctypes.POINTER(UnitBwdDesc), _i32, _vp]),
    'cwn_unit_bwd_grouped': (ctypes.c_int, [ctypes.POINTER(UnitBwdDesc), _i32, _vp]),
    'cwn_unit_bwd_fuses_reduce': (ctypes.c_int, [ctypes.POINTER(UnitBwdDesc), _i32]),
    'cwn_wgrad_finalize_grouped': (ctypes.c_int, [ctypes.POINTER(UnitBwdDesc), _i32, _vp]),
    'cwn_collate': (ctypes.c_int, [ctypes.POINTER(CollateJob), _i32, _vp]),
    'cwn_adam_step_f32': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _i32, _vp]),
    'cwn_allreduce_adam_step_f32': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i64, _f32, _f32, _f32, _f32,
                                                   _f32, _vp, _vp, _i32, _vp, _vp]),
    'cwn_check_index_range': (ctypes.c_int, [_c_i64p, _i64, _i64, _c_i32p, _vp]),
}

_lib = None


def exported_symbols():
    """Every entry point include/cwn_b200.h declares."""
    return sorted(_SIGNATURES)


def load():
    """Load (building in-tree first if the `.so` is absent or stale and nvcc is present). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get('CWN_B200_LIB', _build.LIB)
    try:
        path = _build.build_library()
    except RuntimeError:
        if not os.path.exists(path):
            raise
    if not os.path.exists(path):
        raise RuntimeError(f'cwn_b200: CUDA extension {path} is missing; there is no CPU or PyTorch fallback')
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=''):
    if rc == 0:
        return
    msg = load().cwn_last_error_string().decode()
    if rc < 0:
        raise ValueError(f'cwn_b200 {what}: argument error {rc}: {msg}')
    raise RuntimeError(f'cwn_b200 {what}: CUDA error {rc}: {msg}')


def launch_count():
    return int(load().cwn_launch_count())
