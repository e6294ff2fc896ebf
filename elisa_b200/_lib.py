"""ctypes binding of include/elisa_b200.h.  There is no CPU fallback: a missing library is an error."""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ELISA_B200_LIB: another build of the library (A/B runs of kernel variants, tools/ab_lib.sh)
LIB_PATH = os.environ.get('ELISA_B200_LIB') or os.path.join(HERE, 'lib', 'libelisa_b200.so')

ABI_VERSION = 6
MAX_COMP_PARAMS = 5
MAX_COMPONENTS = 16
MAX_PARAMS = 32
OP_ADD = -1
OP_MUL = -2
OP_NORM0 = -16
MAX_NORMS = 2
MAX_SHIFTS = 2
SHIFT_Z, SHIFT_ZA, SHIFT_V = 1, 2, 3
NORM = {'PhFlux': 0, 'EnFlux': 1}
FLAG_NO_JIT = 1
FLAG_REQUIRE_JIT = 2
FLAG_FOLD_DMMA = 4
FLAG_FOLD_I8 = 8
FLAG_NO_BALANCE = 16
FLAG_NO_SMALL_BATCH_KERNEL = 32
FLAG_FOLD_SLICES_SHIFT = 8
FLAG_FOLD_SLICES_BWD_SHIFT = 12
METHOD = {'trapz': 0, 'simpson': 1}
STAT = {'chi2': 0, 'cstat': 1, 'pstat': 2, 'pgstat': 3, 'wstat': 4}
ERRORS = {-1: 'ELISA_ERR_INVALID', -2: 'ELISA_ERR_CUDA', -3: 'ELISA_ERR_UNSUPPORTED', -4: 'ELISA_ERR_NOMEM'}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class ComponentDesc(C.Structure):
    _fields_ = [
        ('kind', C.c_int32),
        ('method', C.c_int32),
        ('param_src', C.c_int32 * MAX_COMP_PARAMS),
        ('param_const', C.c_double * MAX_COMP_PARAMS),
        ('aux', C.c_double * 2),
        ('table_energy', c_double_p),
        ('table_xsect', c_double_p),
        ('table_len', C.c_int32),
        ('table_log', C.c_int32),
        ('grid_scale', C.c_double),
        ('out_scale', C.c_double),
        ('norm_grid_scale', C.c_double),
        ('norm_out_scale', C.c_double),
        ('shift_kind', C.c_int32 * MAX_SHIFTS),
        ('shift_src', C.c_int32 * MAX_SHIFTS),
    ]


class NormDesc(C.Structure):
    _fields_ = [
        ('kind', C.c_int32),
        ('f_src', C.c_int32),
        ('f_const', C.c_double),
        ('grid', c_double_p),
        ('n_grid', C.c_int32),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ('n_components', C.c_int32),
        ('components', C.POINTER(ComponentDesc)),
        ('n_ops', C.c_int32),
        ('ops', c_int32_p),
        ('n_norms', C.c_int32),
        ('norms', C.POINTER(NormDesc)),
    ]


class SpectrumDesc(C.Structure):
    _fields_ = [
        ('n_energies', C.c_int32),
        ('n_channels', C.c_int32),
        ('photon_egrid', c_double_p),
        ('response_dense', c_double_p),
        ('csr_indptr', c_int32_p),
        ('csr_indices', c_int32_p),
        ('csr_values', c_double_p),
        ('area_scale', c_double_p),
        ('exposure', C.c_double),
        ('stat', C.c_int32),
        ('on_counts', c_double_p),
        ('on_error', c_double_p),
        ('off_counts', c_double_p),
        ('off_error', c_double_p),
        ('back_ratio', c_double_p),
        ('model', ModelDesc),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32),
        ('device', C.c_int32),
        ('n_params', C.c_int32),
        ('n_spectra', C.c_int32),
        ('spectra', C.POINTER(SpectrumDesc)),
        ('chunk', C.c_int64),
        ('flags', C.c_int32),
    ]


# every symbol include/elisa_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    'elisa_b200_plan_create': (C.c_int, [C.POINTER(PlanDesc), C.POINTER(_VP)]),
    'elisa_b200_plan_destroy': (None, [_VP]),
    'elisa_b200_loglike_dev': (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP]),
    'elisa_b200_loglike_grad_dev': (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP]),
    'elisa_b200_normal_eq_dev': (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    'elisa_b200_sites_dev': (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP]),
    'elisa_b200_unfold_dev': (C.c_int, [_VP, C.c_int32, _VP, C.c_int64, _VP, _VP]),
    'elisa_b200_loglike_grad_host': (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP]),
    'elisa_b200_plan_set_channel_width': (C.c_int, [_VP, C.c_int32, _VP]),
    'elisa_b200_plan_set_profiling': (C.c_int, [_VP, C.c_int]),
    'elisa_b200_plan_get_profile': (C.c_int, [_VP, c_double_p, C.POINTER(C.c_int64)]),
    'elisa_b200_plan_workspace_bytes': (C.c_int64, [_VP]),
    'elisa_b200_plan_launch_count': (C.c_int64, [_VP]),
    'elisa_b200_plan_chunk': (C.c_int64, [_VP]),
    'elisa_b200_specialise_check': (C.c_int64, [C.POINTER(ModelDesc), C.c_int32, C.c_char_p, C.c_int64]),
    'elisa_b200_plan_is_specialised': (C.c_int, [_VP, C.c_int32]),
    'elisa_b200_plan_fold_slices': (C.c_int, [_VP, C.c_int32]),
    'elisa_b200_plan_set_fold': (C.c_int, [_VP, C.c_int]),
    'elisa_b200_plan_guard_fallbacks': (C.c_int64, [_VP]),
    'elisa_b200_plan_guard_stats': (C.c_int, [_VP, C.POINTER(C.c_int64)]),
    'elisa_b200_plan_fold_guard': (C.c_int, [_VP, C.POINTER(C.c_int64), c_double_p, C.c_int]),
    'elisa_b200_sliced_gemm_dev': (C.c_int, [_VP, C.c_int64, _VP, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            _VP, C.c_int64, c_double_p, _VP]),
    'elisa_b200_plan_rebalance': (C.c_int, [_VP]),
    'elisa_b200_loglike_grad_guarded_dev': (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, C.POINTER(C.c_int64)]),
    'elisa_b200_plan_guard_snapshot': (C.c_int, [_VP, _VP]),
    'elisa_b200_plan_guard_poll': (C.c_int, [_VP, C.c_int, C.POINTER(C.c_int64), c_double_p]),
    'elisa_b200_jacobian_dev': (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    'elisa_b200_plan_set_debug': (C.c_int, [_VP, C.c_int32]),
    'elisa_b200_sample_dev': (C.c_int, [_VP, C.c_int32, C.c_int32, _VP, C.c_int64, C.c_int64, C.c_uint64, _VP, C.c_int64, _VP]),
    'elisa_b200_csr_fold_dev': (C.c_int, [_VP, C.c_int32, _VP, C.c_int64, C.c_int64, _VP, C.c_int64, _VP]),
    'elisa_b200_plan_fold_slices_bwd': (C.c_int, [_VP, C.c_int32]),
    'elisa_b200_plan_small_batch_kernel': (C.c_int, [_VP, C.c_int32]),
    'elisa_b200_plan_fold_products': (C.c_int, [_VP, C.c_int32, c_double_p]),
    'elisa_b200_abi_version': (C.c_int, []),
    'elisa_b200_peer_alloc': (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    'elisa_b200_peer_open': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'elisa_b200_peer_close': (C.c_int, [_VP]),
    'elisa_b200_peer_free': (C.c_int, [_VP]),
    'elisa_b200_plan_set_gather': (C.c_int, [_VP, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_void_p), C.c_int64]),
    'elisa_b200_sizeof': (C.c_int64, [C.c_int32]),
    'elisa_b200_last_error': (C.c_char_p, []),
}

_lib = None


class ElisaB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libelisa_b200.so (built in-tree by ``elisa_b200.build``); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ElisaB200Error(
            f'{LIB_PATH} is missing: build it with `python -m elisa_b200.build` '
            '(elisa_b200 has no CPU fallback and will not run without its CUDA library)'
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.elisa_b200_abi_version() != ABI_VERSION:
        raise ElisaB200Error('libelisa_b200.so ABI version mismatch: rebuild with `python -m elisa_b200.build --force`')
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().elisa_b200_last_error().decode()
        raise ElisaB200Error(f'{ERRORS.get(rc, rc)}: {msg}')
