"""ctypes binding of libatomistica_b200.so (the C-ABI in include/atomistica_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and if no
CUDA device is present every compute entry point raises RuntimeError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libatomistica_b200.so')

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ssize_p = C.POINTER(C.c_ssize_t)


class AtxSpline(C.Structure):
    _fields_ = [('n', C.c_int), ('x0', C.c_double), ('dx', C.c_double)] + \
        [(k, c_double_p) for k in ('y', 'coeff1', 'coeff2', 'coeff3', 'dcoeff1', 'dcoeff2', 'dcoeff3')]


MAX_EL, MAX_PAIRS = 3, 6


class AtxBopParams(C.Structure):
    _fields_ = [('kind', C.c_int), ('nel', C.c_int), ('Z', C.c_int * MAX_EL)] + \
        [(k, C.c_double * MAX_PAIRS) for k in ('A', 'B', 'xi', 'lambda_', 'mu', 'omega', 'mubo')] + \
        [('m', C.c_int * MAX_PAIRS)] + \
        [(k, C.c_double * MAX_PAIRS) for k in ('D0', 'r0', 'S', 'pbeta', 'gamma', 'pc', 'pd', 'ph', 'pn',
                                               'r1', 'r2')] + \
        [(k, C.c_double * MAX_EL) for k in ('beta', 'n', 'c', 'd', 'h', 'eta', 'delta',
                                            'c1', 'c2', 'c3', 'c4', 'c5')]


class AtxBopScreening(C.Structure):
    _fields_ = [(k, C.c_double * MAX_PAIRS) for k in ('or1', 'or2', 'bor1', 'bor2', 'Cmin', 'Cmax')]


class AtxJuslinParams(C.Structure):
    _fields_ = [('nel', C.c_int), ('Z', C.c_int * MAX_EL)] + \
        [(k, C.c_double * 9) for k in ('D0', 'r0', 'S', 'beta', 'gamma', 'c', 'd', 'h', 'n', 'r1', 'r2')] + \
        [('alpha', C.c_double * 27), ('omega', C.c_double * 27), ('m', C.c_int * 27)]


class AtxJuslinScreening(C.Structure):
    _fields_ = [(k, C.c_double * 9) for k in ('or1', 'or2', 'bor1', 'bor2', 'Cmin', 'Cmax')]


class AtxPairParams(C.Structure):
    _fields_ = [('kind', C.c_int), ('p', C.c_double * 8), ('shift', C.c_int)]


class AtxRebo2Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        'cc_B1', 'cc_B2', 'cc_B3', 'cc_beta1', 'cc_beta2', 'cc_beta3', 'cc_Q', 'cc_A', 'cc_alpha',
        'ch_B1', 'ch_beta1', 'ch_Q', 'ch_A', 'ch_alpha', 'hh_B1', 'hh_beta1', 'hh_Q', 'hh_A', 'hh_alpha')] + [
        ('cc_g_theta', C.c_double * 6), ('cc_g1_coeff', C.c_double * 18), ('cc_g2_coeff', C.c_double * 18),
        ('spgh', C.c_double * 18), ('igh', C.c_int * 25), ('conalp', C.c_double), ('conear', C.c_double * 36),
        ('conpe', C.c_double * 3), ('conan', C.c_double * 3), ('conpf', C.c_double * 3),
        ('cut_in_l', C.c_double * 10), ('cut_in_h', C.c_double * 10), ('cut_in_h2', C.c_double * 10),
        ('with_dihedral', C.c_int)] + [(k, c_double_p) for k in ('Fcc', 'Fch', 'Fhh', 'Tcc', 'Pcc', 'Pch')]


class AtxRebo2Screening(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('cc_ar_r1', 'cc_ar_r2', 'cc_bo_r1', 'cc_bo_r2', 'cc_nc_r1', 'cc_nc_r2',
                                          'Cmin', 'Cmax')]


# every symbol include/atomistica_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    'atx_ctx_create', 'atx_ctx_destroy', 'atx_ctx_synchronize', 'atx_last_error', 'atx_version',
    'atx_kernel_launches',
    'atx_particles_create', 'atx_particles_destroy', 'atx_particles_set_cell', 'atx_particles_set_positions',
    'atx_particles_set_elements', 'atx_particles_set_positions_device',
    'atx_neighbors_create', 'atx_neighbors_destroy', 'atx_neighbors_request_interaction_range',
    'atx_neighbors_set_verlet_shell', 'atx_neighbors_update', 'atx_neighbors_rebuild', 'atx_neighbors_get_info', 'atx_neighbors_get_counters', 'atx_neighbors_get_interaction_range', 'atx_neighbors_request_interaction_range_pair', 'atx_neighbors_get_pair_range',
    'atx_neighbors_copy_to_host', 'atx_neighbors_set_external',
    'atx_neighbors_coordination_numbers', 'atx_neighbors_pair_distribution', 'atx_neighbors_angle_distribution',
    'atx_neighbors_bond_angles',
    'atx_eam_create', 'atx_eam_create_funcfl', 'atx_eam_destroy', 'atx_eam_bind_to', 'atx_eam_energy_and_forces',
    'atx_eam_set_store_outputs', 'atx_bop_set_store_outputs', 'atx_rebo2_set_store_outputs',
    'atx_bop_create', 'atx_bop_create_screened', 'atx_bop_create_juslin', 'atx_bop_create_juslin_screened', 'atx_bop_destroy', 'atx_bop_bind_to', 'atx_bop_energy_and_forces',
    'atx_pair_create', 'atx_pair_destroy', 'atx_pair_bind_to', 'atx_pair_energy_and_forces',
    'atx_pair_set_store_outputs',
    'atx_rebo2_create', 'atx_rebo2_create_screened', 'atx_rebo2_destroy', 'atx_rebo2_bind_to', 'atx_rebo2_energy_and_forces',
    'atx_md_create', 'atx_md_destroy', 'atx_md_run', 'atx_md_get_state', 'atx_md_get_stats',
    'atx_dd_get_unique_id', 'atx_dd_create', 'atx_dd_destroy', 'atx_dd_md_create', 'atx_dd_md_destroy',
    'atx_dd_md_run', 'atx_dd_md_get_count', 'atx_dd_md_get_state', 'atx_dd_md_get_stats', 'atx_dd_md_get_profile',
    'atx_profile_enable', 'atx_profile_read', 'atx_measure_fp64_peak', 'atx_measure_copy_bandwidth',
    'atx_host_alloc_pinned', 'atx_host_free_pinned', 'atx_host_register', 'atx_host_unregister',
    'atx_host_spline_init', 'atx_host_gaussn', 'atx_host_table2d_init', 'atx_host_table3d_init',
    'atx_host_rebo2_g_spline',
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'libatomistica_b200.so is missing (%s). Build it with `python -m atomistica_b200.build`; '
                'there is no CPU fallback.' % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.atx_version.restype = C.c_char_p
        _lib.atx_kernel_launches.restype = C.c_longlong
    return _lib


def last_error():
    buf = C.create_string_buffer(2048)
    lib().atx_last_error(buf, 2048)
    return buf.value.decode(errors='replace')


def check(err):
    """Error convention of the reference's C layer (src/python/c/py_f.c:36-47): RuntimeError."""
    if err != 0:
        raise RuntimeError(last_error() or 'atomistica_b200 error %d' % err)


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


_ctx = {}


def context(device=0):
    """One context (device + stream) per device ordinal, created on first use."""
    if device not in _ctx:
        h = C.c_void_p()
        check(lib().atx_ctx_create(C.c_int(device), C.byref(h)))
        _ctx[device] = h
    return _ctx[device]


def kernel_launches(reset=False):
    return int(lib().atx_kernel_launches(C.c_int(1 if reset else 0)))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _PinnedBlock:
    """owner of one cudaMallocHost allocation; freed when the last numpy view of it is gone"""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().atx_host_alloc_pinned(C.c_size_t(nbytes), C.byref(self.ptr)))

    def __del__(self):
        try:
            lib().atx_host_free_pinned(self.ptr)
        except Exception:
            pass


class PinnedArray:
    """float64 numpy array backed by page-locked host memory (cudaMallocHost through the C ABI).
    `array` (and every view of it handed to a caller) keeps the allocation alive."""

    def __init__(self, shape):
        n = int(np.prod(shape))
        context(0)     # pinned allocation needs a CUDA context
        block = _PinnedBlock(8 * max(n, 1))
        buf = (C.c_double * max(n, 1)).from_address(block.ptr.value)
        buf._block = block          # numpy keeps `buf` as the base object of the array
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        self.array[...] = 0.0
