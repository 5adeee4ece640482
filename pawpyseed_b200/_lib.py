"""ctypes binding of libpawb200.so (the C ABI in include/pawpyseed_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pawpyseed_b200/csrc``.
There is no Python or CPU fallback: if the shared object is missing or no B200 is
visible, calls raise :class:`PAWpyError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpawb200.so")


class PAWpyError(Exception):
    """Same name as pawpyseed.core.utils.PAWpyError (utils.py:9-20)."""


c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class Timers(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "h2d_ms", "scatter_ms", "fft_ms", "project_ms", "table_ms", "gemm_pseudo_ms",
        "gemm_aug_ms", "augment_ms", "d2h_ms")] + [(n, C.c_longlong) for n in (
        "launches", "boxes_scattered", "boxes_fft", "slots_projected", "sphere_samples")]


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    "pawb200_last_error": (C.c_char_p, []),
    "pawb200_clear_error": (None, []),
    "pawb200_device_check": (C.c_int, []),
    "pawb200_version": (C.c_char_p, []),
    "pawb200_read_wavefunctions": (C.c_void_p, [C.c_char_p, c_dbl_p]),
    "pawb200_read_wavefunctions_from_str": (C.c_void_p, [C.c_void_p, c_dbl_p]),
    "pawb200_free_pswf": (None, [C.c_void_p]),
    "pawb200_expand_symm_wf": (C.c_void_p, [C.c_void_p, C.c_int, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p]),
    "pawb200_get_coefficients": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pawb200_get_kpoint": (C.c_int, [C.c_void_p, C.c_int, c_dbl_p, c_dbl_p]),
    "pawb200_get_nband": (C.c_int, [C.c_void_p]),
    "pawb200_get_nwk": (C.c_int, [C.c_void_p]),
    "pawb200_get_nspin": (C.c_int, [C.c_void_p]),
    "pawb200_is_ncl": (C.c_int, [C.c_void_p]),
    "pawb200_get_encut": (C.c_double, [C.c_void_p]),
    "pawb200_get_energy": (C.c_double, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pawb200_get_occ": (C.c_double, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pawb200_get_occs": (C.c_void_p, [C.c_void_p]),
    "pawb200_set_num_sites": (None, [C.c_void_p, C.c_int]),
    "pawb200_free_ptr": (None, [C.c_void_p]),
    "pawb200_get_projector_list": (C.c_void_p, [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p,
                                                c_dbl_p, c_dbl_p, c_dbl_p, C.c_double]),
    "pawb200_free_ppot_list": (None, [C.c_void_p, C.c_int]),
    "pawb200_setup_projections": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_int_p,
                                         c_int_p, c_dbl_p]),
    "pawb200_pseudoprojection": (None, [c_dbl_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pawb200_overlap_setup_real": (None, [C.c_void_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p,
                                          c_dbl_p, c_int_p, c_int_p, c_int_p, c_int_p,
                                          C.c_int, C.c_int, C.c_int]),
    "pawb200_compensation_terms": (None, [c_dbl_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_int, C.c_int] + [c_int_p] * 6 +
                                   [c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, C.c_int]),
    "pawb200_overlap_setup_recip": (None, [C.c_void_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p,
                                           c_dbl_p, c_int_p, c_int_p, c_int_p, c_int_p,
                                           C.c_int, C.c_int, C.c_int]),
    "pawb200_compensation_terms_recip": (None, [c_dbl_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_int, C.c_int, C.c_int] + [c_int_p] * 6 +
                                         [c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, C.c_int]),
    "pawb200_realspace_state": (None, [c_dbl_p, C.c_int, C.c_int, C.c_void_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_ncl_realspace_state": (None, [c_dbl_p, C.c_int, C.c_int, C.c_void_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_remove_phase": (None, [c_dbl_p, C.c_int, C.c_void_p, c_int_p]),
    "pawb200_ae_state_density": (None, [c_dbl_p, C.c_int, C.c_int, C.c_void_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_ae_chg_density": (None, [c_dbl_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_ae_chg_density_bands": (None, [c_dbl_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p, C.c_int, C.c_int]),
    "pawb200_ncl_ae_chg_density": (None, [c_dbl_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_write_volumetric": (None, [C.c_char_p, c_dbl_p, c_int_p, C.c_double]),
    "pawb200_project_realspace_state": (None, [c_dbl_p, C.c_int, C.c_void_p, C.c_void_p, c_int_p, c_int_p, c_dbl_p,
                                              c_int_p, c_dbl_p]),
    "pawb200_fft3d": (None, [c_dbl_p, c_int_p, c_dbl_p, c_dbl_p, c_int_p, C.c_void_p, C.c_int, c_int_p]),
    "pawb200_fwd_fft3d": (None, [c_dbl_p, c_int_p, c_dbl_p, c_dbl_p, c_int_p, C.c_void_p, C.c_int, c_int_p]),
    "pawb200_legendre": (C.c_double, [C.c_int, C.c_int, C.c_double]),
    "pawb200_Ylm": (None, [C.c_int, C.c_int, C.c_double, C.c_double, c_dbl_p]),
    "pawb200_Ylm2": (None, [C.c_int, C.c_int, C.c_double, C.c_double, c_dbl_p]),
    "pawb200_frac_to_cartesian": (None, [c_dbl_p, c_dbl_p]),
    "pawb200_cartesian_to_frac": (None, [c_dbl_p, c_dbl_p]),
    "pawb200_spline_coeff": (C.c_void_p, [c_dbl_p, c_dbl_p, C.c_int]),
    "pawb200_proj_interpolate": (C.c_double, [C.c_double, C.c_double, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p]),
    "pawb200_wave_interpolate": (C.c_double, [C.c_double, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p]),
    "pawb200_spline_integral": (C.c_double, [c_dbl_p, c_dbl_p, c_dbl_p, C.c_int]),
    "pawb200_spherical_bessel_transform": (None, [C.c_double, C.c_int, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "pawb200_reciprocal_offsite_wave_overlap": (None, [c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int,
                                                       c_dbl_p, c_dbl_p, c_dbl_p, C.c_int,
                                                       C.c_int, C.c_int, C.c_int, C.c_int, c_dbl_p]),
    "pawb200_momentum_grid_size": (None, [C.c_void_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, C.c_double]),
    "pawb200_get_momentum_grid": (C.c_int, [c_int_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "pawb200_grid_bounds": (None, [c_int_p, c_int_p, c_int_p, C.c_int]),
    "pawb200_list_to_grid_map": (None, [c_int_p, c_int_p, c_int_p, c_int_p, C.c_int]),
    "pawb200_get_all_transforms": (C.c_void_p, [C.c_void_p, C.c_double]),
    "pawb200_free_density_ft_elem_list": (None, [C.c_void_p, C.c_int]),
    "pawb200_get_momentum_matrix": (None, [c_dbl_p, C.c_int, c_int_p, C.c_void_p, c_int_p, c_dbl_p] + [C.c_int] * 6 +
                                    [C.c_void_p, C.c_double]),
    "pawb200_fullwf_reciprocal": (None, [c_dbl_p, c_int_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_dbl_p]),
    "pawb200_quick_overlap": (None, [c_int_p, c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p]),
    "pawb200_projection_matrix": (None, [c_dbl_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int] + [c_int_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pawb200_projection_matrix_dev": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int] + [c_int_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pawb200_set_kappa_range": (None, [C.c_void_p, C.c_int, C.c_int]),
    "pawb200_set_read_shard": (None, [C.c_int, C.c_int]),
    "pawb200_set_band_shard": (None, [C.c_int, C.c_int]),
    "pawb200_get_device_buffer": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_long), c_int_p, c_int_p,
                                               c_int_p]),
    "pawb200_set_host_threads": (None, [C.c_int]),
    "pawb200_set_async_ingest": (None, [C.c_int]),
    "pawb200_get_projections": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dbl_p]),
    "pawb200_num_projections": (C.c_int, [C.c_void_p, C.c_int]),
    "pawb200_get_channel_index": (C.c_int, [C.c_void_p, c_int_p]),
    "pawb200_get_site_indices": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.c_int]),
    "pawb200_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "pawb200_free_pinned": (None, [C.c_void_p]),
    "pawb200_get_timers": (None, [C.POINTER(Timers)]),
    "pawb200_reset_timers": (None, []),
}


def lib():
    """Load libpawb200.so; raises PAWpyError (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PAWpyError(
                "libpawb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C pawpyseed_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS.keys())


def check():
    """Raise PAWpyError if the last C call on this thread failed."""
    msg = lib().pawb200_last_error()
    if msg:
        lib().pawb200_clear_error()
        raise PAWpyError(msg.decode())


def ip(a):
    return None if a is None or len(a) == 0 else a.ctypes.data_as(c_int_p)


def dp(a):
    return a.ctypes.data_as(c_dbl_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_pinned_free = {}   # nbytes -> [ptr]: page-locked blocks are recycled (cudaMallocHost / cudaFreeHost synchronise)


def _pinned_release(nbytes, ptr):
    cached = _pinned_free.setdefault(nbytes, [])
    if len(cached) < 4:
        cached.append(ptr)          # recycled by the next result of this size
    else:
        try:
            lib().pawb200_free_pinned(ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """Uninitialised numpy array in page-locked host memory; the block returns to a free list with the array."""
    import weakref
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = max(count * dtype.itemsize, 1)
    cached = _pinned_free.get(n)
    if cached:
        ptr = cached.pop()
    else:
        ptr = lib().pawb200_alloc_pinned(n)
        check()
        if not ptr:
            raise PAWpyError("pinned allocation of %d bytes failed" % n)
    buf = (C.c_char * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    weakref.finalize(buf, _pinned_release, n, ptr)
    return arr


def timers() -> dict:
    t = Timers()
    lib().pawb200_get_timers(C.byref(t))
    return {n: getattr(t, n) for n, _ in Timers._fields_}


def reset_timers():
    lib().pawb200_reset_timers()
