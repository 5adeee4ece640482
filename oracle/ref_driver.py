"""ctypes driver for oracle/_ref/libpawpy_ref.so - TEST INFRASTRUCTURE ONLY.

libpawpy_ref.so is the UNMODIFIED reference C core (pawpyseed v0.7.1) built by
oracle/Makefile from /root/reference/pawpyseed/core.  This module mirrors the
structs of the reference's utils.h:30-166 so that tests can read
``bands[b]->projections[s].overlaps`` directly, and wraps the L3->L2 calls of
pawpyc.pyx (file:line cited per method).  Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libpawpy_ref.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class funcset_t(C.Structure):  # utils.h:30-47
    _fields_ = [("l", C.c_int)] + [(n, t) for n, t in [
        ("proj", c_dbl_p), ("proj_spline", C.POINTER(c_dbl_p)),
        ("aewave", c_dbl_p), ("aewave_spline", C.POINTER(c_dbl_p)),
        ("pswave", c_dbl_p), ("pswave_spline", C.POINTER(c_dbl_p)),
        ("diffwave", c_dbl_p), ("diffwave_spline", C.POINTER(c_dbl_p)),
        ("kwave", c_dbl_p), ("kwave_spline", C.POINTER(c_dbl_p)),
        ("smooth_diffwave", c_dbl_p), ("smooth_diffwave_spline", C.POINTER(c_dbl_p)),
        ("dense_kwave", c_dbl_p), ("dense_kwave_spline", C.POINTER(c_dbl_p))]]


class ppot_t(C.Structure):  # utils.h:49-70
    _fields_ = [("num_projs", C.c_int), ("total_projs", C.c_int), ("lmax", C.c_int),
                ("funcs", C.POINTER(funcset_t)), ("rmax", C.c_double),
                ("wave_rmax", C.c_double), ("pspw_overlap_matrix", c_dbl_p),
                ("aepw_overlap_matrix", c_dbl_p), ("diff_overlap_matrix", c_dbl_p),
                ("proj_gridsize", C.c_int), ("wave_gridsize", C.c_int),
                ("num_cart_gridpts", C.c_int), ("wave_grid", c_dbl_p),
                ("kwave_grid", c_dbl_p), ("proj_grid", c_dbl_p),
                ("smooth_grid", c_dbl_p), ("dense_kgrid", c_dbl_p)]


class projection_t(C.Structure):  # utils.h:72-79
    _fields_ = [("num_projs", C.c_int), ("total_projs", C.c_int), ("ns", c_int_p),
                ("ls", c_int_p), ("ms", c_int_p), ("overlaps", c_dbl_p)]


class band_t(C.Structure):  # utils.h:86-99
    _fields_ = [("n", C.c_int), ("num_waves", C.c_int), ("occ", C.c_double),
                ("N", C.c_double), ("energy", C.c_double),
                ("Cs", C.POINTER(C.c_float)), ("CRs", c_dbl_p),
                ("CAs", C.POINTER(C.c_float)),
                ("projections", C.POINTER(projection_t)),
                ("up_projections", C.POINTER(projection_t)),
                ("down_projections", C.POINTER(projection_t)),
                ("wave_projections", C.POINTER(projection_t))]


class kpoint_t(C.Structure):  # utils.h:106-115
    _fields_ = [("up", C.c_short), ("num_waves", C.c_int), ("Gs", c_int_p),
                ("k", c_dbl_p), ("weight", C.c_double), ("num_bands", C.c_int),
                ("bands", C.POINTER(C.POINTER(band_t))), ("expansion", C.c_void_p)]


class pswf_t(C.Structure):  # utils.h:117-141
    _fields_ = [("encut", C.c_double), ("num_elems", C.c_int), ("num_projs", c_int_p),
                ("num_sites", C.c_int), ("pps", C.POINTER(ppot_t)), ("G_bounds", c_int_p),
                ("kpts", C.POINTER(C.POINTER(kpoint_t))), ("nspin", C.c_int),
                ("nband", C.c_int), ("nwk", C.c_int), ("lattice", c_dbl_p),
                ("reclattice", c_dbl_p), ("fftg", c_int_p), ("is_ncl", C.c_int),
                ("wp_num", C.c_int), ("num_aug_overlap_sites", C.c_int),
                ("dcoords", c_dbl_p), ("overlaps", C.POINTER(c_dbl_p))]


class real_proj_t(C.Structure):  # utils.h:147-152
    _fields_ = [("l", C.c_int), ("m", C.c_int), ("func_num", C.c_int), ("values", c_dbl_p)]


class real_proj_site_t(C.Structure):  # utils.h:154-166
    _fields_ = [("index", C.c_int), ("elem", C.c_int), ("num_projs", C.c_int),
                ("total_projs", C.c_int), ("num_indices", C.c_int), ("gridsize", C.c_int),
                ("rmax", C.c_double), ("coord", c_dbl_p), ("indices", c_int_p),
                ("paths", c_dbl_p), ("projs", C.POINTER(real_proj_t))]


class cdouble(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def lib():
    """Load the reference library (import torch first: it carries the MKL DFTI symbols)."""
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (resolves libtorch_cpu / libc10 / libgomp)
        L = C.CDLL(REF_SO, mode=C.RTLD_LOCAL)
        P = C.POINTER(pswf_t)
        L.read_wavefunctions.restype = P
        L.read_wavefunctions.argtypes = [C.c_char_p, c_dbl_p]
        L.read_wavefunctions_from_str.restype = P
        L.read_wavefunctions_from_str.argtypes = [C.c_void_p, c_dbl_p]
        L.free_pswf.argtypes = [P]
        L.expand_symm_wf.restype = P
        L.expand_symm_wf.argtypes = [P, C.c_int, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p]
        L.get_projector_list.restype = C.POINTER(ppot_t)
        L.get_projector_list.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p,
                                         c_dbl_p, c_dbl_p, c_dbl_p, C.c_double]
        L.setup_projections.argtypes = [P, C.POINTER(ppot_t), C.c_int, C.c_int,
                                        c_int_p, c_int_p, c_dbl_p]
        L.pseudoprojection.argtypes = [c_dbl_p, P, P, C.c_int, C.c_int]
        L.overlap_setup_real.argtypes = [P, P, c_int_p, c_int_p, c_dbl_p, c_dbl_p,
                                         c_int_p, c_int_p, c_int_p, c_int_p,
                                         C.c_int, C.c_int, C.c_int]
        L.compensation_terms.argtypes = [c_dbl_p, C.c_int, P, P, C.c_int, C.c_int,
                                         C.c_int, C.c_int] + [c_int_p] * 6 + \
            [c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p, C.c_int]
        L.overlap_setup_recip.argtypes = L.overlap_setup_real.argtypes
        L.compensation_terms_recip.argtypes = L.compensation_terms.argtypes
        for name in ("realspace_state", "ncl_realspace_state"):
            getattr(L, name).argtypes = [c_dbl_p, C.c_int, C.c_int, P, c_int_p, c_int_p, c_dbl_p]
        L.remove_phase.argtypes = [c_dbl_p, C.c_int, P, c_int_p]
        L.ae_state_density.argtypes = [c_dbl_p, C.c_int, C.c_int, P, c_int_p, c_int_p, c_dbl_p]
        L.ae_chg_density.argtypes = [c_dbl_p, P, c_int_p, c_int_p, c_dbl_p]
        L.ncl_ae_chg_density.argtypes = [c_dbl_p, P, c_int_p, c_int_p, c_dbl_p]
        L.write_volumetric.argtypes = [C.c_char_p, c_dbl_p, c_int_p, C.c_double]
        L.project_realspace_state.argtypes = [c_dbl_p, C.c_int, P, P, c_int_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p]
        L.fft3d.argtypes = [c_dbl_p, c_int_p, c_dbl_p, c_dbl_p, c_int_p,
                            C.POINTER(C.c_float), C.c_int, c_int_p]
        L.fwd_fft3d.argtypes = L.fft3d.argtypes
        L.fft_check.restype = C.c_int
        L.fft_check.argtypes = [C.c_char_p, c_dbl_p, c_int_p]
        L.projector_values.restype = C.POINTER(real_proj_site_t)
        L.projector_values.argtypes = [C.c_int, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                       C.POINTER(ppot_t), c_int_p]
        L.smooth_pw_values.restype = C.POINTER(real_proj_site_t)
        L.smooth_pw_values.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                       C.POINTER(ppot_t), c_int_p]
        L.add_num_cart_gridpts.argtypes = [C.POINTER(ppot_t), c_dbl_p, c_int_p]
        L.free_real_proj_site_list.argtypes = [C.POINTER(real_proj_site_t), C.c_int]
        L.Ylm.restype = cdouble  # double complex comes back in xmm0:xmm1 like a 2-double struct
        L.Ylm.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.Ylm2.restype = cdouble
        L.Ylm2.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.legendre.restype = C.c_double
        L.legendre.argtypes = [C.c_int, C.c_int, C.c_double]
        L.get_occ.restype = C.c_double
        L.get_occ.argtypes = [P, C.c_int, C.c_int, C.c_int]
        L.get_energy.restype = C.c_double
        L.get_energy.argtypes = [P, C.c_int, C.c_int, C.c_int]
        L.spline_coeff.restype = C.POINTER(c_dbl_p)
        L.spline_coeff.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
        L.proj_interpolate.restype = C.c_double
        L.proj_interpolate.argtypes = [C.c_double, C.c_double, C.c_int, c_dbl_p, c_dbl_p,
                                       C.POINTER(c_dbl_p)]
        L.wave_interpolate.restype = C.c_double
        L.wave_interpolate.argtypes = [C.c_double, C.c_int, c_dbl_p, c_dbl_p, C.POINTER(c_dbl_p)]
        # momentum.h:56-92 (MomentumMatrix)
        L.momentum_grid_size.argtypes = [P, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, C.c_double]
        L.get_momentum_grid.restype = C.c_int
        L.get_momentum_grid.argtypes = [c_int_p, P, C.c_double, C.c_double, C.c_double, C.c_double]
        L.grid_bounds.argtypes = [c_int_p, c_int_p, c_int_p, C.c_int]
        L.list_to_grid_map.argtypes = [c_int_p, c_int_p, c_int_p, c_int_p, C.c_int]
        L.get_all_transforms.restype = C.c_void_p
        L.get_all_transforms.argtypes = [P, C.c_double]
        L.free_density_ft_elem_list.argtypes = [C.c_void_p, C.c_int]
        L.get_momentum_matrix.argtypes = [c_dbl_p, C.c_int, c_int_p, P, c_int_p, c_dbl_p] + [C.c_int] * 6 + \
            [C.c_void_p, C.c_double]
        L.fullwf_reciprocal.argtypes = [c_dbl_p, c_int_p, P, C.c_int, C.c_int, C.c_int, c_int_p, c_dbl_p]
        L.quick_overlap.restype = cdouble
        L.quick_overlap.argtypes = [c_int_p, c_dbl_p, c_dbl_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None and len(a) else None


def _dp(a):
    return a.ctypes.data_as(c_dbl_p)


class RefWavefunction:
    """Reference-side analogue of pawpyc.CWavefunction (pawpyc.pyx:328-555)."""

    def __init__(self, image_or_path, kws):
        L = lib()
        self.kws = np.ascontiguousarray(kws, dtype=np.float64)
        if isinstance(image_or_path, (str, bytes)):
            p = image_or_path.encode() if isinstance(image_or_path, str) else image_or_path
            self.ptr = L.read_wavefunctions(p, _dp(self.kws))      # pawpyc.pyx:224
        else:
            self._img = np.ascontiguousarray(image_or_path, dtype=np.uint8)
            self.ptr = L.read_wavefunctions_from_str(                # pawpyc.pyx:221
                self._img.ctypes.data_as(C.c_void_p), _dp(self.kws))
        self._describe()

    def _describe(self):
        w = self.ptr.contents
        self.nband, self.nwk, self.nspin, self.ncl = w.nband, w.nwk, w.nspin, bool(w.is_ncl)
        self.encut = w.encut
        self.lattice = np.array([w.lattice[i] for i in range(9)]).reshape(3, 3)
        self.projector_owner = False

    def expand_symm(self, maps, ops, drs, kws, trs):
        """pawpyc.pyx:227-292 (`PWFPointer.from_pointer_and_kpts`) -> expand_symm_wf (utils.c:829)."""
        maps = np.ascontiguousarray(maps, dtype=np.int32)
        ops = np.ascontiguousarray(ops, dtype=np.float64).reshape(-1)
        drs = np.ascontiguousarray(drs, dtype=np.float64).reshape(-1)
        trs = np.ascontiguousarray(trs, dtype=np.int32)
        new = object.__new__(RefWavefunction)
        new.kws = np.ascontiguousarray(kws, dtype=np.float64)
        new.ptr = lib().expand_symm_wf(self.ptr, len(maps), _ip(maps), _dp(ops), _dp(drs), _dp(new.kws), _ip(trs))
        # expand_symm_wf leaves encut unset on the new struct (utils.c:836-866); carry it over like pawpyc.pyx:290
        new.ptr.contents.encut = self.ptr.contents.encut
        new._describe()
        return new

    def free(self):
        if self.ptr:
            lib().free_pswf(self.ptr)
            self.ptr = None

    # ---- raw data accessors -------------------------------------------------
    def gvecs(self, kap):
        k = self.ptr.contents.kpts[kap].contents
        return np.ctypeslib.as_array(k.Gs, shape=(k.num_waves, 3)).copy()

    def kpt(self, kap):
        k = self.ptr.contents.kpts[kap].contents
        return np.array([k.k[0], k.k[1], k.k[2]])

    def coeffs(self, kap, b):
        k = self.ptr.contents.kpts[kap].contents
        bd = k.bands[b].contents
        return np.ctypeslib.as_array(bd.Cs, shape=(bd.num_waves * 2,)).copy().view(np.complex64)

    def _proj_list(self, kap, b, field, nsites):
        bd = self.ptr.contents.kpts[kap].contents.bands[b].contents
        arr = getattr(bd, field)
        out = []
        for s in range(nsites):
            pr = arr[s]
            n = pr.total_projs
            out.append(np.ctypeslib.as_array(pr.overlaps, shape=(2 * n,)).copy().view(np.complex128))
        return out

    def projections(self, kap, b, field="projections"):
        """concatenated <p_i|psi> over sites for band b at kappa (c128[nproj_tot])."""
        n = self.ptr.contents.wp_num if field == "wave_projections" else self.num_sites
        lst = self._proj_list(kap, b, field, n)
        return np.concatenate(lst) if lst else np.zeros(0, np.complex128)

    def channel_index(self, kap=0, b=0, field="projections"):
        """(site, n, l, m) int arrays in storage order - the bit-exact index contract."""
        bd = self.ptr.contents.kpts[kap].contents.bands[b].contents
        arr = getattr(bd, field)
        rows = []
        for s in range(self.num_sites):
            pr = arr[s]
            for p in range(pr.total_projs):
                rows.append((s, pr.ns[p], pr.ls[p], pr.ms[p]))
        return np.array(rows, dtype=np.int32).reshape(-1, 4)

    # ---- L3->L2 calls ---------------------------------------------------------
    def setup_projections(self, pps, labels, coords, dim, grid_encut):
        """pawpyc.pyx:352-412 (`_c_projector_setup`)."""
        from pawpyseed_b200.synth import flatten_pps
        L = lib()
        cl, ls, wg, pr, ae, ps, rm = flatten_pps(pps)
        self._keep = (cl, ls, wg, pr, ae, ps, rm)
        self.pps_ptr = L.get_projector_list(len(pps), _ip(cl), _ip(ls), _dp(wg), _dp(pr),
                                            _dp(ae), _dp(ps), _dp(rm), float(grid_encut))
        self.nums = np.ascontiguousarray(labels, dtype=np.int32)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
        self.dimv = np.ascontiguousarray(dim, dtype=np.int32)
        self.num_sites = len(self.nums)
        self.num_elems = len(pps)
        L.setup_projections(self.ptr, self.pps_ptr, len(pps), self.num_sites,
                            _ip(self.dimv), _ip(self.nums), _dp(self.coords))
        self.projector_owner = True

    def pseudoprojection(self, band_num, basis, flip_spin=False):
        """pawpyc.pyx:311-325."""
        res = np.zeros(basis.nband * basis.nwk * basis.nspin, dtype=np.complex128)
        lib().pseudoprojection(res.ctypes.data_as(c_dbl_p), basis.ptr, self.ptr,
                               int(band_num), int(flip_spin))
        return res

    def realspace_state(self, b, k, s, remove_phase=False):
        """pawpyc.pyx:425-439 / 563-580."""
        L = lib()
        n = int(np.prod(self.dimv))
        mult = 2 if self.ncl else 1
        res = np.zeros(n * mult, dtype=np.complex128)
        fn = L.ncl_realspace_state if self.ncl else L.realspace_state
        fn(res.ctypes.data_as(c_dbl_p), b, k + s * self.nwk, self.ptr, _ip(self.dimv),
           _ip(self.nums), _dp(self.coords))
        if remove_phase:
            for h in range(mult):
                L.remove_phase(res[h * n:].ctypes.data_as(c_dbl_p), k + s * self.nwk,
                               self.ptr, _ip(self.dimv))
        return res.reshape((mult,) + tuple(self.dimv)) if self.ncl else res.reshape(tuple(self.dimv))

    def chg_density(self, fdim=None):
        """pawpyc.pyx:455-461 / 582-588 (ae_chg_density on the fine grid 2*dim)."""
        L = lib()
        if self.ncl:
            fd = self.dimv.copy()
            fn = L.ncl_ae_chg_density
        else:
            fd = (self.dimv * 2).astype(np.int32) if fdim is None else np.asarray(fdim, np.int32)
            fn = L.ae_chg_density
        res = np.zeros(int(np.prod(fd)), dtype=np.float64)
        fn(_dp(res), self.ptr, _ip(fd), _ip(self.nums), _dp(self.coords))
        return res.reshape(tuple(fd))

    def time_projector_values(self):
        """Wall time of the reference's serial per-site table build (`projector_values` -> `setup_site`,
        projector.c:193-208, utils.c:590-696) for the sites given to setup_projections; the tables are freed."""
        import time
        L = lib()
        w = self.ptr.contents
        t0 = time.perf_counter()
        sp = L.projector_values(self.num_sites, _ip(self.nums), _dp(self.coords), w.lattice, w.reclattice,
                                self.pps_ptr, _ip(self.dimv))
        dt = time.perf_counter() - t0
        L.free_real_proj_site_list(sp, self.num_sites)
        return dt

    def site_tables(self, which="proj", site_list=None, indices_only=False):
        """projector_values / smooth_pw_values (projector.c:193-221) -> python lists."""
        L = lib()
        w = self.ptr.contents
        for e in range(self.num_elems):
            L.add_num_cart_gridpts(C.byref(self.pps_ptr[e]), w.lattice, _ip(self.dimv))
        if which == "proj":
            ns = self.num_sites
            sp = L.projector_values(ns, _ip(self.nums), _dp(self.coords), w.lattice,
                                    w.reclattice, self.pps_ptr, _ip(self.dimv))
        else:
            sl = np.ascontiguousarray(site_list, dtype=np.int32)
            ns = len(sl)
            sp = L.smooth_pw_values(ns, _ip(sl), _ip(self.nums), _dp(self.coords), w.lattice,
                                    w.reclattice, self.pps_ptr, _ip(self.dimv))
        out = []
        for s in range(ns):
            st = sp[s]
            n = st.num_indices
            idx = np.ctypeslib.as_array(st.indices, shape=(n,)).copy()
            if indices_only:
                out.append(dict(indices=idx))
                continue
            paths = np.ctypeslib.as_array(st.paths, shape=(n, 3)).copy()
            vals = np.stack([np.ctypeslib.as_array(st.projs[p].values, shape=(2 * n,)).copy()
                             .view(np.complex128) for p in range(st.total_projs)]) \
                if st.total_projs else np.zeros((0, n), np.complex128)
            out.append(dict(indices=idx, paths=paths, values=vals))
        L.free_real_proj_site_list(sp, ns)
        return out


class RefProjector:
    """Reference-side analogue of pawpyc.CProjector (pawpyc.pyx:634-702)."""

    def __init__(self, wf: RefWavefunction, basis: RefWavefunction, site_cat, recip=False):
        self.wf, self.basis = wf, basis
        self.cat = [np.ascontiguousarray(x, dtype=np.int32) for x in site_cat]
        M_R, M_S, N_R, N_S, N_RS_R, N_RS_S = self.cat
        self.recip = recip
        setup = lib().overlap_setup_recip if recip else lib().overlap_setup_real
        setup(basis.ptr, wf.ptr, _ip(basis.nums), _ip(wf.nums),
                                 _dp(basis.coords), _dp(wf.coords), _ip(N_R), _ip(N_S),
                                 _ip(N_RS_R), _ip(N_RS_S), len(N_R), len(N_S), len(N_RS_R))

    def add_augmentation_terms(self, res, band_num, flip_spin=False):
        M_R, M_S, N_R, N_S, N_RS_R, N_RS_S = self.cat
        fn = lib().compensation_terms_recip if self.recip else lib().compensation_terms
        fn(res.ctypes.data_as(c_dbl_p), int(band_num), self.wf.ptr,
                                 self.basis.ptr, len(M_R), len(N_R), len(N_S), len(N_RS_R),
                                 _ip(M_R), _ip(M_S), _ip(N_R), _ip(N_S), _ip(N_RS_R), _ip(N_RS_S),
                                 _ip(self.wf.nums), _dp(self.wf.coords), _ip(self.basis.nums),
                                 _dp(self.basis.coords), _ip(self.wf.dimv), int(flip_spin))
        return res

    def realspace_projection(self, band_num, dim):
        """pawpyc.pyx:723-736 -> project_realspace_state (density.c:205-230)."""
        res = np.zeros(self.basis.nband * self.basis.nwk * self.basis.nspin, dtype=np.complex128)
        dimv = np.ascontiguousarray(dim, dtype=np.int32)
        lib().project_realspace_state(res.ctypes.data_as(c_dbl_p), int(band_num), self.wf.ptr, self.basis.ptr,
                                      _ip(dimv), _ip(self.wf.nums), _dp(self.wf.coords), _ip(self.basis.nums),
                                      _dp(self.basis.coords))
        return res

    def single_band_projection(self, band_num, flip_spin=False):
        """projector.py:210-223 (`aug_real`)."""
        res = self.wf.pseudoprojection(band_num, self.basis, flip_spin)
        return self.add_augmentation_terms(res, band_num, flip_spin)


class RefMomentumMatrix:
    """Reference-side analogue of pawpyc.CMomentumMatrix (pawpyc.pyx:738-807)."""

    def __init__(self, wf: RefWavefunction, encut):
        L = lib()
        self.wf, self.encut = wf, float(encut)
        nb = [C.c_double(0) for _ in range(3)]
        npmax = C.c_int(0)
        L.momentum_grid_size(wf.ptr, C.byref(nb[0]), C.byref(nb[1]), C.byref(nb[2]), C.byref(npmax), self.encut)
        self.nbmax = [v.value for v in nb]
        grid = np.zeros(3 * npmax.value, dtype=np.int32)
        n = L.get_momentum_grid(_ip(grid), wf.ptr, nb[0].value, nb[1].value, nb[2].value, self.encut)
        self.ggrid = np.ascontiguousarray(grid[:3 * n])
        self.gbounds = np.zeros(6, dtype=np.int32)
        self.gdim = np.zeros(3, dtype=np.int32)
        L.grid_bounds(_ip(self.gbounds), _ip(self.gdim), _ip(self.ggrid), n)
        self.grid3d = -np.ones(int(np.prod(self.gdim)), dtype=np.int32)
        L.list_to_grid_map(_ip(self.grid3d), _ip(self.gbounds), _ip(self.gdim), _ip(self.ggrid), n)
        self.transforms = L.get_all_transforms(wf.ptr, self.encut)

    def momentum_matrix_elems(self, b1, k1, s1, b2, k2, s2):
        numg = len(self.ggrid) // 3
        res = np.zeros(numg, dtype=np.complex128)
        lib().get_momentum_matrix(res.ctypes.data_as(c_dbl_p), numg, _ip(self.ggrid), self.wf.ptr, _ip(self.wf.nums),
                                  _dp(self.wf.coords), b1, k1, s1, b2, k2, s2, self.transforms, self.encut)
        return res

    def reciprocal_fullfw(self, b, k, s):
        numg = len(self.ggrid) // 3
        res = np.zeros(numg, dtype=np.complex128)
        lib().fullwf_reciprocal(res.ctypes.data_as(c_dbl_p), _ip(self.ggrid), self.wf.ptr, numg, b,
                                k + s * self.wf.nwk, _ip(self.wf.nums), _dp(self.wf.coords))
        return res

    def g_from_fullfw(self, b1, k1, s1, b2, k2, s2, G):
        v1, v2 = self.reciprocal_fullfw(b1, k1, s1), self.reciprocal_fullfw(b2, k2, s2)
        GP = np.ascontiguousarray(G, dtype=np.int32)
        r = lib().quick_overlap(_ip(GP), v1.ctypes.data_as(c_dbl_p), v2.ctypes.data_as(c_dbl_p), len(self.ggrid) // 3,
                                _ip(self.ggrid), _ip(self.grid3d), _ip(self.gbounds), _ip(self.gdim))
        return complex(r.re, r.im)
