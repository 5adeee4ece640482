"""Host-side mirror of pawpyseed's Cython shim (pawpyseed/core/pawpyc.pyx).

Same class names, method names, argument meaning and error behaviour as the reference
cdef classes - ``PWFPointer`` (:197), ``PseudoWavefunction`` (:270), ``CWavefunction`` (:328),
``CNCLWavefunction`` (:558), ``CProjector`` (:634), ``Timer`` (:37) - but every C call goes
to the GPU engine through the C ABI of include/pawpyseed_b200.h.  pymatgen / monty are
optional: a ``Vasprun``-like object only needs ``actual_kpoints``,
``actual_kpoints_weights`` and ``eigenvalue_band_properties``.
"""
from __future__ import annotations

import ctypes as C
import gzip
import bz2
import sys
import time

import numpy as np

from . import _lib
from ._lib import PAWpyError, check, dp, f64, i32, ip


class Timer:
    """pawpyc.pyx:37-49."""
    ALL_SETUP_TIME = 0
    ALL_OVERLAP_TIME = 0
    ALL_AUGMENTATION_TIME = 0

    @staticmethod
    def setup_time(t):
        Timer.ALL_SETUP_TIME += t

    @staticmethod
    def overlap_time(t):
        Timer.ALL_OVERLAP_TIME += t

    @staticmethod
    def augmentation_time(t):
        Timer.ALL_AUGMENTATION_TIME += t


def el(site):
    """pawpyc.pyx:55-60."""
    return site.specie.symbol


# ---- C utility wrappers (pawpyc.pyx:77-191) ------------------------------------------------
def legendre(l, m, x):
    return _lib.lib().pawb200_legendre(int(l), int(m), float(x))


def Ylm(l, m, theta, phi):
    o = np.zeros(2)
    _lib.lib().pawb200_Ylm(int(l), int(m), float(theta), float(phi), dp(o))
    return complex(o[0], o[1])


def Ylm2(l, m, costheta, phi):
    o = np.zeros(2)
    _lib.lib().pawb200_Ylm2(int(l), int(m), float(costheta), float(phi), dp(o))
    return complex(o[0], o[1])


def frac_to_cartesian(coord, lattice):
    lat = f64(np.asarray(lattice).flatten())
    _lib.lib().pawb200_frac_to_cartesian(dp(coord), dp(lat))


def cartesian_to_frac(coord, reclattice):
    rec = f64(np.asarray(reclattice).flatten())
    _lib.lib().pawb200_cartesian_to_frac(dp(coord), dp(rec))


def _spline(x, y):
    L = _lib.lib()
    n = len(x)
    p = L.pawb200_spline_coeff(dp(x), dp(y), n)
    out = np.ctypeslib.as_array(C.cast(p, _lib.c_dbl_p), shape=(3 * n,)).copy()
    L.pawb200_free_ptr(p)
    return out


def interpolate(res, tst, x, y, rmax, size, tstsize):
    L = _lib.lib()
    x, y = f64(x), f64(y)
    coef = _spline(x[:size], y[:size])
    for i in range(tstsize):
        res[i] = L.pawb200_proj_interpolate(float(tst[i]), float(rmax), int(size), dp(x), dp(y), dp(coef))


def spherical_bessel_transform(encut, l, r, f):
    r, f = f64(r), f64(f)
    k = np.zeros(len(r))
    fk = np.zeros(len(r))
    _lib.lib().pawb200_spherical_bessel_transform(float(encut), int(l), len(r), dp(r), dp(f), dp(k), dp(fk))
    check()
    return k, fk


def reciprocal_offsite_wave_overlap(dcoord, r1, f1, r2, f2, l1, m1, l2, m2):
    encut = 1e5
    k1, fk1 = spherical_bessel_transform(encut, l1, r1, f1)
    k2, fk2 = spherical_bessel_transform(encut, l2, r2, f2)
    s1, s2 = _spline(k1, fk1), _spline(k2, fk2)
    d = f64(dcoord)
    o = np.zeros(2)
    _lib.lib().pawb200_reciprocal_offsite_wave_overlap(dp(d), dp(k1), dp(fk1), dp(s1), len(k1), dp(k2),
                                                       dp(fk2), dp(s2), len(k2), int(l1), int(m1),
                                                       int(l2), int(m2), dp(o))
    check()
    return complex(o[0], o[1])


# ---- base classes ------------------------------------------------------------------------------
class PWFPointer:
    """pawpyc.pyx:197-267.  Holds the opaque engine handle (GPU-resident coefficients)."""

    def __init__(self, filename=None, vr=None):
        self.ptr = None
        self._buf = None
        if filename is None or vr is None:
            return
        self.weights = np.array(vr.actual_kpoints_weights, dtype=np.float64)
        self.kpts = np.array(vr.actual_kpoints, dtype=np.float64)
        self.band_props = np.array(vr.eigenvalue_band_properties)
        self._read(filename)

    @classmethod
    def from_arrays(cls, source, kpts, weights, band_props=(0.0, 0.0, 0.0, False)):
        """Same as the constructor without a Vasprun: `source` is a WAVECAR path (.gz/.bz2
        accepted, pawpyc.pyx:217-222) or an in-memory WAVECAR image (bytes / uint8 array)."""
        self = cls()
        self.weights = np.array(weights, dtype=np.float64)
        self.kpts = np.array(kpts, dtype=np.float64).reshape(-1, 3)
        self.band_props = np.array(band_props)
        self._read(source)
        return self

    @staticmethod
    def from_pointer_and_kpts(ptr, structure, kpts, band_props, allkpts, weights, symprec,
                              time_reversal_symmetry, symmops=None):
        """pawpyc.pyx:227-267: desymmetrised copy of the wavefunction behind `ptr` (the source is left
        untouched).  `symmops` (reciprocal-fractional operators) skips the pymatgen space-group search."""
        from .symmetry import get_kpt_mapping, get_nosym_kpoints, make_c_ops
        if allkpts is None or weights is None:
            allkpts, orig_kptnums, op_nums, symmops, trs = get_nosym_kpoints(
                kpts, structure, symprec=symprec, fil_trsym=time_reversal_symmetry, symmops=symmops)
            weights = np.ones(allkpts.shape[0], dtype=np.float64)
            weights[np.linalg.norm(allkpts, axis=1) < 1e-10] *= 0.5
            weights /= np.sum(weights)
        else:
            orig_kptnums, op_nums, symmops, trs = get_kpt_mapping(allkpts, kpts, structure, symprec=symprec,
                                                                  symmops=symmops)
        ops, drs = make_c_ops(op_nums, symmops)
        pwfp = PWFPointer()
        pwfp.kpts = np.ascontiguousarray(allkpts, dtype=np.float64).reshape(-1, 3)
        pwfp.weights = f64(weights)
        pwfp.band_props = np.array(band_props)
        maps, trs = i32(orig_kptnums), i32(trs)
        pwfp.ptr = _lib.lib().pawb200_expand_symm_wf(ptr, len(maps), ip(maps), dp(ops), dp(drs), dp(pwfp.weights),
                                                    ip(trs))
        check()
        if not pwfp.ptr:
            raise PAWpyError("expand_symm_wf returned NULL")
        return pwfp

    def _read(self, source):
        L = _lib.lib()
        kws = f64(self.weights)
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else source
            self._buf = np.ascontiguousarray(buf, dtype=np.uint8)
            self.ptr = L.pawb200_read_wavefunctions_from_str(self._buf.ctypes.data_as(C.c_void_p), dp(kws))
            # kept alive with the wavefunction: with pawb200_set_async_ingest(1) the copy may still be in flight
        else:
            filename = str(source)
            if ".gz" in filename or ".bz2" in filename:
                opener = gzip.open if ".gz" in filename else bz2.open
                with opener(filename, "rb") as f:
                    self._buf = np.frombuffer(f.read(), dtype=np.uint8)
                self.ptr = L.pawb200_read_wavefunctions_from_str(self._buf.ctypes.data_as(C.c_void_p), dp(kws))
            else:
                self.ptr = L.pawb200_read_wavefunctions(filename.encode("utf-8"), dp(kws))
        check()
        if not self.ptr:
            raise PAWpyError("read_wavefunctions returned NULL")
        sys.stdout.flush()


class PseudoWavefunction:
    """pawpyc.pyx:270-325."""

    def __init__(self, pwf: PWFPointer):
        if pwf.ptr is None:
            raise Exception("NULL PWFPointer ptr!")
        L = _lib.lib()
        self.wf_ptr = pwf.ptr
        self._src_buf = getattr(pwf, "_buf", None)
        pwf.ptr = None   # ownership moves, like the C pointer in the reference
        self.kpts = pwf.kpts.copy(order="C")
        self.kws = pwf.weights.copy(order="C")
        self.ncl = L.pawb200_is_ncl(self.wf_ptr) > 0
        self.nband = L.pawb200_get_nband(self.wf_ptr)
        self.nwk = L.pawb200_get_nwk(self.wf_ptr)
        self.nspin = L.pawb200_get_nspin(self.wf_ptr)
        self.encut = L.pawb200_get_encut(self.wf_ptr)

    def __del__(self):
        try:
            if getattr(self, "wf_ptr", None):
                _lib.lib().pawb200_free_pswf(self.wf_ptr)
                self.wf_ptr = None
        except Exception:
            pass

    def _desymmetrized_pwf(self, structure, band_props, allkpts=None, weights=None, symprec=1e-4,
                           time_reversal_symmetry=True, symmops=None):
        """pawpyc.pyx:317-325."""
        return PWFPointer.from_pointer_and_kpts(self.wf_ptr, structure, self.kpts, band_props, allkpts, weights,
                                                symprec, time_reversal_symmetry, symmops=symmops)

    def _get_coefficients(self, b, kappa):
        """Plane-wave coefficients of (band, kappa) in WAVECAR order (test accessor, complex64)."""
        L = _lib.lib()
        k3 = np.zeros(3)
        n = L.pawb200_get_kpoint(self.wf_ptr, int(kappa), dp(k3), None)
        if n < 0:
            raise ValueError("Invalid kpoint index %d" % kappa)
        out = np.zeros(n, dtype=np.complex64)
        L.pawb200_get_coefficients(self.wf_ptr, int(b), int(kappa), out.ctypes.data_as(C.c_void_p))
        check()
        return out

    def pseudoprojection(self, band_num, basis, flip_spin=False):
        """<psibt_n1k|psit_n2k> for all n1, k and a given n2 (pawpyc.pyx:311-325)."""
        res = np.zeros(basis.nband * basis.nwk * basis.nspin, dtype=np.complex128)
        _lib.lib().pawb200_pseudoprojection(res.ctypes.data_as(_lib.c_dbl_p), basis.wf_ptr, self.wf_ptr,
                                            int(band_num), int(bool(flip_spin)))
        check()
        return res


class CWavefunction(PseudoWavefunction):
    """pawpyc.pyx:328-555."""

    def __init__(self, pwf):
        self.projector_owner = 0
        super().__init__(pwf)

    def _c_projector_setup(self, num_elems, num_sites, grid_encut, nums, coords, dim, pps):
        """pawpyc.pyx:352-412: flatten the PAW data per element (sorted labels) and run
        get_projector_list + setup_projections."""
        L = _lib.lib()
        start = time.monotonic()
        clabels, ls, wgrids, projectors, aewaves, pswaves, rmaxs = [], [], [], [], [], [], []
        for num in sorted(pps.keys()):
            pp = pps[num]
            clabels += [num, len(pp.ls), pp.ndata, len(pp.grid)]
            rmaxs.append(pp.rmax)
            ls += list(pp.ls)
            wgrids.append(np.asarray(pp.grid, dtype=np.float64))
            for i in range(len(pp.ls)):
                projectors.append(np.asarray(pp.realprojs[i], dtype=np.float64))
                aewaves.append(np.asarray(pp.aewaves[i], dtype=np.float64))
                pswaves.append(np.asarray(pp.pswaves[i], dtype=np.float64))
        clabels_v, ls_v = i32(clabels), i32(ls)
        wgrids_v, projectors_v = f64(np.concatenate(wgrids)), f64(np.concatenate(projectors))
        aewaves_v, pswaves_v = f64(np.concatenate(aewaves)), f64(np.concatenate(pswaves))
        rmaxs_v = f64(rmaxs)
        projector_list = L.pawb200_get_projector_list(
            int(num_elems), ip(clabels_v), ip(ls_v), dp(wgrids_v), dp(projectors_v), dp(aewaves_v),
            dp(pswaves_v), dp(rmaxs_v), float(grid_encut))
        check()
        end = time.monotonic()
        Timer.setup_time(end - start)
        self.number_projector_elements = num_elems
        self.nums = np.array(nums, dtype=np.int32, copy=True)
        self.coords = np.array(coords, dtype=np.float64, copy=True).reshape(-1)
        self.update_dimv(dim)
        L.pawb200_setup_projections(self.wf_ptr, projector_list, int(num_elems), int(num_sites),
                                    ip(self.dimv), ip(self.nums), dp(self.coords))
        check()
        self.projector_owner = 1

    def update_dimv(self, dim):
        dim = np.array(dim, dtype=np.int32, order="C")
        self.dimv = dim
        self.fdimv = (dim * 2).astype(np.int32)
        self.gridsize = int(np.cumprod(dim)[-1])
        self.fgridsize = int(np.cumprod(dim * 2)[-1])

    def _check_bks(self, b, k, s):
        if b < 0 or b >= self.nband:
            raise ValueError("Invalid band choice")
        if k < 0 or k >= self.nwk:
            raise ValueError("Invalid k-point choice")
        if s < 0 or s >= self.nspin:
            raise ValueError("Invalid spin choice")

    def _get_realspace_state(self, b, k, s, remove_phase=False):
        self._check_bks(b, k, s)
        L = _lib.lib()
        res = np.zeros(self.gridsize, dtype=np.complex128, order="C")
        L.pawb200_realspace_state(res.ctypes.data_as(_lib.c_dbl_p), b, k + s * self.nwk, self.wf_ptr,
                                  ip(self.dimv), ip(self.nums), dp(self.coords))
        check()
        if remove_phase:
            L.pawb200_remove_phase(res.ctypes.data_as(_lib.c_dbl_p), k + s * self.nwk, self.wf_ptr, ip(self.dimv))
            check()
        res.shape = tuple(self.dimv)
        return res

    def _get_realspace_state_density(self, b, k, s):
        self._check_bks(b, k, s)
        res = np.zeros(self.fgridsize, dtype=np.float64, order="C")
        _lib.lib().pawb200_ae_state_density(dp(res), b, k + s * self.nwk, self.wf_ptr, ip(self.fdimv),
                                            ip(self.nums), dp(self.coords))
        check()
        res.shape = tuple(self.fdimv)
        return res

    def _get_realspace_density(self, bands=None):
        """pawpyc.pyx:455-494."""
        L = _lib.lib()
        res = np.zeros(self.fgridsize, dtype=np.float64, order="C")
        if bands is None:
            L.pawb200_ae_chg_density(dp(res), self.wf_ptr, ip(self.fdimv), ip(self.nums), dp(self.coords))
            check()
        else:
            blist = [bands] if isinstance(bands, (int, np.integer)) else list(bands)
            for b in blist:
                if b < 0 or b >= self.nband:
                    raise ValueError("Invalid band choice")
                for k in range(self.nwk * self.nspin):
                    work = np.zeros(self.fgridsize, dtype=np.float64, order="C")
                    L.pawb200_ae_state_density(dp(work), int(b), k, self.wf_ptr, ip(self.fdimv),
                                               ip(self.nums), dp(self.coords))
                    check()
                    res += work * self.kws[k % self.nwk] / self.nspin
        res.shape = tuple(self.fdimv)
        return res

    def _get_realspace_density_shard(self, band_lo, band_hi):
        """Extension: ae_chg_density restricted to the occupied bands in [band_lo, band_hi) (a rank's band shard)."""
        res = np.zeros(self.fgridsize, dtype=np.float64, order="C")
        _lib.lib().pawb200_ae_chg_density_bands(dp(res), self.wf_ptr, ip(self.fdimv), ip(self.nums), dp(self.coords),
                                                int(band_lo), int(band_hi))
        check()
        res.shape = tuple(self.fdimv)
        return res

    def _write_realspace_state(self, filename1, filename2, scale, b, k, s, remove_phase=False):
        self._check_bks(b, k, s)
        L = _lib.lib()
        res = self._get_realspace_state(b, k, s, remove_phase)
        flat = res.reshape(self.gridsize)
        for fn, part in ((filename1, np.real(flat)), (filename2, np.imag(flat))):
            arr = np.ascontiguousarray(part)
            L.pawb200_write_volumetric(fn.encode("utf-8"), dp(arr), ip(self.dimv), float(scale))
            check()
        return res

    def _write_realspace_density(self, filename, scale, bands=None):
        res = self._get_realspace_density(bands)
        flat = np.ascontiguousarray(res.reshape(self.fgridsize))
        _lib.lib().pawb200_write_volumetric(filename.encode("utf-8"), dp(flat), ip(self.fdimv), float(scale))
        check()
        return res

    def _get_occs(self):
        """pawpyc.pyx:532-538 (through get_occ instead of reaching into the struct)."""
        L = _lib.lib()
        nk = self.nwk * self.nspin
        p = L.pawb200_get_occs(self.wf_ptr)
        res = np.ctypeslib.as_array(C.cast(p, _lib.c_dbl_p), shape=(self.nband * nk,)).copy()
        L.pawb200_free_ptr(p)
        return res

    def _get_energy_list(self, bands):
        L = _lib.lib()
        for b in bands:
            if b < 0 or b >= self.nband:
                raise ValueError("Invalid band choice")
        energy_list = {}
        for b in bands:
            energy_list[b] = []
            for s in range(self.nspin):
                for k in range(self.nwk):
                    energy_list[b].append([L.pawb200_get_energy(self.wf_ptr, b, k, s),
                                           L.pawb200_get_occ(self.wf_ptr, b, k, s)])
        return energy_list

    # ---- extensions: direct access to device-resident projections ------------------------------
    def _get_projections(self, b, kappa, which=0):
        L = _lib.lib()
        n = L.pawb200_num_projections(self.wf_ptr, which)
        out = np.zeros(n, dtype=np.complex128)
        L.pawb200_get_projections(self.wf_ptr, int(b), int(kappa), int(which), out.ctypes.data_as(_lib.c_dbl_p))
        check()
        return out

    def _get_channel_index(self):
        L = _lib.lib()
        n = L.pawb200_get_channel_index(self.wf_ptr, None)
        out = np.zeros((n, 4), dtype=np.int32)
        L.pawb200_get_channel_index(self.wf_ptr, ip(out.reshape(-1)))
        return out

    def _get_site_indices(self, site):
        L = _lib.lib()
        n = L.pawb200_get_site_indices(self.wf_ptr, int(site), None, 0)
        if n < 0:
            raise PAWpyError("site tables not available")
        out = np.zeros(max(n, 1), dtype=np.int32)
        L.pawb200_get_site_indices(self.wf_ptr, int(site), ip(out), n)
        return out[:n]


class CNCLWavefunction(CWavefunction):
    """pawpyc.pyx:558-631."""

    def _get_realspace_state(self, b, k, s, remove_phase=False):
        self._check_bks(b, k, s)
        L = _lib.lib()
        res = np.zeros(self.gridsize * 2, dtype=np.complex128, order="C")
        L.pawb200_ncl_realspace_state(res.ctypes.data_as(_lib.c_dbl_p), b, k + s * self.nwk, self.wf_ptr,
                                      ip(self.dimv), ip(self.nums), dp(self.coords))
        check()
        res0, res1 = res[:self.gridsize], res[self.gridsize:]
        if remove_phase:
            for part in (res0, res1):
                L.pawb200_remove_phase(part.ctypes.data_as(_lib.c_dbl_p), k + s * self.nwk, self.wf_ptr,
                                       ip(self.dimv))
                check()
        return res0.reshape(tuple(self.dimv)), res1.reshape(tuple(self.dimv))

    def _get_realspace_density(self):
        res = np.zeros(self.gridsize, dtype=np.float64, order="C")
        _lib.lib().pawb200_ncl_ae_chg_density(dp(res), self.wf_ptr, ip(self.dimv), ip(self.nums), dp(self.coords))
        check()
        res.shape = tuple(self.dimv)
        return res

    def _write_realspace_state(self, filename1, filename2, filename3, filename4, scale, b, k, s,
                               remove_phase=False):
        self._check_bks(b, k, s)
        L = _lib.lib()
        res0, res1 = self._get_realspace_state(b, k, s, remove_phase=remove_phase)
        for fr, fi, res in ((filename1, filename2, res0), (filename3, filename4, res1)):
            flat = res.reshape(self.gridsize)
            for fn, part in ((fr, np.real(flat)), (fi, np.imag(flat))):
                arr = np.ascontiguousarray(part)
                L.pawb200_write_volumetric(fn.encode("utf-8"), dp(arr), ip(self.dimv), float(scale))
                check()
        return res0, res1

    def _write_realspace_density(self, filename, scale):
        res = self._get_realspace_density()
        flat = np.ascontiguousarray(res.reshape(self.gridsize))
        _lib.lib().pawb200_write_volumetric(filename.encode("utf-8"), dp(flat), ip(self.dimv), float(scale))
        check()
        return res


class CProjector:
    """pawpyc.pyx:634-736."""

    def __init__(self, wf, basis):
        self.wf = wf
        self.basis = basis

    def _setup_overlap(self, site_cat, recip):
        """pawpyc.pyx:642-683."""
        names = ("M_R", "M_S", "N_R", "N_S", "N_RS_R", "N_RS_S")
        for n, lst in zip(names, site_cat):
            setattr(self, n, np.array(lst, dtype=np.int32, order="C"))
            setattr(self, "num_" + n, len(lst))
        self._recip = bool(recip)
        fn = _lib.lib().pawb200_overlap_setup_recip if recip else _lib.lib().pawb200_overlap_setup_real
        fn(
            self.basis.wf_ptr, self.wf.wf_ptr, ip(self.basis.nums), ip(self.wf.nums),
            dp(self.basis.coords), dp(self.wf.coords), ip(self.N_R), ip(self.N_S), ip(self.N_RS_R),
            ip(self.N_RS_S), self.num_N_R, self.num_N_S, self.num_N_RS_R)
        check()

    def _add_augmentation_terms(self, res, band_num, flip_spin):
        if res.dtype != np.complex128 or not res.flags["C_CONTIGUOUS"]:
            raise ValueError("res must be a contiguous complex128 array")
        _lib.lib().pawb200_compensation_terms(
            res.ctypes.data_as(_lib.c_dbl_p), int(band_num), self.wf.wf_ptr, self.basis.wf_ptr,
            self.num_M_R, self.num_N_R, self.num_N_S, self.num_N_RS_R, ip(self.M_R), ip(self.M_S),
            ip(self.N_R), ip(self.N_S), ip(self.N_RS_R), ip(self.N_RS_S), ip(self.wf.nums),
            dp(self.wf.coords), ip(self.basis.nums), dp(self.basis.coords), ip(self.wf.dimv),
            int(bool(flip_spin)))
        check()

    def _projection_recip(self, res, band_num, flip_spin):
        """pawpyc.pyx:704-721."""
        if res.dtype != np.complex128 or not res.flags["C_CONTIGUOUS"]:
            raise ValueError("res must be a contiguous complex128 array")
        _lib.lib().pawb200_compensation_terms_recip(
            res.ctypes.data_as(_lib.c_dbl_p), int(band_num), self.wf.wf_ptr, self.basis.wf_ptr,
            self.num_M_R, self.num_N_R, self.num_N_S, self.num_N_RS_R, ip(self.M_R), ip(self.M_S),
            ip(self.N_R), ip(self.N_S), ip(self.N_RS_R), ip(self.N_RS_S), ip(self.wf.nums),
            dp(self.wf.coords), ip(self.basis.nums), dp(self.basis.coords), ip(self.wf.dimv),
            int(bool(flip_spin)))
        check()

    def _realspace_projection(self, band_num, dim):
        """pawpyc.pyx:723-736 -> project_realspace_state (density.c:205-230)."""
        res = np.zeros(self.basis.nband * self.basis.nwk * self.basis.nspin, dtype=np.complex128, order="C")
        dimv = self.wf.dimv if dim is None else np.array(dim, dtype=np.int32, order="C")
        _lib.lib().pawb200_project_realspace_state(
            res.ctypes.data_as(_lib.c_dbl_p), int(band_num), self.wf.wf_ptr, self.basis.wf_ptr, ip(dimv),
            ip(self.wf.nums), dp(self.wf.coords), ip(self.basis.nums), dp(self.basis.coords))
        check()
        return res

    # ---- extension: all band pairs in one call ---------------------------------------------------
    def _projection_matrix(self, flip_spin=False, kappa_range=None, pseudo_only=False):
        """out[kappa, b_wf, b_basis]; row b_wf of block kappa equals
        single_band_projection(b_wf)[b_basis*NK + kappa]."""
        NK = self.basis.nwk * self.basis.nspin
        lo, hi = (0, NK) if kappa_range is None else kappa_range
        out = _lib.pinned_empty((hi - lo, self.wf.nband, self.basis.nband), np.complex128)   # fully overwritten
        have = hasattr(self, "M_R")
        z = np.zeros(0, np.int32)
        lists = [getattr(self, n) if have else z for n in ("M_R", "M_S", "N_R", "N_S", "N_RS_R", "N_RS_S")]
        _lib.lib().pawb200_projection_matrix(
            out.ctypes.data_as(_lib.c_dbl_p), self.wf.wf_ptr, self.basis.wf_ptr, len(lists[0]),
            len(lists[2]), len(lists[3]), len(lists[4]), *[ip(a) for a in lists],
            int(bool(flip_spin)), int(lo), int(hi),
            1 if (pseudo_only or not have) else (2 if getattr(self, "_recip", False) else 0))
        check()
        return out


    def _projection_matrix_dev(self, out_dev, flip_spin=False, kappa_range=None, pseudo_only=False):
        """Same blocks, written to the DEVICE tensor `out_dev` (torch complex128 / float64 view, contiguous,
        [(hi - lo), nband_wf, nband_basis]).  Nothing is copied to the host and nothing is synchronised: the work is
        queued on the legacy default stream, which is torch's default stream."""
        NK = self.basis.nwk * self.basis.nspin
        lo, hi = (0, NK) if kappa_range is None else kappa_range
        need = (hi - lo) * self.wf.nband * self.basis.nband * 16
        if out_dev.numel() * out_dev.element_size() < need or not out_dev.is_cuda or not out_dev.is_contiguous():
            raise ValueError("out_dev must be a contiguous CUDA tensor of at least %d bytes" % need)
        have = hasattr(self, "M_R")
        z = np.zeros(0, np.int32)
        lists = [getattr(self, n) if have else z for n in ("M_R", "M_S", "N_R", "N_S", "N_RS_R", "N_RS_S")]
        _lib.lib().pawb200_projection_matrix_dev(
            C.c_void_p(out_dev.data_ptr()), self.wf.wf_ptr, self.basis.wf_ptr, len(lists[0]),
            len(lists[2]), len(lists[3]), len(lists[4]), *[ip(a) for a in lists],
            int(bool(flip_spin)), int(lo), int(hi),
            1 if (pseudo_only or not have) else (2 if getattr(self, "_recip", False) else 0))
        check()
        return out_dev


class CMomentumMatrix:
    """pawpyc.pyx:738-807."""

    def __init__(self, wf, encut):
        self.wf = wf
        self.momentum_encut = float(encut)
        self.elem_density_transforms = None
        self._setup_momentum_grid()
        self._setup_transforms()

    def __del__(self):
        try:
            if getattr(self, "elem_density_transforms", None):
                _lib.lib().pawb200_free_density_ft_elem_list(self.elem_density_transforms, 0)
                self.elem_density_transforms = None
        except Exception:
            pass

    def _setup_momentum_grid(self):
        L = _lib.lib()
        nb = [C.c_double(0), C.c_double(0), C.c_double(0)]
        npmax = C.c_int(0)
        L.pawb200_momentum_grid_size(self.wf.wf_ptr, C.byref(nb[0]), C.byref(nb[1]), C.byref(nb[2]),
                                     C.byref(npmax), self.momentum_encut)
        check()
        grid = np.zeros(3 * max(npmax.value, 1), dtype=np.int32)
        actual_size = L.pawb200_get_momentum_grid(ip(grid), self.wf.wf_ptr, nb[0].value, nb[1].value, nb[2].value,
                                                  self.momentum_encut)
        check()
        self.ggrid = np.ascontiguousarray(grid[:3 * actual_size])
        self.gbounds = np.zeros(6, dtype=np.int32)
        self.gdim = np.zeros(3, dtype=np.int32)
        L.pawb200_grid_bounds(ip(self.gbounds), ip(self.gdim), ip(self.ggrid), actual_size)
        self.grid3d = -1 * np.ones(int(self.gdim[0]) * int(self.gdim[1]) * int(self.gdim[2]), dtype=np.int32)
        L.pawb200_list_to_grid_map(ip(self.grid3d), ip(self.gbounds), ip(self.gdim), ip(self.ggrid), actual_size)

    def _setup_transforms(self):
        self.elem_density_transforms = _lib.lib().pawb200_get_all_transforms(self.wf.wf_ptr, self.momentum_encut)
        check()

    def _get_ggrid(self):
        return self.ggrid.copy()

    def _get_momentum_matrix_elems(self, b1, k1, s1, b2, k2, s2):
        numg = self.ggrid.shape[0] // 3
        res = np.zeros(numg, dtype=np.complex128)
        _lib.lib().pawb200_get_momentum_matrix(
            res.ctypes.data_as(_lib.c_dbl_p), numg, ip(self.ggrid), self.wf.wf_ptr, ip(self.wf.nums),
            dp(self.wf.coords), int(b1), int(k1), int(s1), int(b2), int(k2), int(s2), self.elem_density_transforms,
            self.momentum_encut)
        check()
        return res

    def _get_reciprocal_fullfw(self, b, k, s):
        numg = self.ggrid.shape[0] // 3
        res = np.zeros(numg, dtype=np.complex128)
        _lib.lib().pawb200_fullwf_reciprocal(res.ctypes.data_as(_lib.c_dbl_p), ip(self.ggrid), self.wf.wf_ptr, numg,
                                             int(b), int(k + s * self.wf.nwk), ip(self.wf.nums), dp(self.wf.coords))
        check()
        return res

    def _get_g_from_fullfw(self, b1, k1, s1, b2, k2, s2, G):
        vec1 = self._get_reciprocal_fullfw(b1, k1, s1)
        vec2 = self._get_reciprocal_fullfw(b2, k2, s2)
        GP = np.array(G, dtype=np.int32)
        out = np.zeros(2)
        _lib.lib().pawb200_quick_overlap(ip(GP), vec1.ctypes.data_as(_lib.c_dbl_p), vec2.ctypes.data_as(_lib.c_dbl_p),
                                         self.ggrid.shape[0] // 3, ip(self.ggrid), ip(self.grid3d), ip(self.gbounds),
                                         ip(self.gdim), dp(out))
        return complex(out[0], out[1])
