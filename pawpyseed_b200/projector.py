"""pawpyseed/core/projector.py on the GPU engine: `Projector(wf, basis)` with
`single_band_projection`, `proportion_conduction`, `defect_band_analysis`.

This module restates the Python API layer of pawpyseed (pawpyseed/core/projector.py, Copyright (c) 2017 Kyle Bystrom,
BSD 3-clause licence - see the upstream LICENSE) on top of the B200 engine: class / method names, argument
meaning and the bookkeeping code around the compute calls follow the reference so that user code switches by
import path only.  It is the drop-in surface required by the integration boundary, not an independent design.
"""
from __future__ import annotations

import time

import numpy as np

from . import pawpyc
from .pawpyc import Timer
from .utils import PAWpyError, el


class Projector(pawpyc.CProjector):
    METHODS = ["pseudo", "realspace", "aug_recip", "aug_real"]

    def __init__(self, wf, basis, unsym_basis=False, unsym_wf=False, method="aug_real", symmops=None):
        """projector.py:42-113.  `symmops` (extension): space-group operators in reciprocal fractional coordinates
        for the unsym_* options, instead of the pymatgen space-group search (see symmetry.get_symmops)."""
        self.method = method
        if self.method == "pseudo":
            self._single_band_projection = self._single_band_projection_pseudo
        elif self.method == "aug_real":
            self._single_band_projection = self._single_band_projection_aug_real
        elif self.method == "realspace":
            self._single_band_projection = self._single_band_projection_realspace
        elif self.method == "aug_recip":
            self._single_band_projection = self._single_band_projection_aug_recip
        else:
            raise PAWpyError("method not recognized for Projector")
        if wf.ncl or basis.ncl:
            raise PAWpyError("Projection not supported for noncollinear case!")
        # projector.py:77-95: bring both onto one unreduced mesh (GPU remap, pawb200_expand_symm_wf)
        if unsym_basis and unsym_wf:
            basis = basis.desymmetrized_copy(symmops=symmops)
            wf = wf.desymmetrized_copy(basis.kpts, basis.kws, symmops=symmops)
        elif unsym_wf:
            if basis.kpts.shape[0] < wf.kpts.shape[0]:
                raise PAWpyError("Basis doesn't have enough kpoints, needs to be desymmetrized!")
            wf = wf.desymmetrized_copy(basis.kpts, basis.kws, symmops=symmops)
        elif unsym_basis:
            if wf.kpts.shape[0] < basis.kpts.shape[0]:
                raise PAWpyError("Defect doesn't have enough kpoints, needs to be desymmetrized!")
            basis = basis.desymmetrized_copy(wf.kpts, wf.kws, symmops=symmops)
        if basis.kpts.shape != wf.kpts.shape:
            raise PAWpyError("k-point grids for projection are not matched.")
        if np.linalg.norm(basis.kpts - wf.kpts) > 1e-10:
            raise PAWpyError("k-point grids for projection are not matched.")
        if np.linalg.norm(basis.kws - wf.kws) > 1e-10:
            raise PAWpyError("k-point weights for projection are not matched.")
        if wf.structure.lattice != basis.structure.lattice:
            raise PAWpyError("Need the lattice to be the same for projections, and they are not")
        if wf.nspin != basis.nspin:
            # the reference reads out of bounds here; refuse instead
            raise PAWpyError("wf and basis must have the same number of spin channels")
        if self.method != "pseudo":
            basis.check_c_projectors()
            wf.check_c_projectors()
        super().__init__(wf, basis)
        if "aug" in self.method:
            self.setup_overlap()

    def make_site_lists(self):
        """projector.py:115-160 (including its `sites[i]` quirk for rmax2, guarded against IndexError)."""
        ref_sites = self.basis.structure.sites
        sites = self.wf.structure.sites
        M_R, M_S = [], []
        for i in range(len(ref_sites)):
            for j in range(len(sites)):
                if ref_sites[i].distance(sites[j]) <= 0.02 and el(ref_sites[i]) == el(sites[j]):
                    M_R.append(i)
                    M_S.append(j)
        N_R = [i for i in range(len(ref_sites)) if i not in M_R]
        N_S = [j for j in range(len(sites)) if j not in M_S]
        N_RS = []
        for i in N_R:
            for j in N_S:
                rmax1 = self.basis.cr.pps[el(ref_sites[i])].rmax
                rmax2 = self.wf.cr.pps[el(sites[min(i, len(sites) - 1)])].rmax
                if ref_sites[i].distance(sites[j]) < rmax1 + rmax2:
                    N_RS.append((i, j))
        return M_R, M_S, N_R, N_S, N_RS

    def setup_overlap(self, site_cat=None):
        """projector.py:162-187.  `site_cat` lets callers supply precomputed site lists."""
        if site_cat is None:
            M_R, M_S, N_R, N_S, N_RS = self.make_site_lists()
            if len(N_RS) > 0:
                N_RS_R, N_RS_S = zip(*N_RS)
            else:
                N_RS_R, N_RS_S = [], []
            site_cat = [M_R, M_S, N_R, N_S, N_RS_R, N_RS_S]
        self.site_cat = [list(x) for x in site_cat]
        start = time.monotonic()
        if self.method not in ("aug_recip", "aug_real"):
            raise PAWpyError("method must be aug type for setup_overlap call")
        self._setup_overlap(self.site_cat, self.method == "aug_recip")
        Timer.overlap_time(time.monotonic() - start)

    def _single_band_projection_pseudo(self, band_num):
        return self.wf.pseudoprojection(band_num, self.basis)

    def _single_band_projection_realspace(self, band_num, dim=None):
        """projector.py:199-208: AE states on a real-space grid (default: the fine grid 2*dim), integrated."""
        if dim is None:
            dim = self.wf.dim * 2
        return self._realspace_projection(band_num, dim)

    def _single_band_projection_aug_real(self, band_num, flip_spin=False):
        """projector.py:210-223."""
        res = self.wf.pseudoprojection(band_num, self.basis, flip_spin)
        start = time.monotonic()
        self._add_augmentation_terms(res, band_num, flip_spin)
        Timer.augmentation_time(time.monotonic() - start)
        return res

    def _single_band_projection_aug_recip(self, band_num, flip_spin=False):
        """projector.py:225-236: augmentation carried through the plane-wave basis (low-pass filtered partial
        waves -> FFT grid -> forward FFT), then plane-wave dot products."""
        res = self.wf.pseudoprojection(band_num, self.basis, flip_spin)
        self._projection_recip(res, band_num, flip_spin)
        return res

    def single_band_projection(self, band_num, **kwargs):
        """projector.py:238-272: result[b*nwk*nspin + s*nwk + k] = <basis;b,k,s|wf;band_num,k,s>."""
        if band_num >= self.wf.nband or band_num < 0:
            raise ValueError("Band index out of range (0-indexed)")
        return self._single_band_projection(band_num, **kwargs)

    def projection_matrix(self, flip_spin=False, kappa_range=None):
        """Extension: every band pair at once, out[kappa, b_wf, b_basis] (one GEMM pass on the GPU)."""
        return self._projection_matrix(flip_spin, kappa_range, pseudo_only=(self.method == "pseudo"))

    def proportion_conduction(self, band_num, spinpol=False):
        """projector.py:384-435."""
        basis = self.basis
        nband, nwk, nspin = basis.nband, basis.nwk, basis.nspin
        occs = self.basis._get_occs()
        res = self.single_band_projection(band_num)
        if spinpol:
            c, v = np.zeros(nspin), np.zeros(nspin)
            for b in range(nband):
                for s in range(nspin):
                    ind = b * nspin + s
                    prop = np.absolute(res[ind * nwk:(ind + 1) * nwk]) ** 2
                    c[s] += np.dot(prop, (1 - occs[ind * nwk:(ind + 1) * nwk]) * self.wf.kws)
                    v[s] += np.dot(prop, occs[ind * nwk:(ind + 1) * nwk] * self.wf.kws)
        else:
            c, v = 0, 0
            for i in range(nband * nwk * nspin):
                if occs[i] > 0.5:
                    v += np.absolute(res[i]) ** 2 * self.wf.kws[i % nwk] / nspin
                else:
                    c += np.absolute(res[i]) ** 2 * self.wf.kws[i % nwk] / nspin
        if self.method == "pseudo":
            t = v + c
            v /= t
            c /= t
        if spinpol:
            v, c = v.tolist(), c.tolist()
        return v, c

    def defect_band_analysis(self, num_below_ef=20, num_above_ef=20, spinpol=False, return_energies=False,
                             vbmband=None, band_list=None, analyze_all=False):
        """projector.py:437-503."""
        if num_below_ef < 0 or num_above_ef < 0:
            raise ValueError("num_above_ef and num_below_ef must both be nonnegative.")
        nband = self.basis.nband
        occs = self.wf._get_occs()
        if analyze_all:
            totest = [i for i in range(nband)]
        elif band_list:
            totest = band_list[:]
        else:
            vbm = 0
            for i in range(self.wf.nband):
                if occs[i * self.wf.nwk * self.wf.nspin] > 0.5:
                    vbm = i
            if vbmband is not None:
                vbm = vbmband
            min_band, max_band = max(vbm - num_below_ef, 0), min(vbm + num_above_ef, self.wf.nband - 1)
            totest = [i for i in range(min_band, max_band + 1)]
        results = {}
        for b in totest:
            results[b] = self.proportion_conduction(b, spinpol=spinpol)
        if return_energies:
            return results, self.wf._get_energy_list(totest)
        return results
