"""Python-facing API of the reference's pawpyseed/core/wavefunction.py on the GPU engine.

`Pseudopotential`, `CoreRegion` and `Wavefunction` keep the reference's names, constructor
arguments and method names (file:line cited per method) so user code switches by import
path only.  pymatgen is optional: `from_files` / `from_directory` need it (they parse
POSCAR / POTCAR / vasprun.xml exactly like the reference); everything else accepts either
pymatgen objects or the light `structure.Structure`.

This module restates the Python API layer of pawpyseed (pawpyseed/core/wavefunction.py, Copyright (c) 2017 Kyle Bystrom,
BSD 3-clause licence - see the upstream LICENSE) on top of the B200 engine: class / method names, argument
meaning and the bookkeeping code around the compute calls follow the reference so that user code switches by
import path only.  It is the drop-in surface required by the integration boundary, not an independent design.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import pawpyc
from .utils import PAWpyError, el


class Pseudopotential:
    """POTCAR text parser, wavefunction.py:17-140 (attribute names identical)."""

    def __init__(self, data):
        nonradial, radial = data.split("PAW radial sets", 1)
        partial_waves = radial.split("pseudo wavefunction")
        gridstr, partial_waves = partial_waves[0], partial_waves[1:]
        self.pswaves, self.aewaves, self.recipprojs, self.realprojs = [], [], [], []
        self.nonlocalprojs, self.ls = [], []
        auguccstr, gridstr = gridstr.split("grid", 1)
        gridstr, aepotstr = gridstr.split("aepotential", 1)
        aepotstr, corechgstr = aepotstr.split("core charge-density", 1)
        try:
            corechgstr, kenstr = corechgstr.split("kinetic energy-density", 1)
            kenstr, pspotstr = kenstr.split("pspotential", 1)
        except ValueError:
            corechgstr, pspotstr = corechgstr.split("pspotential", 1)
        self.grid = self.make_nums(gridstr)
        augstr, _ = auguccstr.split("uccopancies in atom", 1)
        _, augstr = augstr.split("augmentation charges (non sperical)", 1)
        self.augs = self.make_nums(augstr)
        for pwave in partial_waves:
            lst = pwave.split("ae wavefunction", 1)
            self.pswaves.append(self.make_nums(lst[0]))
            self.aewaves.append(self.make_nums(lst[1]))
        projstrs = nonradial.split("Non local Part")
        topstr, projstrs = projstrs[0], projstrs[1:]
        self.T = float(topstr[-22:-4])
        topstr, _ = topstr[:-22].split("atomic pseudo charge-density", 1)
        try:
            topstr, _ = topstr.split("core charge-density (partial)", 1)
        except ValueError:
            pass
        settingstr, _ = topstr.split("local part", 1)
        for projstr in projstrs:
            lst = projstr.split("Reciprocal Space Part")
            nonlocalvals, projs = lst[0], lst[1:]
            self.rmax = self.make_nums(nonlocalvals.split()[2])[0]
            nonlocalvals = self.make_nums(nonlocalvals)
            l = nonlocalvals[0]
            self.nonlocalprojs.append(nonlocalvals[2:])
            for proj in projs:
                recipproj, realproj = proj.split("Real Space Part")
                self.recipprojs.append(self.make_nums(recipproj))
                self.realprojs.append(self.make_nums(realproj))
                self.ls.append(l)
        settingstr, _ = settingstr.split("STEP   =")
        self.ndata = int(settingstr.split()[-1])
        self.projgrid = np.arange(len(self.realprojs[0])) * self.rmax / len(self.realprojs[0])
        self.step = (self.projgrid[0], self.projgrid[1])

    @staticmethod
    def make_nums(numstring):
        return np.array(numstring.split(), dtype=np.float64)


class CoreRegion:
    """wavefunction.py:143-166.  Accepts a pymatgen Potcar or a dict {symbol: pseudopotential-like}."""

    def __init__(self, potcar):
        self.pps = {}
        if isinstance(potcar, dict):
            self.pps = dict(potcar)
        else:
            for potsingle in potcar:
                self.pps[potsingle.element] = Pseudopotential(potsingle.data[:-15])


class Wavefunction(pawpyc.CWavefunction):
    """wavefunction.py:169-606."""

    def __init__(self, struct, pwf, cr, dim, symprec=1e-4, setup_projectors=False):
        self.band_props = pwf.band_props.copy(order="C")
        super().__init__(pwf)
        if self.ncl:
            raise PAWpyError("Pseudowavefunction is noncollinear! Call NCLWavefunction(...) instead")
        self.structure = struct
        self.symprec = symprec
        self.cr = cr
        self.dim = np.array(dim).astype(np.int32)
        if len(dim) != 3:
            raise PAWpyError("Grid dimensions must be length 3")
        if setup_projectors:
            self.check_c_projectors()

    # -- index checks (wavefunction.py:218-243) ------------------------------------------------
    def check_band_index(self, b):
        if b < 0 or b >= self.nband:
            raise ValueError("Invalid band {}. Should be in range [{}, {}]".format(b, 0, self.nband - 1))

    def check_kpoint_index(self, k):
        if k < 0 or k >= self.nwk:
            raise ValueError("Invalid kpoint index {}. Should be in range [{}, {}]".format(k, 0, self.nwk - 1))

    def check_spin_index(self, s):
        if s < 0 or s >= self.nspin:
            raise ValueError("Spin must be 0 for non-spin-polarized or 0 or 1 for spin-polarized.")

    def check_bks_spec(self, b, k, s):
        self.check_band_index(b)
        self.check_kpoint_index(k)
        self.check_spin_index(s)

    def update_dim(self, dim):
        self.dim = np.array(dim, dtype=np.int32)
        self.update_dimv(dim)

    def desymmetrized_copy(self, allkpts=None, weights=None, symprec=None, time_reversal_symmetry=True,
                           symmops=None):
        """Copy of self on a k-point mesh that is not reduced by crystal symmetry (wavefunction.py:249-279).
        The remap runs on the GPU (pawb200_expand_symm_wf).  `symmops` (operators in reciprocal fractional
        coordinates, see symmetry.get_symmops) is an extension that skips the pymatgen space-group search."""
        if not symprec:
            symprec = self.symprec
        pwf = self._desymmetrized_pwf(self.structure, self.band_props, allkpts, weights, symprec,
                                      time_reversal_symmetry, symmops=symmops)
        return Wavefunction(self.structure, pwf, self.cr, self.dim, symprec=symprec)

    # -- constructors (wavefunction.py:281-384) ---------------------------------------------------
    @staticmethod
    def from_files(struct="CONTCAR", wavecar="WAVECAR", cr="POTCAR", vr="vasprun.xml", setup_projectors=False):
        for fname in [struct, wavecar, cr, vr]:
            if not os.path.isfile(fname):
                raise FileNotFoundError(f"File {fname} does not exist.")
        try:
            from pymatgen.io.vasp.inputs import Poscar, Potcar
            from pymatgen.io.vasp.outputs import Vasprun
        except ImportError as e:
            raise PAWpyError("from_files needs pymatgen to parse VASP text files: %s" % e)
        vr = Vasprun(vr)
        dim = np.array([vr.parameters["NGX"], vr.parameters["NGY"], vr.parameters["NGZ"]])
        symprec = vr.parameters["SYMPREC"]
        pwf = pawpyc.PWFPointer(wavecar, vr)
        return Wavefunction(Poscar.from_file(struct).structure, pwf, CoreRegion(Potcar.from_file(cr)), dim,
                            symprec, setup_projectors)

    @staticmethod
    def from_directory(path, setup_projectors=False):
        filepaths = [str(os.path.join(path, d)) for d in ["CONTCAR", "WAVECAR", "POTCAR", "vasprun.xml"]]
        return Wavefunction.from_files(*(filepaths + [setup_projectors]))

    @staticmethod
    def from_atomate_directory(path, setup_projectors=False):
        paths = []
        for file in ["CONTCAR", "WAVECAR", "POTCAR", "vasprun.xml"]:
            for suffix in (".relax2.gz", ".relax1.gz", ".gz", ""):
                filepat = os.path.join(path, file + suffix)
                if os.path.exists(filepat):
                    break
            else:
                print(f"Could not find {file}! Skipping this defect...")
                return False
            paths.append(filepat)
        return Wavefunction.from_files(*(paths + [setup_projectors]))

    @classmethod
    def from_arrays(cls, struct, wavecar, cr, dim, kpts, weights, band_props=(0.0, 0.0, 0.0, False),
                    symprec=1e-4, setup_projectors=False):
        """Extension: build from an in-memory WAVECAR image / path plus k-points and weights,
        without vasprun.xml (what tests and bench.py use)."""
        pwf = pawpyc.PWFPointer.from_arrays(wavecar, kpts, weights, band_props)
        return cls(struct, pwf, cr, dim, symprec, setup_projectors)

    # -- projector setup (wavefunction.py:386-429) --------------------------------------------------
    def _make_c_projectors(self):
        pps, labels = {}, {}
        for label, e in enumerate(self.cr.pps):
            pps[label] = self.cr.pps[e]
            labels[e] = label
        nums = np.array([labels[el(s)] for s in self.structure], dtype=np.int32)
        coords = np.array([], dtype=np.float64)
        self.num_sites = len(self.structure)
        self.num_elems = len(pps)
        for s in self.structure:
            coords = np.append(coords, s.frac_coords)
        grid_encut = (np.pi * self.dim / np.asarray(self.structure.lattice.abc)) ** 2 / 0.262
        self._c_projector_setup(self.num_elems, self.num_sites, max(grid_encut), nums, coords, self.dim, pps)

    def check_c_projectors(self):
        if not self.projector_owner:
            start = time.monotonic()
            self._make_c_projectors()
            end = time.monotonic()
            pawpyc.Timer.setup_time(end - start)

    # -- real-space (wavefunction.py:431-580) ---------------------------------------------------------
    def get_state_realspace(self, b, k, s, dim=None, remove_phase=False):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim))
        return self._get_realspace_state(b, k, s, remove_phase)

    def get_state_realspace_density(self, b, k, s, dim=None):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim) // 2)
        return self._get_realspace_state_density(b, k, s)

    def get_realspace_density(self, dim=None, bands=None):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim) // 2)
        return self._get_realspace_density()

    def _convert_to_vasp_volumetric(self, filename, dim):
        """wavefunction.py:481-510: prepend a POSCAR-style header to the raw dump."""
        latt = np.asarray(self.structure.lattice.matrix)
        symbols, counts = [], []
        for s in self.structure:
            sym = el(s)
            if symbols and symbols[-1] == sym:
                counts[-1] += 1
            else:
                symbols.append(sym)
                counts.append(1)
        lines = filename + "\n   1.00000000000000\n"
        for r in range(3):
            lines += " %12.6f%12.6f%12.6f\n" % tuple(latt[r, :])
        lines += "".join(["%5s" % s for s in symbols]) + "\n"
        lines += "".join(["%6d" % x for x in counts]) + "\n"
        lines += "Direct\n"
        for site in self.structure:
            lines += "%10.6f%10.6f%10.6f\n" % tuple(site.frac_coords)
        lines += " \n"
        with open(filename) as f:
            nums = f.read()
        with open(filename, "w") as f:
            f.write(lines + "%d %d %d\n" % (dim[0], dim[1], dim[2]) + nums)

    def write_state_realspace(self, b, k, s, fileprefix="", dim=None, scale=1, remove_phase=False):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim))
        filename_base = "%sB%dK%dS%d" % (fileprefix, b, k, s)
        filename1 = "%s_REAL.vasp" % filename_base
        filename2 = "%s_IMAG.vasp" % filename_base
        res = self._write_realspace_state(filename1, filename2, scale, b, k, s, remove_phase)
        self._convert_to_vasp_volumetric(filename1, self.dim)
        self._convert_to_vasp_volumetric(filename2, self.dim)
        return res

    def write_density_realspace(self, filename="PYAECCAR.vasp", dim=None, scale=1, bands=None):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim) // 2)
        res = self._write_realspace_density(filename, scale, bands)
        self._convert_to_vasp_volumetric(filename, self.dim * 2)
        return res
