"""pawpyseed/core/noncollinear.py on the GPU engine (spinor wavefunctions)."""
from __future__ import annotations

import numpy as np

from . import pawpyc
from .utils import PAWpyError
from .wavefunction import Wavefunction


class NCLWavefunction(pawpyc.CNCLWavefunction, Wavefunction):
    def __init__(self, struct, pwf, cr, dim, symprec=1e-4, setup_projectors=False):
        """noncollinear.py:5-31."""
        self.band_props = pwf.band_props.copy(order="C")
        pawpyc.CWavefunction.__init__(self, pwf)
        if not self.ncl:
            raise PAWpyError("Pseudowavefunction is collinear! Call Wavefunction(...) instead")
        self.structure = struct
        self.symprec = symprec
        self.cr = cr
        self.dim = np.array(dim).astype(np.int32)
        if setup_projectors:
            self.check_c_projectors()

    @classmethod
    def from_arrays(cls, struct, wavecar, cr, dim, kpts, weights, band_props=(0.0, 0.0, 0.0, False),
                    symprec=1e-4, setup_projectors=False):
        pwf = pawpyc.PWFPointer.from_arrays(wavecar, kpts, weights, band_props)
        return cls(struct, pwf, cr, dim, symprec, setup_projectors)

    def desymmetrized_copy(self, allkpts=None, weights=None):
        raise NotImplementedError()

    def get_realspace_density(self, dim=None):
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim))
        return self._get_realspace_density()

    def write_state_realspace(self, b, k, s, fileprefix="", dim=None, scale=1, remove_phase=False):
        """noncollinear.py:98-147."""
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim))
        base = "%sB%dK%dS%d" % (fileprefix, b, k, s)
        names = ["%s_%s.vasp" % (base, t) for t in ("UP_REAL", "UP_IMAG", "DOWN_REAL", "DOWN_IMAG")]
        res0, res1 = self._write_realspace_state(*names, scale, b, k, s, remove_phase=remove_phase)
        for n in names:
            self._convert_to_vasp_volumetric(n, self.dim)
        return res0, res1

    def write_density_realspace(self, filename="PYAECCAR.vasp", dim=None, scale=1):
        """noncollinear.py:149-173."""
        self.check_c_projectors()
        if dim is not None:
            self.update_dim(np.array(dim))
        res = self._write_realspace_density(filename, scale)
        self._convert_to_vasp_volumetric(filename, self.dim)
        return res
