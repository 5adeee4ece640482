"""MomentumMatrix - same public class as pawpyseed.core.momentum (momentum.py:4-91)."""
from __future__ import annotations

from . import pawpyc


class MomentumMatrix(pawpyc.CMomentumMatrix):
    """< psi1 | exp(i (G + k1 - k2).r) | psi2 > for the G vectors of a cutoff sphere (`momentum_grid`), and the
    plane-wave expansion of all-electron bands.  `encut` defaults to 4 * wf.encut.  The grid is k-independent."""

    def __init__(self, wf, encut=None):
        wf.check_c_projectors()
        if encut is None:
            encut = 4 * wf.encut
        super().__init__(wf, encut)

    @property
    def momentum_grid(self):
        grid = self._get_ggrid()
        return grid.reshape((grid.shape[0] // 3, 3))

    def get_momentum_matrix_elems(self, b1, k1, s1, b2, k2, s2):
        """< b1,k1,s1 | exp(i (G + k1 - k2).r) | b2,k2,s2 > for each G in momentum_grid (momentum.py:45-62)."""
        self.wf.check_bks_spec(b1, k1, s1)
        self.wf.check_bks_spec(b2, k2, s2)
        return self._get_momentum_matrix_elems(b1, k1, s1, b2, k2, s2)

    def get_reciprocal_fullfw(self, b, k, s):
        """C(b,k,s,G) with |b,k,s> = V^-1/2 sum_G C exp(i (k+G).r) (momentum.py:64-77)."""
        self.wf.check_bks_spec(b, k, s)
        return self._get_reciprocal_fullfw(b, k, s)

    def g_from_wf(self, b1, k1, s1, b2, k2, s2, G):
        """Slow cross-check of one matrix element through the two plane-wave expansions (momentum.py:79-91)."""
        self.wf.check_bks_spec(b1, k1, s1)
        self.wf.check_bks_spec(b2, k2, s2)
        return self._get_g_from_fullfw(b1, k1, s1, b2, k2, s2, G)
