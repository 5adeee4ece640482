"""Symmetry-equivalent k-points and the operators that connect them (host side of SURVEY 8 row f2).

Mirrors the interface of pawpyseed.core.symmetry (symmetry.py:11-164): ``get_symmops``,
``get_nosym_kpoints`` and ``get_kpt_mapping`` return the (k-point index, operator index,
time-reversal flag) triples that `expand_symm_wf` consumes.  Finding the space group itself needs
pymatgen/spglib; every function also takes ``symmops=`` (operators already expressed in reciprocal
fractional coordinates) so the path stays usable, and testable, without them.
"""
from __future__ import annotations

import numpy as np

from ._lib import PAWpyError


class SymmOp:
    """Minimal stand-in for pymatgen's SymmOp: `.rotation_matrix`, `.translation_vector`."""

    def __init__(self, rotation, translation=(0.0, 0.0, 0.0)):
        self.rotation_matrix = np.array(rotation, dtype=np.float64).reshape(3, 3)
        self.translation_vector = np.array(translation, dtype=np.float64).reshape(3)

    @classmethod
    def from_rotation_and_translation(cls, rotation, translation=(0.0, 0.0, 0.0)):
        return cls(rotation, translation)


def get_symmops(structure, symprec):
    """Space-group operations in reciprocal-lattice fractional coordinates (symmetry.py:11-31):
    R' = A R_cart A^-1 and t' = t_cart A^-1 with A the row-vector lattice matrix."""
    try:
        from pymatgen.symmetry.analyzer import SpacegroupAnalyzer
    except ImportError as e:
        raise PAWpyError("space-group detection needs pymatgen (%s); pass symmops= explicitly" % e)
    sga = SpacegroupAnalyzer(structure, symprec * max(structure.lattice.abc))
    A = np.asarray(structure.lattice.matrix)
    Ainv = np.asarray(structure.lattice.inv_matrix)
    return [SymmOp(A @ op.rotation_matrix @ Ainv, op.translation_vector @ Ainv)
            for op in sga.get_symmetry_operations(cartesian=True)]


def _same_mod_lattice(a, b, tol=1e-4):
    d = (a - b) % 1
    return bool(np.all((np.abs(d) < tol) | (np.abs(1 - d) < tol)))


def _in_trs_lower_half(k):
    """True for the half of the zone that time reversal makes redundant (symmetry.py:58-67, 93-102)."""
    return (k[2] < -1e-6 or (abs(k[2]) < 1e-6 and k[1] < -1e-6)
            or (abs(k[2]) < 1e-6 and abs(k[1]) < 1e-6 and k[0] < -1e-6))


def get_nosym_kpoints(kpts, structure=None, init_kpts=None, symprec=1e-4, gen_trsym=True, fil_trsym=True,
                      symmops=None):
    """All k-points generated from `kpts` by the crystal symmetry, each with the index of its source
    k-point, the operator used and whether time reversal was applied (symmetry.py:34-118).
    Returns (allkpts, orig_kptnums, op_nums, symmops, trs)."""
    if symmops is None:
        symmops = get_symmops(structure, symprec)
    allkpts = [] if init_kpts is None else [np.asarray(k, dtype=np.float64) for k in init_kpts]
    orig_kptnums, op_nums, trs = [], [], []
    kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    for tr in ((0, 1) if gen_trsym else (0,)):
        sign = -1.0 if tr else 1.0
        for i, op in enumerate(symmops):
            for k, kpt in enumerate(kpts):
                new = sign * (op.rotation_matrix @ kpt)
                new -= np.around(new)
                new[np.abs(new + 0.5) < 1e-5] = 0.5
                if fil_trsym and (_in_trs_lower_half(new) or (tr and new[2] < -1e-10)):
                    continue
                if any(_same_mod_lattice(new, other) for other in allkpts):
                    continue
                allkpts.append(new)
                orig_kptnums.append(k)
                op_nums.append(i)
                trs.append(tr)
    return np.array(allkpts), orig_kptnums, op_nums, symmops, trs


def get_kpt_mapping(allkpts, kpts, structure=None, symprec=1e-4, gen_trsym=True, symmops=None):
    """For each k-point of `allkpts`, the (source k-point, operator, time-reversal) that produces it
    (symmetry.py:121-164); plain operations are preferred over time-reversed ones.
    Returns (orig_kptnums, op_nums, symmops, trs)."""
    if symmops is None:
        symmops = get_symmops(structure, symprec)
    kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    orig_kptnums, op_nums, trs = [], [], []
    for target in np.asarray(allkpts, dtype=np.float64).reshape(-1, 3):
        found = None
        for tr in (0, 1):
            sign = -1.0 if tr else 1.0
            for i, op in enumerate(symmops):
                for k, kpt in enumerate(kpts):
                    if _same_mod_lattice(sign * (op.rotation_matrix @ kpt), target):
                        found = (k, i, tr)
                        break
                if found:
                    break
            if found:
                break
        if not found:
            raise PAWpyError("Could not find kpoint mapping to %s" % str(target))
        orig_kptnums.append(found[0])
        op_nums.append(found[1])
        trs.append(found[2])
    return orig_kptnums, op_nums, symmops, trs


def make_c_ops(op_nums, symmops):
    """Flatten the selected operators for the C ABI (pawpyc.pyx:64-71): ops[9*n], drs[3*n]."""
    ops = np.concatenate([symmops[i].rotation_matrix.reshape(-1) for i in op_nums]) if len(op_nums) else np.zeros(0)
    drs = np.concatenate([symmops[i].translation_vector for i in op_nums]) if len(op_nums) else np.zeros(0)
    return np.ascontiguousarray(ops, dtype=np.float64), np.ascontiguousarray(drs, dtype=np.float64)
