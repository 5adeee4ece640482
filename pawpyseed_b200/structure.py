"""Minimal stand-ins for the pymatgen objects the reference API touches
(``Structure``: iteration, ``.sites``, ``.lattice.matrix/.abc``, ``site.frac_coords``,
``site.specie.symbol``, ``site.distance``).  A real pymatgen Structure works as well;
these exist because pymatgen is not a dependency of the engine."""
from __future__ import annotations

import numpy as np


class Lattice:
    def __init__(self, matrix):
        self.matrix = np.asarray(matrix, dtype=np.float64).reshape(3, 3)

    @property
    def abc(self):
        return tuple(np.linalg.norm(self.matrix, axis=1))

    @property
    def volume(self):
        return abs(np.linalg.det(self.matrix))

    def __eq__(self, other):
        return np.allclose(self.matrix, np.asarray(other.matrix), atol=1e-8)

    def __ne__(self, other):
        return not self.__eq__(other)


class _Specie:
    def __init__(self, symbol):
        self.symbol = symbol


class Site:
    def __init__(self, symbol, frac_coords, lattice):
        self.specie = _Specie(symbol)
        self.frac_coords = np.asarray(frac_coords, dtype=np.float64)
        self.lattice = lattice

    def distance(self, other):
        """Shortest periodic-image distance (what pymatgen's PeriodicSite.distance returns)."""
        d = self.frac_coords - other.frac_coords
        d -= np.round(d)
        best = np.inf
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                for k in (-1, 0, 1):
                    best = min(best, np.linalg.norm((d + np.array([i, j, k])) @ self.lattice.matrix))
        return best


class Structure:
    def __init__(self, lattice, species, frac_coords):
        self.lattice = lattice if isinstance(lattice, Lattice) else Lattice(lattice)
        self.sites = [Site(s, c, self.lattice) for s, c in zip(species, np.asarray(frac_coords).reshape(-1, 3))]

    def __iter__(self):
        return iter(self.sites)

    def __len__(self):
        return len(self.sites)

    def __getitem__(self, i):
        return self.sites[i]
