"""pawpyseed_b200 - B200-native PAW band-projection engine behind pawpyseed's API.

    from pawpyseed_b200 import Wavefunction, Projector, CoreRegion, NCLWavefunction

Compute lives in libpawb200.so (hand-written sm_100a CUDA, C ABI in include/pawpyseed_b200.h);
this package is the host-side mirror of the reference's Python / Cython interface.
"""
from ._lib import PAWpyError  # noqa: F401

__all__ = ["Wavefunction", "CoreRegion", "Pseudopotential", "Projector", "NCLWavefunction", "MomentumMatrix",
           "PAWpyError"]


def __getattr__(name):
    if name in ("Wavefunction", "CoreRegion", "Pseudopotential"):
        from . import wavefunction
        return getattr(wavefunction, name)
    if name == "Projector":
        from .projector import Projector
        return Projector
    if name == "NCLWavefunction":
        from .noncollinear import NCLWavefunction
        return NCLWavefunction
    if name == "MomentumMatrix":
        from .momentum import MomentumMatrix
        return MomentumMatrix
    raise AttributeError(name)
