"""Small Python helpers with the names of pawpyseed/core/utils.py."""
from ._lib import PAWpyError  # noqa: F401


class PAWpyWarning(Warning):
    def __init__(self, msg):
        self.msg = msg


def check_spin(spin, nspin):
    """utils.py:26-40."""
    if spin >= 0:
        if spin >= nspin:
            raise PAWpyError("spin must be less than nspin. spin is %d, nspin is %d" % (spin, nspin))
        return 1
    return nspin


def el(site):
    """utils.py:43-48."""
    return site.specie.symbol
