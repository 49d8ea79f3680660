"""pynngp_b200 -- B200-native NNGP likelihood engine behind the pyNNGP API.

``from pynngp_b200 import NNGP`` (or the drop-in alias ``from pyNNGP import NNGP``).
"""
from .kernels import Exponential, Kernel, Matern  # noqa: F401
from .nngp import NNGP, NeighborSets  # noqa: F401
