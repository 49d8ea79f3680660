# Drop-in import path of the reference package (pyNNGP/__init__.py:1 upstream): the same name,
# now backed by the B200 engine in pynngp_b200.
from pynngp_b200.nngp import NNGP  # noqa: F401
