"""graphconvgeo_b200 -- B200-native (sm_100a) GCN propagation hot path of afcarl/graphconvgeo.

Only what the path needs: ``csrc/`` (hand-written CUDA kernels behind the C ABI of
include/gcg.h, built into libgcg.so), the ctypes binding, and the host-side mirror of the
reference's layer / model interface (lasagne_layers.py, mlpconv.py).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "sparse", "lasagne_layers", "mlpconv", "synth"]
__version__ = "0.1.0"
