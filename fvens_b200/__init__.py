"""fvens_b200: B200-native residual engine behind the FVENS class surface.

The product is the C-ABI shared library `libfvens_b200.so` (include/fvens_b200.h) and the C++ host
classes in fvens_b200/host/. This Python package is the binding used by the tests and bench.py.
"""
from . import lib, synth  # noqa: F401
