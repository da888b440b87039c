"""mdgrad_b200 - B200-native (sm_100a) implementation of torchmd/mdgrad's MD hot path.

Modules mirror the reference's `torchmd` package (system, topology, potentials, interface, md,
sovlers, observable); `_lib` is the ctypes binding of the C ABI in include/mdgrad_b200.h.
"""
__version__ = "0.1.0"
