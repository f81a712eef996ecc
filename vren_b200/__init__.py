"""vren_b200 — B200-native (sm_100a) implementation of vren's data-parallel compute core.

The product is the CUDA shared library `libvrenb200.so` (C ABI in include/vrenb200.h) plus the C++ facade
in include/vren/ that mirrors the reference's functor classes.  This Python package is the ctypes harness
used by the tests and bench; it has no CPU fallback.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
