"""c3sc_b200 -- B200-native Bellman-backup hot path of goroda/c3sc.

The product is the C-ABI shared library `lib/libc3sc_b200.so` (hand-written
sm_100a kernels, `csrc/`), declared in `include/c3sc_b200.h`, plus the C host
mirror of the reference API.  The Python modules here are plumbing for tests
and the benchmark: `capi` (ctypes binding), `configs` (BASELINE.json problem
definitions) and `synthetic` (seeded inputs).
"""
from . import configs, synthetic  # noqa: F401

__all__ = ["configs", "synthetic", "capi"]
