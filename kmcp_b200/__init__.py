"""kmcp_b200 — B200 (sm_100a) implementation of the `kmcp search` hot path.

The product is kmcp_b200/libkmcp_gpu.so (C ABI: include/kmcp_gpu.h, sources: kmcp_b200/csrc) and the
kmcp-gpu CLI.  `kmcp_b200.api` is the ctypes binding used by tests and bench.py.
"""
from . import api  # noqa: F401

__all__ = ["api"]
