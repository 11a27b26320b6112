"""swarm_b200 — Blackwell-native (sm_100a) engine for swarm's amplicon neighbour-search hot path.

The product is two in-tree shared libraries with plain-C ABIs:
  * ``libswarm_b200.so``       CUDA engine            (include/swarm_b200.h)
  * ``libswarm_b200_host.so``  FASTA database/writers (include/swarm_b200_host.h)
This package is only the thin ctypes binding used by the tests, ``bench.py`` and the multi-GPU
launcher; there is no Python or CPU implementation of the algorithms here — if the CUDA library is
missing or no GPU is present every engine call raises.
"""
from .ffi import Engine, HostDb, D1Result, DnResult, DerepResult, scoring, EngineError, lib_paths, ENUM_FULL, ENUM_HALF, ENUM_JOIN, NONE  # noqa: F401

__all__ = ["Engine", "HostDb", "D1Result", "DnResult", "DerepResult", "scoring", "EngineError", "lib_paths", "ENUM_FULL", "ENUM_HALF", "ENUM_JOIN", "NONE"]
