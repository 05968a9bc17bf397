"""nmch_b200 -- B200-native (sm_100a) Heston Monte-Carlo engine behind the edo01/NMCH method API.

Layout: csrc/ (CUDA kernels + the C ABI of include/nmch_b200.h), capi.py (ctypes binding),
engine.py (handle wrapper), methods.py (mirror of nmch::methods::NMCH_FE_* / NMCH_EM_*),
distributed.py (path sharding + one NCCL allreduce of the moments).
There is no CPU fallback: importing is cheap, computing needs the CUDA library and a GPU.
"""
from . import capi, engine, methods  # noqa: F401
from .engine import Engine, Group, Moments  # noqa: F401

__all__ = ["capi", "engine", "methods", "Engine", "Group", "Moments"]
