"""zkm_b200 — B200-native STARK proving path for zkMIPS/zkm behind a C ABI (include/zkm_b200.h).

The Python layer is a thin ctypes mirror used by tests and bench.py; the product is
libzkm_b200.so (CUDA sm_100a kernels + C++ host).  There is no CPU fallback: loading fails loudly
if the library has not been built, and zkm_b200_init fails if no Blackwell GPU is present.
"""
from .lib import load, ZkmError, Table, StarkConfig  # noqa: F401
