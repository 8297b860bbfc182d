"""ctypes binding of oracle/liborc.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY: may be
imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
the product package zkm_b200 never imports it."""
import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "liborc.so"
_lib = None


def build():
    subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = list(_HERE.glob("*.cpp")) + list(_HERE.glob("*.h")) + list((_HERE.parent / "zkm_b200/csrc/tables").glob("*.h"))
    if not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs):
        build()
    lib = C.CDLL(str(LIB_PATH))
    u64p = C.POINTER(C.c_uint64)
    u64 = C.c_uint64
    lib.orc_last_error.restype = C.c_char_p
    for f in ("orc_mul", "orc_add", "orc_sub"):
        getattr(lib, f).argtypes = [u64, u64]
        getattr(lib, f).restype = u64
    lib.orc_inv.argtypes = [u64]; lib.orc_inv.restype = u64
    lib.orc_root_of_unity.argtypes = [C.c_uint]; lib.orc_root_of_unity.restype = u64
    lib.orc_ext_mul.argtypes = [u64p, u64p, u64p]
    lib.orc_ext_inv.argtypes = [u64p, u64p]
    lib.orc_poseidon_permute.argtypes = [u64p, C.c_int]
    lib.orc_poseidon_permute_many.argtypes = [u64p, C.c_size_t]
    lib.orc_hash_or_noop.argtypes = [u64p, C.c_size_t, u64p]
    lib.orc_two_to_one.argtypes = [u64p, u64p, u64p]
    lib.orc_ntt.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_int]
    lib.orc_commit.argtypes = [C.POINTER(u64p), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u64p]
    lib.orc_commit.restype = C.c_void_p
    lib.orc_batch_free.argtypes = [C.c_void_p]
    lib.orc_batch_get_coeffs.argtypes = [C.c_void_p, C.c_uint32, u64p]
    lib.orc_batch_get_lde.argtypes = [C.c_void_p, C.c_uint32, u64p]
    lib.orc_batch_open.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p]
    _lib = lib
    return lib


def u64ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def col_ptrs(cols: np.ndarray):
    ncols, n = cols.shape
    arr = (C.POINTER(C.c_uint64) * ncols)()
    for i in range(ncols):
        arr[i] = C.cast(cols.ctypes.data + i * n * 8, C.POINTER(C.c_uint64))
    return arr
