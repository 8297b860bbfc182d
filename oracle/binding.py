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
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_prove_system.restype = C.c_void_p
    lib.orc_prove_system.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]
    lib.orc_free.argtypes = [C.c_void_p]
    lib.orc_verify_system.argtypes = [C.c_int, u64p, C.c_size_t, C.POINTER(C.c_uint32)]
    lib.orc_check_table_constraints.restype = C.c_long
    lib.orc_check_table_constraints.argtypes = [C.c_int, C.POINTER(u64p), C.c_uint32, C.c_uint32]
    lib.orc_table_fingerprint.argtypes = [C.c_int, u64, u64p]
    lib.orc_ctl_fingerprint.argtypes = [C.c_int, C.c_int, C.c_int, u64, u64p]
    lib.orc_num_ctls.argtypes = [C.c_int]
    lib.orc_lookup_fingerprint.argtypes = [C.c_int, C.c_int, u64, u64p]
    lib.orc_gen_poseidon_rows.argtypes = [u64p, u64p, C.c_size_t, u64p]
    lib.orc_table_constraint_values.restype = C.c_long
    lib.orc_table_constraint_values.argtypes = [C.c_int, u64, u64p, C.c_size_t]
    lib.orc_stage_table.restype = C.c_long
    lib.orc_stage_table.argtypes = [C.c_int, C.c_uint32, C.POINTER(u64p), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), u64p, u64p, u64p, u64p,
                                    C.POINTER(C.c_uint32), u64p, u64p]
    lib.orc_hash_pages.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.orc_poseidon_bytes.argtypes = [C.c_char_p, C.c_size_t, u64p]
    lib.orc_const_hash_page.argtypes = [C.c_int, C.c_void_p]
    lib.orc_pagetree_create.restype = C.c_void_p
    lib.orc_pagetree_destroy.argtypes = [C.c_void_p]
    lib.orc_pagetree_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.orc_pagetree_page.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.orc_pagetree_count.argtypes = [C.c_void_p]
    lib.orc_pagetree_count.restype = C.c_size_t
    _lib = lib
    return lib


STANDARD_FAST_CONFIG = (2, 4, 16, 37, 2, 4, 5)     # reference prover/src/config.rs:17-29


def tables_arg(traces):
    """traces: list of (ncols, n) uint64 arrays -> (void* tables, ncols[], log_n[], keepalive)."""
    T = len(traces)
    per = [col_ptrs(t) for t in traces]
    arr = (C.c_void_p * T)(*[C.cast(p, C.c_void_p) for p in per])
    ncols = (C.c_uint32 * T)(*[t.shape[0] for t in traces])
    logn = (C.c_uint32 * T)(*[t.shape[1].bit_length() - 1 for t in traces])
    return arr, ncols, logn, (per, traces)


def prove_system(lib, system_id, traces, roots_before=None, roots_after=None, userdata=bytes(32), cfg=STANDARD_FAST_CONFIG):
    """Runs the oracle prover; returns the proof as a uint64 array."""
    arr, ncols, logn, keep = tables_arg(traces)
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    cw = (C.c_uint32 * 7)(*cfg)
    words = C.c_size_t()
    ptr = lib.orc_prove_system(system_id, C.cast(arr, C.c_void_p), ncols, logn, rb, ra, userdata, len(userdata), cw, C.byref(words))
    if not ptr:
        raise RuntimeError(lib.orc_last_error().decode())
    out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(words.value,)).copy()
    lib.orc_free(ptr)
    return out


def verify_system(lib, system_id, proof: np.ndarray, cfg=STANDARD_FAST_CONFIG):
    """Returns None if the oracle verifier accepts, else the rejection message."""
    cw = (C.c_uint32 * 7)(*cfg)
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    rc = lib.orc_verify_system(system_id, u64ptr(proof), proof.size, cw)
    return None if rc == 0 else lib.orc_last_error().decode()


def u64ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def col_ptrs(cols: np.ndarray):
    ncols, n = cols.shape
    arr = (C.POINTER(C.c_uint64) * ncols)()
    for i in range(ncols):
        arr[i] = C.cast(cols.ctypes.data + i * n * 8, C.POINTER(C.c_uint64))
    return arr


class OrcPageTree:
    """oracle/pagehash.h PageTree (emulator update_page_hash + compute_image_id restated on the CPU)."""

    def __init__(self, lib):
        self.lib, self.h = lib, lib.orc_pagetree_create()

    def split(self, indices, pages, registers: bytes, pc: int):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        pg = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1, 4096)
        image_id, root = np.zeros(32, dtype=np.uint8), np.zeros(32, dtype=np.uint8)
        rc = self.lib.orc_pagetree_split(self.h, idx.ctypes.data, pg.ctypes.data, idx.size, registers, pc, image_id.ctypes.data, root.ctypes.data)
        if rc != 0:
            raise RuntimeError("compute image ID fail")
        return bytes(image_id), bytes(root)

    def page(self, index: int):
        out = np.zeros(4096, dtype=np.uint8)
        return out if self.lib.orc_pagetree_page(self.h, index, out.ctypes.data) else None

    def count(self):
        return self.lib.orc_pagetree_count(self.h)

    def close(self):
        if self.h:
            self.lib.orc_pagetree_destroy(self.h)
            self.h = None


def stage_table(lib, system_id, table_index, cols, ctl_challenges, alphas, zeta, cfg=STANDARD_FAST_CONFIG, max_aux=256):
    """orc_stage_table -> (aux columns (num_aux, n), quotient coefficients (num_challenges, 2n), openings words)."""
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    ncols, n = a.shape
    nc = cfg[4]
    cc = np.ascontiguousarray(ctl_challenges, dtype=np.uint64).reshape(-1)
    al = np.ascontiguousarray(alphas, dtype=np.uint64)
    ze = np.ascontiguousarray(zeta, dtype=np.uint64)
    aux = np.zeros(max_aux * n, dtype=np.uint64)
    quot = np.zeros(nc * 2 * n, dtype=np.uint64)
    opn = np.zeros(4 * (ncols + max_aux) + max_aux + 4 * nc + 8, dtype=np.uint64)
    naux = C.c_uint32()
    cw = (C.c_uint32 * 7)(*cfg)
    w = lib.orc_stage_table(system_id, table_index, col_ptrs(a), ncols, n.bit_length() - 1, cw, u64ptr(cc), u64ptr(al), u64ptr(ze), u64ptr(aux),
                            C.byref(naux), u64ptr(quot), u64ptr(opn))
    if w < 0:
        raise RuntimeError(lib.orc_last_error().decode())
    return aux[:naux.value * n].reshape(naux.value, n).copy(), quot.reshape(nc, 2 * n), opn[:w].copy()
