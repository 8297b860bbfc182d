"""ctypes binding of include/zkm_b200.h."""
import ctypes as C
import os
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
# ZKM_B200_LIB_TAG=x selects an A/B tuning build (zkm_b200/build.py ZKM_BUILD_TAG); unset = the product library
_TAG = os.environ.get("ZKM_B200_LIB_TAG", "")
LIB_PATH = _HERE / (f"libzkm_b200_{_TAG}.so" if _TAG else "libzkm_b200.so")

P = 0xFFFFFFFF00000001


class ZkmError(RuntimeError):
    pass


class Table(C.Structure):
    _fields_ = [("cols", C.POINTER(C.POINTER(C.c_uint64))), ("ncols", C.c_uint32), ("log_n", C.c_uint32)]


class StarkConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("rate_bits", "cap_height", "pow_bits", "num_queries", "num_challenges", "arity_bits", "final_poly_bits")]


_lib = None

# every symbol include/zkm_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "zkm_b200_free_string", "zkm_b200_free", "zkm_b200_standard_fast_config", "zkm_b200_init", "zkm_b200_shutdown",
    "zkm_b200_launch_count", "zkm_b200_sync", "zkm_b200_commit_values", "zkm_b200_commit_coeffs",
    "zkm_b200_commit_values_device", "zkm_b200_batch_free", "zkm_b200_batch_get_coeffs", "zkm_b200_batch_get_lde",
    "zkm_b200_batch_open", "zkm_b200_ntt", "zkm_b200_poseidon_permute", "zkm_b200_transcript_permute",
    "zkm_b200_worker_create", "zkm_b200_worker_bind", "zkm_b200_worker_destroy",
    "zkm_b200_shard_unique_id", "zkm_b200_shard_init", "zkm_b200_shard_shutdown",
    "zkm_b200_prove_with_traces", "zkm_b200_prove_with_trace_rows", "zkm_b200_memory_trace", "zkm_b200_prove_with_memory_ops", "zkm_b200_prove_with_ops", "zkm_b200_table_from_ops", "zkm_b200_prove_system", "zkm_b200_prove_system_device", "zkm_b200_synth_columns_device", "zkm_b200_synth_trace_device", "zkm_b200_synth_trace", "zkm_b200_system_shape", "zkm_b200_timer_start", "zkm_b200_timer_stop", "zkm_b200_profile_enable", "zkm_b200_profile_reset", "zkm_b200_profile_get", "zkm_b200_profile_get_traffic", "zkm_b200_timing_enable", "zkm_b200_last_timing", "zkm_b200_layout_check", "zkm_b200_layout_describe", "zkm_b200_proof_table_json", "zkm_b200_public_values_json", "zkm_b200_profile_families",
    "zkm_b200_stage_table", "zkm_b200_segment_json", "zkm_b200_hash_pages", "zkm_b200_pagetree_create", "zkm_b200_pagetree_destroy", "zkm_b200_pagetree_split", "zkm_b200_pagetree_page", "zkm_b200_pagetree_set_page",
    "zkm_b200_splitter_create", "zkm_b200_splitter_destroy", "zkm_b200_splitter_pagetree", "zkm_b200_splitter_segment_count", "zkm_b200_splitter_split",
]


def load():
    """Loads libzkm_b200.so; raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ZkmError(f"{LIB_PATH} not built: run `python -m zkm_b200.build` (nvcc, sm_100a). No CPU fallback exists.")
    lib = C.CDLL(str(LIB_PATH))
    u64p = C.POINTER(C.c_uint64)
    errp = C.POINTER(C.c_char_p)
    lib.zkm_b200_free_string.argtypes = [C.c_void_p]
    lib.zkm_b200_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.zkm_b200_shutdown.argtypes = [C.POINTER(C.c_void_p)]
    lib.zkm_b200_sync.argtypes = [C.POINTER(C.c_void_p)]
    lib.zkm_b200_launch_count.restype = C.c_uint64
    lib.zkm_b200_standard_fast_config.argtypes = [C.POINTER(StarkConfig)]
    for name in ("zkm_b200_commit_values", "zkm_b200_commit_coeffs"):
        getattr(lib, name).argtypes = [C.POINTER(Table), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), u64p,
                                       C.POINTER(C.c_void_p)]
    lib.zkm_b200_commit_values_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                  C.POINTER(C.c_void_p), u64p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_batch_free.argtypes = [C.c_void_p]
    lib.zkm_b200_batch_get_coeffs.argtypes = [C.c_void_p, C.c_uint32, u64p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_batch_get_lde.argtypes = [C.c_void_p, C.c_uint32, u64p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_batch_open.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_ntt.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    lib.zkm_b200_poseidon_permute.argtypes = [u64p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.zkm_b200_transcript_permute.argtypes = [u64p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.zkm_b200_synth_columns_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
    u32p = C.POINTER(C.c_uint32)
    lib.zkm_b200_prove_with_traces.argtypes = [C.POINTER(Table), u32p, u32p, C.c_char_p, C.c_uint32, C.POINTER(StarkConfig),
                                               C.POINTER(u64p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.zkm_b200_prove_system.argtypes = [C.c_int, C.POINTER(Table), C.c_uint32, u32p, u32p, C.c_char_p, C.c_uint32,
                                          C.POINTER(StarkConfig), C.POINTER(u64p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.zkm_b200_prove_system_device.argtypes = [C.c_int, C.POINTER(Table), C.POINTER(C.c_void_p), C.c_uint32, u32p, u32p, C.c_char_p,
                                                 C.c_uint32, C.POINTER(StarkConfig), C.POINTER(u64p), C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_void_p)]
    lib.zkm_b200_free.argtypes = [C.c_void_p]
    lib.zkm_b200_worker_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.zkm_b200_worker_bind.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_worker_destroy.argtypes = [C.c_void_p]
    lib.zkm_b200_worker_destroy.restype = None
    lib.zkm_b200_synth_trace_device.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_synth_trace.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, u64p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_system_shape.argtypes = [C.c_int, u32p, u32p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.zkm_b200_timer_start.argtypes = [C.POINTER(C.c_void_p)]
    lib.zkm_b200_timer_stop.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
    lib.zkm_b200_profile_enable.argtypes = [C.c_int]
    lib.zkm_b200_profile_reset.argtypes = [C.POINTER(C.c_void_p)]
    lib.zkm_b200_profile_get.argtypes = [C.c_char_p, C.POINTER(C.c_double), u64p, C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
    lib.zkm_b200_shard_unique_id.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_shard_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.zkm_b200_shard_shutdown.argtypes = [C.POINTER(C.c_void_p)]
    lib.zkm_b200_profile_get_traffic.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
    lib.zkm_b200_profile_families.restype = C.c_void_p
    lib.zkm_b200_layout_check.argtypes = [C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_void_p)]
    lib.zkm_b200_layout_describe.argtypes = [C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.zkm_b200_proof_table_json.argtypes = [u64p, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.zkm_b200_public_values_json.argtypes = [u64p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.zkm_b200_timing_enable.argtypes = [C.c_int]
    lib.zkm_b200_last_timing.restype = C.c_void_p
    del errp
    _lib = lib
    return lib


def check(lib, rc, err):
    if rc != 0:
        msg = C.cast(err, C.c_char_p).value if err.value else b"unknown error"
        text = msg.decode(errors="replace")
        if err.value:
            lib.zkm_b200_free_string(err)
        raise ZkmError(text)


def u64ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def make_table(cols: np.ndarray):
    """cols: (ncols, n) uint64 C-contiguous -> (Table, keepalive)."""
    assert cols.dtype == np.uint64 and cols.ndim == 2 and cols.flags["C_CONTIGUOUS"]
    ncols, n = cols.shape
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    ptrs = (C.POINTER(C.c_uint64) * ncols)()
    base = cols.ctypes.data
    for i in range(ncols):
        ptrs[i] = C.cast(base + i * n * 8, C.POINTER(C.c_uint64))
    t = Table(C.cast(ptrs, C.POINTER(C.POINTER(C.c_uint64))), ncols, log_n)
    return t, (ptrs, cols)


_inited = False


def init(device: int = 0):
    global _inited
    lib = load()
    if not _inited:
        err = C.c_void_p()
        check(lib, lib.zkm_b200_init(device, C.byref(err)), err)
        _inited = True
    return lib


def standard_fast_config(lib=None):
    lib = lib or load()
    c = StarkConfig()
    lib.zkm_b200_standard_fast_config(C.byref(c))
    return c


def prove_system(lib, system_id, traces, roots_before=None, roots_after=None, userdata=bytes(32), cfg=None):
    """traces: list of (ncols, n) uint64 arrays (host).  Returns the proof buffer as a uint64 array."""
    cfg = cfg or standard_fast_config(lib)
    T = len(traces)
    made = [make_table(np.ascontiguousarray(t)) for t in traces]
    arr = (Table * T)(*[m[0] for m in made])
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    out = C.POINTER(C.c_uint64)()
    words = C.c_size_t()
    err = C.c_void_p()
    rc = lib.zkm_b200_prove_system(system_id, arr, T, rb, ra, userdata, len(userdata), C.byref(cfg), C.byref(out), C.byref(words),
                                   C.byref(err))
    check(lib, rc, err)
    proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
    lib.zkm_b200_free(out)
    return proof


class Worker:
    """A worker context bound to the calling thread for the duration of a `with` block (include/zkm_b200.h)."""

    def __init__(self, lib):
        self.lib = lib
        self.h = C.c_void_p()
        err = C.c_void_p()
        check(lib, lib.zkm_b200_worker_create(C.byref(self.h), C.byref(err)), err)

    def __enter__(self):
        err = C.c_void_p()
        check(self.lib, self.lib.zkm_b200_worker_bind(self.h, C.byref(err)), err)
        return self

    def __exit__(self, *exc):
        err = C.c_void_p()
        self.lib.zkm_b200_worker_bind(None, C.byref(err))
        return False

    def close(self):
        if self.h:
            self.lib.zkm_b200_worker_destroy(self.h)
            self.h = C.c_void_p()


def prove_with_traces(lib, traces, roots_before=None, roots_after=None, userdata=bytes(32), cfg=None):
    """The drop-in entry point (reference prover.rs:130-140 prove_with_traces): 12 host traces in `Table` order
    -> AllProof buffer."""
    cfg = cfg or standard_fast_config(lib)
    assert len(traces) == 12
    made = [make_table(np.ascontiguousarray(t)) for t in traces]
    arr = (Table * 12)(*[m[0] for m in made])
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    out = C.POINTER(C.c_uint64)()
    words = C.c_size_t()
    err = C.c_void_p()
    rc = lib.zkm_b200_prove_with_traces(arr, rb, ra, userdata, len(userdata), C.byref(cfg), C.byref(out), C.byref(words), C.byref(err))
    check(lib, rc, err)
    proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
    lib.zkm_b200_free(out)
    return proof


class TableRows(C.Structure):
    _fields_ = [("rows", C.POINTER(C.c_uint64)), ("ncols", C.c_uint32), ("log_n", C.c_uint32)]


def prove_with_trace_rows(lib, traces, as_rows, roots_before=None, roots_after=None, userdata=bytes(32), cfg=None):
    """zkm_b200_prove_with_trace_rows: traces[t] is (ncols, n) column-major, except for t in `as_rows`, which is passed as the
    (n, ncols) row-major block its generator produced."""
    cfg = cfg or standard_fast_config(lib)
    keep, cols, rows = [], [], []
    for t, a in enumerate(traces):
        if t in as_rows:
            r = np.ascontiguousarray(a, dtype=np.uint64)
            assert r.ndim == 2
            keep.append(r)
            cols.append(Table(None, r.shape[1], r.shape[0].bit_length() - 1))
            rows.append(TableRows(r.ctypes.data_as(C.POINTER(C.c_uint64)), r.shape[1], r.shape[0].bit_length() - 1))
        else:
            m = make_table(np.ascontiguousarray(a))
            keep.append(m)
            cols.append(m[0])
            rows.append(TableRows(None, 0, 0))
    carr, rarr = (Table * 12)(*cols), (TableRows * 12)(*rows)
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
    lib.zkm_b200_prove_with_trace_rows.argtypes = [C.POINTER(Table), C.POINTER(TableRows), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                                   C.c_char_p, C.c_uint32, C.POINTER(StarkConfig), C.POINTER(C.POINTER(C.c_uint64)),
                                                   C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    rc = lib.zkm_b200_prove_with_trace_rows(carr, rarr, rb, ra, userdata, len(userdata), C.byref(cfg), C.byref(out), C.byref(words),
                                            C.byref(err))
    check(lib, rc, err)
    proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
    lib.zkm_b200_free(out)
    return proof


def prove_with_memory_ops(lib, traces, memory_ops, roots_before=None, roots_after=None, userdata=bytes(32), cfg=None):
    """zkm_b200_prove_with_memory_ops: traces[0..10] column-major host tables (traces[11] is ignored), Memory from the log."""
    cfg = cfg or standard_fast_config(lib)
    made = [make_table(np.ascontiguousarray(t)) for t in traces[:11]]
    arr = (Table * 12)(*([m[0] for m in made] + [Table(None, 0, 0)]))
    ops = np.ascontiguousarray(memory_ops, dtype=np.uint64)
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
    lib.zkm_b200_prove_with_memory_ops.argtypes = [C.POINTER(Table), C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_uint32),
                                                   C.POINTER(C.c_uint32), C.c_char_p, C.c_uint32, C.POINTER(StarkConfig),
                                                   C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    rc = lib.zkm_b200_prove_with_memory_ops(arr, None, ops.ctypes.data_as(C.POINTER(C.c_uint64)), ops.shape[0], rb, ra, userdata,
                                            len(userdata), C.byref(cfg), C.byref(out), C.byref(words), C.byref(err))
    check(lib, rc, err)
    proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
    lib.zkm_b200_free(out)
    return proof


def memory_trace(lib, ops):
    """zkm_b200_memory_trace: ops = (n_ops, 7) uint64 (context, segment, virt, timestamp, is_read, value, filter) in push
    order -> (13, n) Memory table."""
    a = np.ascontiguousarray(ops, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 7
    out, lg, err = C.POINTER(C.c_uint64)(), C.c_uint32(), C.c_void_p()
    lib.zkm_b200_memory_trace.argtypes = [C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_void_p)]
    check(lib, lib.zkm_b200_memory_trace(a.ctypes.data_as(C.POINTER(C.c_uint64)), a.shape[0], C.byref(out), C.byref(lg), C.byref(err)), err)
    n = 1 << lg.value
    t = np.ctypeslib.as_array(out, shape=(13 * n,)).copy().reshape(13, n)
    lib.zkm_b200_free(out)
    return t


def system_shape(lib, system_id):
    n = C.c_uint32()
    nc = (C.c_uint32 * 32)()
    err = C.c_void_p()
    check(lib, lib.zkm_b200_system_shape(system_id, C.byref(n), nc, 32, C.byref(err)), err)
    return [int(nc[i]) for i in range(n.value)]


def synth_traces(lib, system_id, log_heights, seed=0x5EED000000000000):
    """Host copies of the synthetic traces (generated on the device) for every table of the System."""
    out = []
    for t, (nc, lg) in enumerate(zip(system_shape(lib, system_id), log_heights)):
        a = np.zeros((nc, 1 << lg), dtype=np.uint64)
        err = C.c_void_p()
        check(lib, lib.zkm_b200_synth_trace(system_id, t, lg, seed | (t << 16), u64ptr(a), C.byref(err)), err)
        out.append(a)
    return out


def proof_table_json(lib, proof: np.ndarray, table: int) -> str:
    """serde_json::to_string(&all_proof.stark_proofs[table].proof) of the reference (include/zkm_b200.h "proof wire format")."""
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
    check(lib, lib.zkm_b200_proof_table_json(u64ptr(proof), proof.size, table, C.byref(out), C.byref(n), C.byref(err)), err)
    s = C.string_at(out, n.value).decode()
    lib.zkm_b200_free_string(out)
    return s


def public_values_json(lib, proof: np.ndarray) -> str:
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
    check(lib, lib.zkm_b200_public_values_json(u64ptr(proof), proof.size, C.byref(out), C.byref(n), C.byref(err)), err)
    s = C.string_at(out, n.value).decode()
    lib.zkm_b200_free_string(out)
    return s


class OpLog(C.Structure):
    _fields_ = [("ops", C.POINTER(C.c_uint64)), ("n_ops", C.c_size_t)]


NCOLS_ALL_STARK = [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]


def sponge_log(ops):
    """The variable-width log of the two byte sponges (tables 3 and 5, include/zkm_b200.h): ops = [(virt of every input word,
    timestamp, input bytes, context, segment)] -> (1-D uint64 log whose word 0 is its word count, number of operations)."""
    words = [0]
    for virts, ts, data, ctx, seg in ops:
        data = bytes(data)
        words += [ctx, seg, ts, len(data), len(virts)] + [int(v) for v in virts]
        padded = data + bytes(-len(data) % 8)
        words += [int.from_bytes(padded[i:i + 8], "little") for i in range(0, len(padded), 8)]
    words[0] = len(words)
    return np.array(words, dtype=np.uint64), len(ops)


def table_from_ops(lib, table: int, ops, min_rows: int = 64, n_ops=None):
    """zkm_b200_table_from_ops: ops = (n_ops, words_per_op) uint64 -> (ncols, n) table generated on the device; for the two
    byte sponges ops is the 1-D log of sponge_log() and n_ops its number of operations."""
    a = np.ascontiguousarray(ops, dtype=np.uint64)
    assert a.ndim == 2 or n_ops is not None
    n_ops = a.shape[0] if n_ops is None else n_ops
    out, lg, err = C.POINTER(C.c_uint64)(), C.c_uint32(), C.c_void_p()
    lib.zkm_b200_table_from_ops.argtypes = [C.c_uint32, C.POINTER(C.c_uint64), C.c_size_t, C.c_uint32, C.POINTER(C.POINTER(C.c_uint64)),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)]
    check(lib, lib.zkm_b200_table_from_ops(table, a.ctypes.data_as(C.POINTER(C.c_uint64)), n_ops, min_rows, C.byref(out), C.byref(lg),
                                           C.byref(err)), err)
    n, nc = 1 << lg.value, NCOLS_ALL_STARK[table]
    t = np.ctypeslib.as_array(out, shape=(nc * n,)).copy().reshape(nc, n)
    lib.zkm_b200_free(out)
    return t


def prove_with_ops(lib, traces, op_logs, roots_before=None, roots_after=None, userdata=bytes(32), cfg=None):
    """zkm_b200_prove_with_ops: traces[t] column-major (ignored for tables with a log), op_logs = {table: (n_ops, k) uint64}."""
    cfg = cfg or standard_fast_config(lib)
    keep, cols = [], []
    for t, a in enumerate(traces):
        if t in op_logs:
            cols.append(Table(None, 0, 0))
        else:
            tb, k = make_table(np.ascontiguousarray(a, dtype=np.uint64))
            keep.append(k)
            cols.append(tb)
    arr = (Table * 12)(*cols)
    logs = (OpLog * 12)()
    for t, ops in op_logs.items():
        n_ops = None
        if isinstance(ops, tuple):                  # (1-D sponge log, number of operations)
            ops, n_ops = ops
        a = np.ascontiguousarray(ops, dtype=np.uint64)
        keep.append(a)
        logs[t].ops = a.ctypes.data_as(C.POINTER(C.c_uint64))
        logs[t].n_ops = a.shape[0] if n_ops is None else n_ops
    rb = (C.c_uint32 * 8)(*(roots_before or range(1, 9)))
    ra = (C.c_uint32 * 8)(*(roots_after or range(11, 19)))
    out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
    lib.zkm_b200_prove_with_ops.argtypes = [C.POINTER(Table), C.c_void_p, C.POINTER(OpLog), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.c_char_p, C.c_uint32, C.POINTER(StarkConfig), C.POINTER(C.POINTER(C.c_uint64)),
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    rc = lib.zkm_b200_prove_with_ops(arr, None, logs, rb, ra, userdata, len(userdata), C.byref(cfg), C.byref(out), C.byref(words), C.byref(err))
    check(lib, rc, err)
    proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
    lib.zkm_b200_free(out)
    return proof


def hash_pages(lib, pages):
    """zkm_b200_hash_pages: pages = (n, 4096) uint8 -> (n, 32) uint8 digests (emulator hash_page on the device)."""
    a = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1, 4096)
    out = np.zeros((a.shape[0], 32), dtype=np.uint8)
    err = C.c_void_p()
    lib.zkm_b200_hash_pages.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
    check(lib, lib.zkm_b200_hash_pages(a.ctypes.data, a.shape[0], out.ctypes.data, C.byref(err)), err)
    return out


class PageTree:
    """zkm_pagetree_t: the emulator's hash pages + update_page_hash / compute_image_id with the hashing on the device."""

    def __init__(self, lib):
        self.lib, self.h = lib, C.c_void_p()
        err = C.c_void_p()
        lib.zkm_b200_pagetree_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        lib.zkm_b200_pagetree_destroy.argtypes = [C.c_void_p]
        lib.zkm_b200_pagetree_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                                C.POINTER(C.c_void_p)]
        lib.zkm_b200_pagetree_page.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        lib.zkm_b200_pagetree_set_page.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]
        check(lib, lib.zkm_b200_pagetree_create(C.byref(self.h), C.byref(err)), err)

    def set_page(self, index: int, data):
        """Seeds one hash page (a state resumed from a segment file carries them in its memory image)."""
        pg = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        assert pg.size == 4096
        err = C.c_void_p()
        check(self.lib, self.lib.zkm_b200_pagetree_set_page(self.h, index, pg.ctypes.data, C.byref(err)), err)

    def split(self, indices, pages, registers: bytes, pc: int):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        pg = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1, 4096)
        assert idx.size == pg.shape[0] and len(registers) == 156
        image_id, root = np.zeros(32, dtype=np.uint8), np.zeros(32, dtype=np.uint8)
        err = C.c_void_p()
        check(self.lib, self.lib.zkm_b200_pagetree_split(self.h, idx.ctypes.data, pg.ctypes.data, idx.size, registers, pc, image_id.ctypes.data,
                                                         root.ctypes.data, C.byref(err)), err)
        return bytes(image_id), bytes(root)

    def page(self, index: int):
        out, present, err = np.zeros(4096, dtype=np.uint8), C.c_int(), C.c_void_p()
        check(self.lib, self.lib.zkm_b200_pagetree_page(self.h, index, out.ctypes.data, C.byref(present), C.byref(err)), err)
        return out if present.value else None

    def close(self):
        if self.h:
            self.lib.zkm_b200_pagetree_destroy(self.h)
            self.h = C.c_void_p()


def stage_table(lib, system_id: int, table_index: int, cols, ctl_challenges, alphas, zeta, cfg=None):
    """zkm_b200_stage_table -> (aux columns (num_aux, n), quotient coefficients (num_challenges, 2n), openings words)."""
    cfg = cfg or standard_fast_config(lib)
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    t, keep = make_table(a)
    cc = np.ascontiguousarray(ctl_challenges, dtype=np.uint64).reshape(-1)
    al = np.ascontiguousarray(alphas, dtype=np.uint64)
    ze = np.ascontiguousarray(zeta, dtype=np.uint64)
    aux, quot, opn = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
    naux, words, err = C.c_uint32(), C.c_size_t(), C.c_void_p()
    u64p = C.POINTER(C.c_uint64)
    lib.zkm_b200_stage_table.argtypes = [C.c_int, C.c_uint32, C.POINTER(Table), C.POINTER(StarkConfig), u64p, u64p, u64p, C.POINTER(u64p),
                                         C.POINTER(C.c_uint32), C.POINTER(u64p), C.POINTER(u64p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    check(lib, lib.zkm_b200_stage_table(system_id, table_index, C.byref(t), C.byref(cfg), u64ptr(cc), u64ptr(al), u64ptr(ze), C.byref(aux),
                                        C.byref(naux), C.byref(quot), C.byref(opn), C.byref(words), C.byref(err)), err)
    n, nc = a.shape[1], cfg.num_challenges
    out = (np.ctypeslib.as_array(aux, shape=(naux.value * n,)).copy().reshape(naux.value, n),
           np.ctypeslib.as_array(quot, shape=(nc * 2 * n,)).copy().reshape(nc, 2 * n),
           np.ctypeslib.as_array(opn, shape=(words.value,)).copy())
    for p in (aux, quot, opn):
        lib.zkm_b200_free(p)
    return out


class SegmentC(C.Structure):
    _fields_ = [("page_indices", C.POINTER(C.c_uint32)), ("pages", C.c_void_p), ("n_pages", C.c_size_t), ("pc", C.c_uint32),
                ("segment_id", C.c_uint32), ("pre_image_id", C.c_uint8 * 32), ("pre_hash_root", C.c_uint8 * 32), ("image_id", C.c_uint8 * 32),
                ("page_hash_root", C.c_uint8 * 32), ("end_pc", C.c_uint32), ("step", C.c_uint64), ("input_stream", C.POINTER(C.c_char_p)),
                ("input_stream_lens", C.POINTER(C.c_size_t)), ("n_input_streams", C.c_size_t), ("input_stream_ptr", C.c_uint64),
                ("public_values_stream", C.c_char_p), ("public_values_stream_len", C.c_size_t), ("public_values_stream_ptr", C.c_uint64)]


def segment_json(lib, page_indices, pages, pc, segment_id, pre_image_id, pre_hash_root, image_id, page_hash_root, end_pc, step, input_stream,
                 input_stream_ptr, public_values_stream, public_values_stream_ptr) -> bytes:
    """zkm_b200_segment_json: the emulator's Segment file (serde_json).  Host-only."""
    idx = np.ascontiguousarray(page_indices, dtype=np.uint32)
    pg = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1, 4096)
    s = SegmentC()
    s.page_indices, s.pages, s.n_pages = idx.ctypes.data_as(C.POINTER(C.c_uint32)), pg.ctypes.data, idx.size
    s.pc, s.segment_id, s.end_pc, s.step = pc, segment_id, end_pc, step
    for name, val in (("pre_image_id", pre_image_id), ("pre_hash_root", pre_hash_root), ("image_id", image_id), ("page_hash_root", page_hash_root)):
        setattr(s, name, (C.c_uint8 * 32)(*val))
    streams = [bytes(b) for b in input_stream]
    arr = (C.c_char_p * max(1, len(streams)))(*streams)
    lens = (C.c_size_t * max(1, len(streams)))(*[len(b) for b in streams])
    s.input_stream, s.input_stream_lens, s.n_input_streams, s.input_stream_ptr = arr, lens, len(streams), input_stream_ptr
    pvs = bytes(public_values_stream)
    s.public_values_stream, s.public_values_stream_len, s.public_values_stream_ptr = pvs, len(pvs), public_values_stream_ptr
    out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
    lib.zkm_b200_segment_json.argtypes = [C.POINTER(SegmentC), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    check(lib, lib.zkm_b200_segment_json(C.byref(s), C.byref(out), C.byref(n), C.byref(err)), err)
    text = C.string_at(out, n.value)
    lib.zkm_b200_free_string(out)
    return text


class SplitState(C.Structure):
    _fields_ = [("dirty_page_indices", C.c_void_p), ("dirty_pages", C.c_void_p), ("n_dirty_pages", C.c_size_t),
                ("read_page_indices", C.c_void_p), ("read_pages", C.c_void_p), ("n_read_pages", C.c_size_t), ("registers", C.c_char_p),
                ("pc", C.c_uint32), ("step", C.c_uint64), ("input_stream", C.POINTER(C.c_char_p)), ("input_stream_lens", C.POINTER(C.c_size_t)),
                ("n_input_streams", C.c_size_t), ("input_stream_ptr", C.c_uint64), ("public_values_stream", C.c_char_p),
                ("public_values_stream_len", C.c_size_t), ("public_values_stream_ptr", C.c_uint64)]


class Splitter:
    """zkm_splitter_t: InstrumentedState::split_segment (hashing on the device, pre_* bookkeeping, segment file text)."""

    def __init__(self, lib):
        self.lib, self.h = lib, C.c_void_p()
        err = C.c_void_p()
        lib.zkm_b200_splitter_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        lib.zkm_b200_splitter_destroy.argtypes = [C.c_void_p]
        lib.zkm_b200_splitter_segment_count.argtypes = [C.c_void_p]
        lib.zkm_b200_splitter_segment_count.restype = C.c_uint32
        lib.zkm_b200_splitter_split.argtypes = [C.c_void_p, C.POINTER(SplitState), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p,
                                                C.c_void_p, C.POINTER(C.c_void_p)]
        check(lib, lib.zkm_b200_splitter_create(C.byref(self.h), C.byref(err)), err)

    def split(self, dirty, read, registers: bytes, pc: int, step: int, input_stream, input_stream_ptr, public_values_stream, public_values_stream_ptr,
              proof: bool):
        """dirty / read = (ascending page indices, (n, 4096) uint8).  Returns (segment json bytes or None, image id, page hash root)."""
        di, dp = np.ascontiguousarray(dirty[0], dtype=np.uint32), np.ascontiguousarray(dirty[1], dtype=np.uint8).reshape(-1, 4096)
        ri, rp = np.ascontiguousarray(read[0], dtype=np.uint32), np.ascontiguousarray(read[1], dtype=np.uint8).reshape(-1, 4096)
        st = SplitState()
        st.dirty_page_indices, st.dirty_pages, st.n_dirty_pages = di.ctypes.data, dp.ctypes.data, di.size
        st.read_page_indices, st.read_pages, st.n_read_pages = ri.ctypes.data, rp.ctypes.data, ri.size
        st.registers, st.pc, st.step = registers, pc, step
        streams = [bytes(b) for b in input_stream]
        arr = (C.c_char_p * max(1, len(streams)))(*streams)
        lens = (C.c_size_t * max(1, len(streams)))(*[len(b) for b in streams])
        st.input_stream, st.input_stream_lens, st.n_input_streams, st.input_stream_ptr = arr, lens, len(streams), input_stream_ptr
        pvs = bytes(public_values_stream)
        st.public_values_stream, st.public_values_stream_len, st.public_values_stream_ptr = pvs, len(pvs), public_values_stream_ptr
        out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
        image_id, root = np.zeros(32, dtype=np.uint8), np.zeros(32, dtype=np.uint8)
        check(self.lib, self.lib.zkm_b200_splitter_split(self.h, C.byref(st), int(proof), C.byref(out), C.byref(n), image_id.ctypes.data,
                                                         root.ctypes.data, C.byref(err)), err)
        text = None
        if out.value:
            text = C.string_at(out, n.value)
            self.lib.zkm_b200_free_string(out)
        return text, bytes(image_id), bytes(root)

    def segment_count(self):
        return self.lib.zkm_b200_splitter_segment_count(self.h)

    def close(self):
        if self.h:
            self.lib.zkm_b200_splitter_destroy(self.h)
            self.h = C.c_void_p()
