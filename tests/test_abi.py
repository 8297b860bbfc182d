"""CPU test: the product library loads and exports every symbol include/zkm_b200.h declares; with
no GPU it must fail loudly (no CPU fallback)."""
import ctypes as C
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _built():
    from zkm_b200 import build
    return build.build(verbose=False)


def test_header_symbols_exported():
    lib = C.CDLL(str(_built()))
    header = (ROOT / "include/zkm_b200.h").read_text()
    names = set(re.findall(r"\b(zkm_b200_\w+)\s*\(", header))
    assert len(names) >= 16
    from zkm_b200.lib import EXPORTS
    assert names == set(EXPORTS), names ^ set(EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_standard_fast_config_matches_reference():
    # reference prover/src/config.rs:17-29
    from zkm_b200.lib import load, StarkConfig
    lib = load()
    c = StarkConfig()
    lib.zkm_b200_standard_fast_config(C.byref(c))
    assert (c.rate_bits, c.cap_height, c.pow_bits, c.num_queries, c.num_challenges, c.arity_bits, c.final_poly_bits) == \
        (2, 4, 16, 37, 2, 4, 5)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zkm_b200.lib import load
    lib = load()
    err = C.c_void_p()
    rc = lib.zkm_b200_init(0, C.byref(err))
    assert rc == -1 and err.value
    msg = C.cast(err, C.c_char_p).value.decode()
    assert "no CPU fallback" in msg or "CUDA" in msg
    lib.zkm_b200_free_string(err)


def test_transcript_permutation_matches_oracle(orc):
    """The host-side Fiat-Shamir permutation of the prover (fast CPU schedule, csrc/poseidon_host.h) against the oracle's
    naive schedule and the known answers of SURVEY Appendix D.  Needs no GPU."""
    import numpy as np
    from oracle import binding
    from zkm_b200.lib import load, u64ptr
    from conftest import splitmix64_stream, P
    lib = load()
    n = 4096
    st = splitmix64_stream(0xC0FFEE, 12 * n).reshape(n, 12).copy()
    st[0] = 0
    st[1] = np.arange(12, dtype=np.uint64)
    st[2] = np.uint64(P - 1)
    st[3, ::2] = np.uint64(P - 1)
    want = st.copy()
    for i in range(n):
        row = np.ascontiguousarray(want[i])
        orc.orc_poseidon_permute(binding.u64ptr(row), 0)
        want[i] = row
    got = np.ascontiguousarray(st)
    err = C.c_void_p()
    assert lib.zkm_b200_transcript_permute(u64ptr(got), n, C.byref(err)) == 0
    assert (got == want).all()
    assert int(got[0, 0]) == 0x3c18a9786cb0b359 and int(got[1, 0]) == 0xd64e1e3efc5b8e9e


def test_column_layout_handshake_against_reference_derived_fixture():
    """tests/golden/column_layout_v1.json holds the CPU-table field offsets derived from the REFERENCE's struct declarations
    (tools/gen_layout_golden.py, run where /root/reference exists) as the (key, value) pairs of the layout handshake
    (include/zkm_b200.h): the constants the kernels were compiled with must agree with every one of them, and a shifted
    field must be reported by key.  This is the call the Rust shim makes with the offsets of ITS compiler before proving."""
    import json
    from zkm_b200.lib import load
    lib = load()
    g = json.loads((ROOT / "tests/golden/column_layout_v1.json").read_text())
    pairs = g["pairs"]
    assert len(pairs) >= 60
    flat = [x for p in pairs for x in p]
    arr = (C.c_uint32 * len(flat))(*flat)
    err = C.c_void_p()
    assert lib.zkm_b200_layout_check(arr, len(pairs), C.byref(err)) == 0, C.cast(err, C.c_char_p).value
    # the library's own description covers exactly the fixture's keys with the same values
    n = C.c_size_t()
    out = (C.c_uint32 * 512)()
    assert lib.zkm_b200_layout_describe(out, 256, C.byref(n), C.byref(err)) == 0
    mine = {out[2 * i]: out[2 * i + 1] for i in range(n.value)}
    assert mine == {k: v for k, v in pairs}
    # a rustc that had reordered CpuMiscView (rd_index <-> auxs) must be caught
    k = g["keys"]["ZKM_LK_CPU_G_MISC_AUXS_REL"]
    bad = [(a, b + 1 if a == k else b) for a, b in pairs]
    flat = [x for p in bad for x in p]
    arr = (C.c_uint32 * len(flat))(*flat)
    assert lib.zkm_b200_layout_check(arr, len(bad), C.byref(err)) == -1
    msg = C.cast(err, C.c_char_p).value.decode()
    lib.zkm_b200_free_string(err)
    assert f"key {k}" in msg and "mismatch" in msg
    unknown = (C.c_uint32 * 2)(9999, 1)
    assert lib.zkm_b200_layout_check(unknown, 1, C.byref(err)) == -1
    lib.zkm_b200_free_string(err)
