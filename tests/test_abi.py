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


def test_header_is_plain_c_and_cxx():
    """include/zkm_b200.h is the boundary a cgo / Rust-bindgen / C caller binds: it must compile on its own as C99 and as C++11
    (pedantic, warnings as errors), and a C program using the host-only entry points links against the library."""
    import subprocess
    hdr = str(ROOT / "include/zkm_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)


def test_c_program_links_and_uses_the_host_only_entry_points(tmp_path):
    import subprocess
    lib_path = _built()
    src = tmp_path / "c_user.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "zkm_b200.h"
int main(void) {
    zkm_stark_config_t cfg;
    zkm_b200_standard_fast_config(&cfg);
    printf("config %u %u %u %u\n", cfg.rate_bits, cfg.cap_height, cfg.pow_bits, cfg.num_queries);
    char* err = NULL;
    zkm_pagetree_t* t = NULL;
    if (zkm_b200_pagetree_create(&t, &err)) return 1;
    unsigned char page[4096], back[4096];
    memset(page, 0x5A, sizeof page);
    if (zkm_b200_pagetree_set_page(t, 0x81020u, page, &err)) return 2;
    int present = 0;
    if (zkm_b200_pagetree_page(t, 0x81020u, back, &present, &err) || !present || memcmp(page, back, sizeof page)) return 3;
    if (zkm_b200_pagetree_set_page(t, 5u, page, &err) == 0) return 4;
    printf("error: %s\n", err);
    zkm_b200_free_string(err);
    zkm_b200_pagetree_destroy(t);
    uint64_t junk[4] = {1, 2, 3, 4};
    char* json = NULL; size_t len = 0;
    if (zkm_b200_public_values_json(junk, 4, &json, &len, &err) == 0) return 5;
    printf("error: %s\n", err);
    zkm_b200_free_string(err);
    return 0;
}
''')
    exe = tmp_path / "c_user"
    cuda = "/usr/local/cuda/lib64"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(lib_path.parent), "-lzkm_b200",
                    f"-Wl,-rpath,{lib_path.parent}", f"-Wl,-rpath,{cuda}", f"-Wl,-rpath-link,{cuda}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.splitlines() == ["config 2 4 16 37", "error: not a hash page index", "error: bad proof magic"]
