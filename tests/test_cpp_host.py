"""include/zkm_b200.hpp: the C++ host-side mirror of the reference's prover API (prove_with_traces, AllProof and the plonky2 proof
structs, StarkConfig, TimingTree, the serde wire format) over the C ABI -- compiled with g++ and driven as a program, the way a
C++ host would use it.  CPU: decoding an AllStark proof made by the oracle (typed rebuild -> re-encode must be the identity,
shapes, JSON).  GPU: prove_with_traces from C++ gives the same proof as the Python mirror."""
import json
import pathlib
import subprocess

import numpy as np
import pytest

import traces as tr
from oracle import binding

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEIGHTS = [16, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from zkm_b200 import build
    build.build(verbose=False)
    out = tmp_path_factory.mktemp("cpp") / "host_mirror"
    cuda = "/usr/local/cuda/lib64"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests/cpp/host_mirror.cpp"), "-o", str(out),
                    "-L", str(ROOT / "zkm_b200"), "-lzkm_b200", f"-Wl,-rpath,{ROOT / 'zkm_b200'}", f"-Wl,-rpath,{cuda}", f"-Wl,-rpath-link,{cuda}"],
                   check=True)
    return out


def _run(exe, *args):
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_decodes_an_all_stark_proof(exe, orc, tmp_path):
    from zkm_b200 import lib as zl
    lib = zl.load()
    traces = tr.all_stark_valid_traces(orc)                  # a valid 12-table trace (MIPS program with the hash precompiles)
    heights = [t.shape[1].bit_length() - 1 for t in traces]
    proof = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, traces)
    path = tmp_path / "proof.bin"
    proof.tofile(path)
    out = _run(exe, "decode", path, 1)
    lines = out.splitlines()
    assert lines[0] == "challenges 2 userdata 32"
    for t, lg in enumerate(heights):
        f = lines[1 + t].split()
        assert f[:4] == ["table", str(t), "degree_bits", str(lg)], lines[1 + t]
        assert f[5:8] == ["16", "16", "16"] and f[f.index("local") + 1] == str(zl.NCOLS_ALL_STARK[t]) and f[f.index("queries") + 1] == "37"
    js = [l for l in lines if l.startswith("JSON ")][0][5:]
    assert js == zl.proof_table_json(lib, proof, 1) and len(json.loads(js)["trace_cap"]) == 16
    pv = json.loads([l for l in lines if l.startswith("PV ")][0][3:])
    assert pv["userdata"] == [0] * 32
    assert "truncated: proof buffer truncated" in out
    err = _run(exe, "errors")
    assert "config 2 4 16 37 2 4 5" in err and "junk: bad proof magic" in err and "empty: null/empty table" in err


@pytest.mark.gpu
def test_cpp_prove_with_traces_equals_the_python_mirror(exe, zkm, tmp_path):
    from zkm_b200 import lib as zl
    out = tmp_path / "proof.bin"
    txt = _run(exe, "prove", out, *HEIGHTS)
    assert "timing scopes, first scope: compute all trace commitments" in txt and "ragged: ragged table" in txt
    got = np.fromfile(out, dtype=np.uint64)
    want = zl.prove_with_traces(zkm, zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, HEIGHTS))
    assert got.size == want.size and (got == want).all()
