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
    refused = int([l for l in lines if l.startswith("hostile lengths refused: ")][0].split(": ")[1])
    assert refused >= 4 * 12                                 # at least the header counts and the first table's vec lengths
    err = _run(exe, "errors")
    assert "config 2 4 16 37 2 4 5" in err and "junk: bad proof magic" in err and "empty: null/empty table" in err


def _dump_traces(ops, rng):
    """The operations of tests/traces.py as the reference's TYPED `Traces` (witness/traces.rs:46-60: byte arrays, MemoryAddress
    triples), one field per word, in the order tests/cpp/host_mirror.cpp reads them.  Contexts and segments are randomised (the
    test program runs in context 0) and written back into `ops`, so that a swapped pair shows."""
    w = []
    le4 = lambda v: list(int(v).to_bytes(4, "little"))
    addr = lambda ctx, seg, virt: [int(ctx), int(seg), int(virt)]
    w.append(len(ops[0])); w += [int(x) for x in ops[0].ravel()]
    w.append(len(ops[10])); w += [int(x) for x in ops[10].ravel()]
    w.append(len(ops[11]))
    for ctx, seg, virt, ts, is_read, value, filt in ops[11].tolist():
        w += addr(ctx, seg, virt) + [ts, is_read, value, filt]
    w.append(len(ops[2])); w += [int(x) for x in ops[2].ravel()]
    for t in (3, 5):
        sponge = []
        w.append(len(ops[t]))
        for virts, ts, data, _ctx, _seg in ops[t]:
            ctx, seg = int(rng.integers(1, 1 << 20)), int(rng.integers(1, 8))
            sponge.append((virts, ts, data, ctx, seg))
            w.append(len(virts))
            for v in virts:
                w += addr(ctx, seg, v)
            w += [ts, len(bytes(data))] + list(bytes(data))
        ops[t] = sponge
    w.append(len(ops[4])); w += [int(x) for x in ops[4].ravel()]
    w.append(len(ops[6]))
    for *ins, ts in ops[6].tolist():
        w += sum((le4(v) for v in ins), []) + [ts]
    ops[7][:, 10:12] = rng.integers(1, 1 << 16, size=(len(ops[7]), 2))
    w.append(len(ops[7]))
    for rnd, i0, i1, i2, i3, v0, v1, v2, v3, out_virt, ctx, seg, ts in ops[7].tolist():
        w.append(4)
        for v in (v0, v1, v2, v3):
            w += addr(ctx, seg, v)
        w += [ts] + sum((le4(v) for v in (i0, i1, i2, i3)), []) + [rnd] + addr(ctx, seg, out_virt)
    ops[8][:, 12:14] = rng.integers(1, 1 << 16, size=(len(ops[8]), 2))
    w.append(len(ops[8]))
    for row in ops[8].tolist():
        w += sum((le4(v) for v in row[:10]), []) + [row[10]] + addr(row[13], row[12], row[11]) + [row[14]]
    ops[9][:, 81:85] = rng.integers(1, 1 << 16, size=(len(ops[9]), 4))
    w.append(len(ops[9]))
    for row in ops[9].tolist():
        hx, ws, virts, (w_ptr, w_seg, w_ctx, ctx, seg, ts) = row[:8], row[8:72], row[72:80], row[80:]
        w.append(9)
        for v in virts:
            w += addr(ctx, seg, v)
        w += addr(w_ctx, w_seg, w_ptr) + [ts] + sum((le4(v) for v in hx), []) + [64] + sum((le4(v) for v in ws), [])
    return np.array(w, dtype=np.uint64)


def test_cpp_traces_marshal_into_the_operation_logs(exe, orc, tmp_path):
    """op_logs(const Traces&) of zkm_b200.hpp (the compiled counterpart of shim/src/b200_ops.rs) = the logs the Python mirror
    feeds zkm_b200_prove_with_ops, on the 12-table test program (2 SHA-256 blocks, Keccak and Poseidon sponges)."""
    from zkm_b200 import lib as zl
    _tables, ops = tr.all_stark_valid_traces(orc, sha_blocks=2, return_ops=True)
    dump = _dump_traces(ops, np.random.default_rng(5))
    (tmp_path / "traces.bin").write_bytes(dump.tobytes())
    out = _run(exe, "oplogs", tmp_path / "traces.bin", tmp_path / "logs.bin")
    assert "short: sha compress sponge operation needs nine addresses" in out and "cpu: the Cpu rows must be padded" in out
    got, at = np.fromfile(tmp_path / "logs.bin", dtype=np.uint64), 0
    for t in range(12):
        n_ops, words = int(got[at]), int(got[at + 1])
        log = got[at + 2:at + 2 + words]
        at += 2 + words
        if t == 1:
            assert n_ops == 0 and words == 0
            continue
        want, want_n = zl.sponge_log(ops[t]) if t in (3, 5) else (ops[t].ravel(), len(ops[t]))
        assert n_ops == want_n > 0, t
        assert log.size == want.size and (log == want).all(), t
    assert at == got.size


@pytest.mark.gpu
def test_cpp_prove_with_traces_equals_the_python_mirror(exe, zkm, tmp_path):
    from zkm_b200 import lib as zl
    out = tmp_path / "proof.bin"
    txt = _run(exe, "prove", out, *HEIGHTS)
    assert "timing scopes, first scope: compute all trace commitments" in txt and "ragged: ragged table" in txt
    got = np.fromfile(out, dtype=np.uint64)
    want = zl.prove_with_traces(zkm, zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, HEIGHTS))
    assert got.size == want.size and (got == want).all()
