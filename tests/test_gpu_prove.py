"""GPU parity tests for the full prover (through the C ABI): the proof of the B200 path must equal the
CPU oracle's proof word for word on the same traces, and the oracle's verifier (reference verifier.rs
logic) must accept it.  Systems: zkm_b200/csrc/tables/systems.h."""
import ctypes as C

import numpy as np
import pytest

import traces as tr
from oracle import binding
from zkm_b200 import lib as zl

pytestmark = pytest.mark.gpu


def _traces(orc, sid, scale=0):
    if sid == tr.SYSTEM_LOGIC:
        return [tr.logic_trace(6 + scale)]
    if sid == tr.SYSTEM_POSEIDON:
        return [tr.poseidon_trace(orc, 6 + scale)]
    if sid == tr.SYSTEM_MEMORY:
        return [tr.memory_trace(7 + scale)]
    if sid == tr.SYSTEM_ARITH:
        return [tr.arithmetic_trace()]
    if sid == tr.SYSTEM_KECCAK:
        return tr.keccak_system_traces()
    if sid == tr.SYSTEM_POSEIDON_SPONGE:
        return tr.poseidon_system_traces(orc)
    if sid == tr.SYSTEM_SHA_EXTEND:
        return tr.sha_extend_system_traces()
    if sid == tr.SYSTEM_SHA_COMPRESS:
        return tr.sha_compress_system_traces()
    if sid == tr.SYSTEM_CPU:
        return tr.cpu_system_traces()
    return [tr.poseidon_trace(orc, 6 + scale), tr.logic_trace(8 + scale), tr.memory_trace(7 + scale)]


def _first_diff(a, b):
    if a.size != b.size:
        return f"length {a.size} vs {b.size}"
    d = np.nonzero(a != b)[0]
    return None if d.size == 0 else f"first differing word {d[0]} of {a.size} ({d.size} differ)"


@pytest.mark.parametrize("sid", [tr.SYSTEM_LOGIC, tr.SYSTEM_POSEIDON, tr.SYSTEM_MEMORY, tr.SYSTEM_MINI3, tr.SYSTEM_ARITH, tr.SYSTEM_KECCAK, tr.SYSTEM_POSEIDON_SPONGE,
                                 tr.SYSTEM_SHA_EXTEND, tr.SYSTEM_SHA_COMPRESS, tr.SYSTEM_CPU])
def test_gpu_proof_equals_oracle_proof_and_verifies(zkm, orc, sid):
    traces = _traces(orc, sid)
    gpu = zl.prove_system(zkm, sid, traces)
    assert binding.verify_system(orc, sid, gpu) is None
    cpu = binding.prove_system(orc, sid, traces)
    assert _first_diff(gpu, cpu) is None


@pytest.mark.parametrize("sid,scale", [(tr.SYSTEM_LOGIC, 7), (tr.SYSTEM_MINI3, 4), (tr.SYSTEM_MEMORY, 9)])
def test_gpu_proof_larger_sizes(zkm, orc, sid, scale):
    """Sizes that exercise the two-pass NTT and several FRI rounds (degree_bits 13..16)."""
    traces = _traces(orc, sid, scale)
    gpu = zl.prove_system(zkm, sid, traces)
    assert binding.verify_system(orc, sid, gpu) is None
    cpu = binding.prove_system(orc, sid, traces)
    assert _first_diff(gpu, cpu) is None


def test_gpu_proof_other_public_values_and_config(zkm, orc):
    traces = _traces(orc, tr.SYSTEM_LOGIC)
    cfg = zl.standard_fast_config(zkm)
    cfg.num_queries = 11
    cfg.pow_bits = 10
    ud = bytes(range(32))
    gpu = zl.prove_system(zkm, tr.SYSTEM_LOGIC, traces, roots_before=[7] * 8, roots_after=[9] * 8, userdata=ud, cfg=cfg)
    cw = (2, 4, 10, 11, 2, 4, 5)
    assert binding.verify_system(orc, tr.SYSTEM_LOGIC, gpu, cfg=cw) is None
    cpu = binding.prove_system(orc, tr.SYSTEM_LOGIC, traces, roots_before=[7] * 8, roots_after=[9] * 8, userdata=ud, cfg=cw)
    assert _first_diff(gpu, cpu) is None


def test_gpu_invalid_trace_proof_is_rejected(zkm, orc):
    t = tr.logic_trace(6)
    t[68, 5] = (int(t[68, 5]) + 1) % tr.P
    gpu = zl.prove_system(zkm, tr.SYSTEM_LOGIC, [t])
    assert binding.verify_system(orc, tr.SYSTEM_LOGIC, gpu) is not None
    # still identical to what the reference algorithm produces on the same (invalid) trace
    assert _first_diff(gpu, binding.prove_system(orc, tr.SYSTEM_LOGIC, [t])) is None


def test_gpu_non_binary_filter_error(zkm):
    t = tr.logic_trace(6)
    t[0, 2] = 2
    with pytest.raises(zl.ZkmError, match="Non-binary filter"):
        zl.prove_system(zkm, tr.SYSTEM_LOGIC, [t])


def test_gpu_wrong_shape_errors(zkm):
    with pytest.raises(zl.ZkmError, match="wrong number of trace columns"):
        zl.prove_system(zkm, tr.SYSTEM_LOGIC, [np.zeros((5, 64), dtype=np.uint64)])
    with pytest.raises(zl.ZkmError, match="wrong number of tables"):
        zl.prove_system(zkm, tr.SYSTEM_MINI3, [tr.logic_trace(6)])


def test_all_stark_synthetic_proof_equals_oracle(zkm, orc):
    """The full 12-table AllStark (all_stark.rs) on synthetic traces at small heights:
    every table's constraint evaluator, the 15 CTLs and the Arithmetic/Memory logUp lookups run on the
    device and the proof must equal the oracle's word for word.  (Synthetic traces are not valid
    executions, so the proof is not expected to verify.)"""
    heights = [16, 7, 6, 6, 6, 6, 6, 7, 6, 6, 8, 9]
    traces = zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, heights)
    assert [t.shape[0] for t in traces] == [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]
    gpu = zl.prove_system(zkm, tr.SYSTEM_ALL_STARK, traces)
    cpu = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, traces)
    assert _first_diff(gpu, cpu) is None


def test_all_stark_synthetic_proof_equals_oracle_mid_heights(zkm, orc):
    """Same as above with the Cpu, Logic and Memory tables at 2^13 rows and Arithmetic at 2^16: these sizes take the
    one-thread-per-point quotient kernels, the two-pass NTT and three FRI rounds (the 2^6-row tables take the
    cooperative small-table kernels), i.e. the code paths the 2^20-row benchmark runs."""
    heights = [16, 13, 6, 6, 6, 6, 6, 6, 6, 6, 13, 13]
    traces = zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, heights)
    gpu = zl.prove_system(zkm, tr.SYSTEM_ALL_STARK, traces)
    cpu = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, traces)
    assert _first_diff(gpu, cpu) is None


@pytest.mark.parametrize("workload", ["U18", "U20"])
def test_all_stark_benchmark_config_equals_oracle(zkm, orc, workload):
    """The configuration bench.py reports (Arithmetic/Cpu/Memory at 2^20 rows, Logic 2^18, the other
    tables 2^6; U18 is the same shape at a quarter of the rows): the GPU proof through zkm_b200_prove_with_traces (host
    columns, the drop-in call) equals the oracle's proof of the same traces word for word.  One oracle proof at U20 takes a few
    minutes of host time; nothing is sampled or extrapolated."""
    import hashlib
    import bench
    heights = bench.workload_log_heights(workload)
    traces = zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, heights)
    gpu = zl.prove_with_traces(zkm, traces)
    orc.orc_set_threads(bench.host_threads())
    cpu = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, traces)
    assert _first_diff(gpu, cpu) is None
    assert hashlib.sha256(gpu.tobytes()).hexdigest() == hashlib.sha256(cpu.tobytes()).hexdigest()


def test_all_stark_valid_trace_proof_verifies(zkm, orc):
    """The drop-in entry point on a VALID 12-table trace (MIPS program with syscalls and the Keccak / SHA-256 / Poseidon
    precompiles, tests/traces.py all_stark_valid_traces): the GPU proof is accepted by the restated verifier
    (verifier.rs logic, all 15 cross-table lookups) and equals the oracle's proof word for word."""
    traces = tr.all_stark_valid_traces(orc)
    gpu = zl.prove_with_traces(zkm, traces)
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, gpu) is None
    cpu = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, traces)
    assert _first_diff(gpu, cpu) is None
    # a precompile result the sponge table does not back: the GPU proof is rejected like the oracle's
    import cpu_gen as cg
    bad = [t.copy() for t in traces]
    r = int(np.nonzero(bad[1][cg.IS_KECCAK_SPONGE])[0][0])
    bad[1][cg.GENERAL, r] += 1
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, zl.prove_with_traces(zkm, bad)) is not None


def test_sha2_guest_segment_valid_proof(zkm, orc):
    """BASELINE config 1 analogue (a SHA-256 guest, one 2^16-row segment): the test program followed by a loop of 500
    sha_extend + sha_compress precompile calls -> CPU table 2^16 rows, Logic 2^19, Memory 2^20, SHA tables 2^15..2^16, all
    valid.  The GPU proof must be accepted by the restated verifier; at 40 blocks (CPU 2^13) it must also equal the
    oracle's proof word for word."""
    small = tr.all_stark_valid_traces(orc, sha_blocks=40)
    gpu = zl.prove_with_traces(zkm, small)
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, gpu) is None
    assert _first_diff(gpu, binding.prove_system(orc, tr.SYSTEM_ALL_STARK, small)) is None
    big = tr.all_stark_valid_traces(orc, sha_blocks=500)
    heights = [t.shape[1].bit_length() - 1 for t in big]
    assert heights[1] == 16 and heights[10] == 19 and heights[11] == 20, heights
    gpu = zl.prove_with_traces(zkm, big)
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, gpu) is None
    # and a single wrong message-schedule word anywhere in the segment is fatal
    import cpu_gen as cg
    rows = np.nonzero(big[1][cg.IS_SHA_EXTEND_SPONGE])[0]
    big[1][cg.GENERAL, int(rows[len(rows) // 2])] += 1
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, zl.prove_with_traces(zkm, big)) is not None


def test_grouped_upload_commit_is_identical(zkm, orc, monkeypatch):
    """Host tables above a size threshold are uploaded and transformed in 4 column groups (NTTs overlap the upload); the proof
    must not change.  The threshold is lowered so that the small test tables take that path."""
    traces = tr.keccak_system_traces()
    ref = zl.prove_system(zkm, tr.SYSTEM_KECCAK, traces)
    monkeypatch.setenv("ZKM_GROUP_BYTES", "4096")
    grouped = zl.prove_system(zkm, tr.SYSTEM_KECCAK, traces)
    assert _first_diff(ref, grouped) is None
    assert _first_diff(grouped, binding.prove_system(orc, tr.SYSTEM_KECCAK, traces)) is None


def test_two_worker_contexts_prove_concurrently(zkm, orc):
    """Two host threads, each bound to its own worker context (include/zkm_b200.h), prove different Systems at the same time;
    every proof equals the one computed alone on the process-wide context."""
    import threading
    jobs = [(tr.SYSTEM_KECCAK, tr.keccak_system_traces()), (tr.SYSTEM_CPU, tr.cpu_system_traces())]
    alone = [zl.prove_system(zkm, sid, t) for sid, t in jobs]
    workers = [zl.Worker(zkm) for _ in jobs]
    out, errors = [[], []], []

    def body(i):
        try:
            with workers[i]:
                for _ in range(3):
                    out[i].append(zl.prove_system(zkm, jobs[i][0], jobs[i][1]))
        except BaseException as e:
            errors.append(e)
    threads = [threading.Thread(target=body, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for w in workers:
        w.close()
    assert not errors, errors
    for i in range(2):
        assert len(out[i]) == 3
        for p in out[i]:
            assert _first_diff(p, alone[i]) is None
    # and the process-wide context still works after the workers are gone
    assert _first_diff(zl.prove_system(zkm, jobs[0][0], jobs[0][1]), alone[0]) is None


def test_row_major_tables_give_the_same_proof(zkm, orc):
    """zkm_b200_prove_with_trace_rows: the ten tables the reference generates row by row (everything but Arithmetic and Memory,
    witness/traces.rs:274-305) are handed over as rows and transposed on the device (reference util.rs:37-47 does it on
    the CPU); the proof must be the one the column-major call returns."""
    traces = tr.all_stark_valid_traces(orc, sha_blocks=12)
    ref = zl.prove_with_traces(zkm, traces)
    as_rows = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10}
    mixed = [np.ascontiguousarray(t.T) if k in as_rows else t for k, t in enumerate(traces)]
    got = zl.prove_with_trace_rows(zkm, mixed, as_rows)
    assert _first_diff(ref, got) is None
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, got) is None


def test_arithmetic_rows_get_their_range_checks_on_the_device(zkm, orc):
    """Arithmetic rows as ArithmeticStark::generate_trace has them before generate_range_checks (arithmetic_stark.rs:155-192):
    the counter and frequency columns come from the device (reference :127-153) and the proof equals the column-major one."""
    import arith_gen as ag
    traces = tr.all_stark_valid_traces(orc)
    ref = zl.prove_with_traces(zkm, traces)
    rows = np.ascontiguousarray(traces[0].T).copy()
    rows[:, ag.RANGE_COUNTER] = 0
    rows[:, ag.RC_FREQUENCIES] = 0
    mixed = [rows] + traces[1:]
    got = zl.prove_with_trace_rows(zkm, mixed, {0})
    assert _first_diff(ref, got) is None
    rows[5, ag.IN2] = 1 << 16                  # not range-checkable
    with pytest.raises(zl.ZkmError, match="exceeds the max range value"):
        zl.prove_with_trace_rows(zkm, [rows] + traces[1:], {0})


def test_memory_table_generated_on_the_device(zkm, orc):
    """zkm_b200_memory_trace (sort, fill_gaps, padding, flags, range check, frequencies on the device) against the restated
    MemoryStark::generate_trace (tests/cpu_gen.py memory_generate_trace, memory_stark.rs:133-244), bit for bit."""
    import cpu_gen as cg
    import cpu_program as cp
    cases = []
    image, end = cp.build(with_syscalls=True, sha_blocks=3)
    cpu = cg.MiniCpu(image, cp.ENTRY, image_id_words=(cp.IMAGE_ID, list(range(1, 10))))
    while cpu.pc != end:
        cpu.step()
    cases.append(cpu.mem_ops)                                   # address gaps between code, data and hash regions
    rng = np.random.default_rng(9)
    for n_ops, amax, tmax in ((1, 10, 10), (2, 1 << 20, 5), (64, 50, 1 << 16), (1000, 1 << 24, 1 << 20), (4096, 300, 4000), (5000, 40, 10)):
        ops = []
        for i in range(n_ops):
            seg = int(rng.integers(0, 6))
            ops.append((int(rng.integers(0, 2)), seg, int(rng.integers(0, 3 if seg == 4 else amax)), int(rng.integers(0, tmax)) * 10,
                        int(rng.integers(0, 2)), int(rng.integers(0, 1 << 32)), int(rng.integers(0, 2))))
        cases.append(ops)
    for ops in cases:
        want = cg.memory_generate_trace(list(ops))
        got = zl.memory_trace(zkm, np.array(ops, dtype=np.uint64))
        assert got.shape == want.shape, (got.shape, want.shape)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (len(ops), bad[:5], got[:, bad[0][1]], want[:, bad[0][1]])
    # and the device-generated table drives the AllStark proof like the host-generated one
    traces, cpu = tr.all_stark_valid_traces(orc, return_cpu=True)
    dev_mem = zl.memory_trace(zkm, np.array(cpu.mem_ops, dtype=np.uint64))
    assert (dev_mem == traces[11]).all()
    proof = zl.prove_with_traces(zkm, traces[:11] + [dev_mem])
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, proof) is None
    # one call: the table is generated on the device and stays there
    direct = zl.prove_with_memory_ops(zkm, traces, np.array(cpu.mem_ops, dtype=np.uint64))
    assert _first_diff(direct, proof) is None


def test_timing_scopes_use_the_reference_names(zkm, orc):
    """zkm_b200_last_timing: device-time scopes keyed by the reference's TimingTree strings (prover.rs:146,152,204,250-411,513,
    545,578,620), nested like the reference's."""
    traces = tr.all_stark_valid_traces(orc)
    zkm.zkm_b200_timing_enable(1)
    try:
        zl.prove_with_traces(zkm, traces)
        p = zkm.zkm_b200_last_timing()
        text = C.string_at(p).decode()
        zkm.zkm_b200_free_string(p)
    finally:
        zkm.zkm_b200_timing_enable(0)
    rows = [ln.split("\t") for ln in text.splitlines()]
    names = [r[2] for r in rows]
    assert names[0] == "compute all trace commitments" and rows[0][0] == "0"
    for t in ("Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge", "ShaExtend", "ShaExtendSponge", "ShaCompress",
              "ShaCompressSponge", "Logic", "Memory"):
        assert f"compute trace commitment for {t}" in names
    assert "compute all proofs given commitments" in names
    for t in ("Arithmetic", "CPU", "Poseidon", "Poseidon sponge", "Keccak", "Keccak sponge", "SHA Extend", "SHA Extend sponge", "SHA Compress",
              "SHA Compress sponge", "Logic", "Memory"):
        assert f"prove {t} STARK" in names
    i = names.index("prove CPU STARK")
    inner = [n for r, n in zip(rows[i + 1:i + 7], names[i + 1:i + 7]) if r[0] == "2"]
    assert inner == ["compute CTL data + lookup helper columns", "compute auxiliary polynomials commitment", "compute quotient polys",
                     "compute quotient commitment", "compute openings", "compute openings proof"]
    assert all(float(r[1]) >= 0 for r in rows) and sum(float(r[1]) for r in rows if r[0] == "0") > 0
    # switched off again: the next proof records nothing
    zl.prove_with_traces(zkm, traces)
    p = zkm.zkm_b200_last_timing()
    assert C.string_at(p) == b""
    zkm.zkm_b200_free_string(p)


def test_logic_and_poseidon_tables_generated_on_the_device(zkm, orc):
    """zkm_b200_table_from_ops: the Logic table from (operator, input0, input1) logs and the Poseidon table from
    (12 inputs, timestamp) logs equal the reference generators as restated in tests/traces.py (logic.rs:108-183) and in the
    oracle (poseidon_stark.rs:51-145), including the padding rows and the minimum height."""
    rng = np.random.default_rng(5)
    for n_ops in (0, 1, 63, 64, 65, 1000, 5000):
        ops = [(int(k), int(x), int(y)) for k, x, y in zip(rng.integers(0, 4, n_ops), rng.integers(0, 1 << 32, n_ops), rng.integers(0, 1 << 32, n_ops))]
        if n_ops:
            ops[0] = (3, 0xFFFFFFFF, 0)
        arr = np.array(ops, dtype=np.uint64).reshape(n_ops, 3)
        got = zl.table_from_ops(zkm, 10, arr)
        n = max(64, 1 << max(0, (n_ops - 1).bit_length()))
        assert got.shape == (69, n)
        assert (got == tr.logic_trace_from_ops(ops, n.bit_length() - 1)).all()
    with pytest.raises(zl.ZkmError, match="logic operation out of range"):
        zl.table_from_ops(zkm, 10, np.array([[4, 1, 2]], dtype=np.uint64))
    with pytest.raises(zl.ZkmError, match="no device-side generator"):
        zl.table_from_ops(zkm, 1, np.zeros((1, 26), dtype=np.uint64))      # Cpu: the interpreter's own rows
    from oracle.binding import u64ptr
    for n_ops in (0, 3, 64, 200):
        inputs = rng.integers(0, tr.P, size=(n_ops, 12), dtype=np.uint64)
        if n_ops:
            inputs[0] = tr.P - 1
        ts = np.arange(7, 7 + n_ops, dtype=np.uint64)
        got = zl.table_from_ops(zkm, 2, np.concatenate([inputs, ts[:, None]], axis=1).reshape(n_ops, 13))
        n = max(64, 1 << max(0, (n_ops - 1).bit_length()))
        full_in = np.zeros((n, 12), dtype=np.uint64); full_in[:n_ops] = inputs
        full_ts = np.zeros(n, dtype=np.uint64); full_ts[:n_ops] = ts
        rows = np.zeros((n, 262), dtype=np.uint64)
        orc.orc_gen_poseidon_rows(u64ptr(full_in), u64ptr(full_ts), n, u64ptr(rows))
        rows[n_ops:, 0] = 0                                   # padding rows: FILTER = 0 (poseidon_stark.rs:119-122)
        assert got.shape == (262, n) and (got == rows.T).all()
    with pytest.raises(zl.ZkmError, match="canonical"):
        zl.table_from_ops(zkm, 2, np.full((1, 13), tr.P, dtype=np.uint64))


def test_prove_with_ops_gives_the_same_proof(zkm, orc):
    """zkm_b200_prove_with_ops: the Logic table enters as its operation log and is generated on the device inside the prove
    call; the proof equals the proof over the host-built table (same 12-table valid trace as the drop-in test)."""
    traces = tr.all_stark_valid_traces(orc)
    logic = traces[10]
    used = int(logic[:4].sum(axis=0).astype(bool).sum())
    assert (logic[:4, :used].sum(axis=0) == 1).all() and not logic[:, used:].any()      # operations first, zero padding after
    x = sum(logic[4 + i, :used].astype(np.uint64) << np.uint64(i) for i in range(32))
    y = sum(logic[36 + i, :used].astype(np.uint64) << np.uint64(i) for i in range(32))
    k = np.argmax(logic[:4, :used], axis=0).astype(np.uint64)
    ops = np.stack([k, x, y], axis=1)
    assert (zl.table_from_ops(zkm, 10, ops) == logic).all()
    ref = zl.prove_with_traces(zkm, traces)
    got = zl.prove_with_ops(zkm, traces, {10: ops})
    assert _first_diff(got, ref) is None


def test_arithmetic_table_generated_on_the_device(zkm, orc):
    """zkm_b200_table_from_ops for the Arithmetic table: all 26 operations incl. the two-row ones (DIV, DIVU, SRL(V), SRA(V)) and
    edge operands, against tests/arith_gen.py (the Python restatement of arithmetic/*.rs generate_*), range-check columns
    included; the generated table satisfies the constraints and proves to the same proof as the host-built one."""
    import arith_gen as ag
    for count, seed in ((0, 1), (1, 2), (300, 3), (30000, 11)):
        ops = ag.random_ops(count, seed)
        rows = sum(2 if o[0] in (ag.IS_DIV, ag.IS_DIVU, ag.IS_SRL, ag.IS_SRLV, ag.IS_SRA, ag.IS_SRAV) else 1 for o in ops)
        log_n = max(16, (rows - 1).bit_length() if rows else 0)
        want = ag.arithmetic_trace(ops, log_n)
        got = zl.table_from_ops(zkm, 0, np.array(ops, dtype=np.uint64).reshape(count, 3))
        assert got.shape == want.shape
        bad = np.argwhere(got != want)
        assert bad.size == 0, (count, bad[:5], [ops[int(r)] if int(r) < len(ops) else None for r in bad[:5, 1]])
    ops = ag.random_ops(30000, 11)
    arr = np.array(ops, dtype=np.uint64)
    t = zl.table_from_ops(zkm, 0, arr)
    assert orc.orc_check_table_constraints(0, binding.col_ptrs(t), 54, 16) == 0
    assert _first_diff(zl.prove_system(zkm, tr.SYSTEM_ARITH, [t]), zl.prove_system(zkm, tr.SYSTEM_ARITH, [tr.arithmetic_trace()])) is None
    with pytest.raises(zl.ZkmError, match="arithmetic operation out of range"):
        zl.table_from_ops(zkm, 0, np.array([[ag.IS_DIVU, 5, 0]], dtype=np.uint64))


def test_keccak_table_generated_on_the_device(zkm, orc):
    """zkm_b200_table_from_ops for the Keccak-f table (2431 columns, 24 rows per permutation) against tests/hash_gen.py (the
    Python restatement of keccak_stark.rs:62-237, itself checked against hashlib's SHA3): every cell, the zero padding and
    the minimum height; the generated table satisfies the transcribed constraints."""
    import hash_gen as hg
    rng = np.random.default_rng(9)
    for perms in (0, 1, 2, 3, 11):
        ins = [([int(v) for v in rng.integers(0, 1 << 64, size=25, dtype=np.uint64)], 10 + i) for i in range(perms)]
        ops = np.array([list(i) + [ts] for i, ts in ins], dtype=np.uint64).reshape(perms, 26)
        got = zl.table_from_ops(zkm, 4, ops)
        n = max(64, 1 << max(0, (24 * perms - 1).bit_length()))
        want = hg.keccak_trace(ins, n.bit_length() - 1)
        assert got.shape == want.shape == (2431, n)
        assert (got == want).all()
    assert orc.orc_check_table_constraints(4, binding.col_ptrs(got), 2431, 9) == 0


def test_row_major_table_with_2_21_rows(zkm):
    """the row -> column transpose put the row tiles on gridDim.y (<= 65535), so any row-major table with >= 2^21 rows
    failed to launch.  A Logic table of 2^21 rows handed over as rows must give the proof of the same table handed over as
    columns (synthetic traces: the two proofs are compared with each other)."""
    heights = [16, 6, 6, 6, 6, 6, 6, 6, 6, 6, 21, 6]
    traces = zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, heights)
    ref = zl.prove_with_traces(zkm, traces)
    mixed = [np.ascontiguousarray(t.T) if k == 10 else t for k, t in enumerate(traces)]
    got = zl.prove_with_trace_rows(zkm, mixed, {10})
    assert _first_diff(ref, got) is None


def test_mixed_shape_workers_soak(zkm):
    """every worker context caches freed device blocks in its own arena.  Three workers proving segments of DIFFERENT
    shapes, shapes rotating between iterations (so that blocks cached for one shape are useless for the next), must keep
    producing the proofs computed alone -- the arenas are registered, a failed cudaMalloc trims all of them, and the bytes
    cached device-wide are capped."""
    import threading
    shapes = [[16, 14, 6, 6, 6, 6, 6, 6, 6, 6, 12, 14], [16, 15, 6, 6, 6, 6, 6, 7, 6, 6, 13, 13], [17, 13, 6, 6, 7, 6, 6, 6, 6, 6, 15, 16]]
    segs = [zl.synth_traces(zkm, tr.SYSTEM_ALL_STARK, h, seed=0x5EED000000000000 + (i << 40)) for i, h in enumerate(shapes)]
    alone = [zl.prove_with_traces(zkm, s) for s in segs]
    workers = [zl.Worker(zkm) for _ in range(3)]
    errors = []

    def body(i):
        try:
            with workers[i]:
                for it in range(4):
                    k = (i + it) % 3
                    assert _first_diff(zl.prove_with_traces(zkm, segs[k]), alone[k]) is None
        except BaseException as e:
            errors.append(e)
    threads = [threading.Thread(target=body, args=(i,)) for i in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for w in workers:
        w.close()
    assert not errors, errors
