"""Stage-level parity (SURVEY §8(b) "finer-grained seam"): for one table of a System and caller-given challenges, the auxiliary
columns (cross_table_lookup_data + lookup_helper_columns, rows a4/a5), the quotient coefficients (compute_quotient_polys + coset
iFFT, a6/a7) and StarkOpeningSet::new (a8) of the CUDA path against the oracle's stage functions -- every word, not only through
the final proof bytes.  The CPU test pins the oracle's stage function to the oracle's own proof (same code path the verifier
accepts): with the proof's challenges it must reproduce the proof's openings."""
import numpy as np
import pytest

import traces as tr
from oracle import binding

CH = dict(ctl=[[0x1234567890ABCDEF % tr.P, 0x0FEDCBA987654321], [0x1111111122222222, 0x3333333344444444]],
          alphas=[0x55555555AAAAAAAA % tr.P, 0x0123456701234567], zeta=[0x7777777788888888, 0x99999999AAAAAAAB % tr.P])


def _openings_from_proof(proof, table_index, ncols_by_table):
    """Walks the flat proof buffer (include/zkm_b200.h) to table `table_index`; returns (ctl challenges, openings words in the
    stage layout)."""
    w = [int(x) for x in proof]
    p = 3
    nch = w[p]; p += 1
    ctl = [[w[p + 2 * k], w[p + 2 * k + 1]] for k in range(nch)]; p += 2 * nch
    p += 16
    p += 1 + w[p]                                  # userdata
    for t in range(table_index + 1):
        p += 12                                    # init_challenger_state
        for _ in range(3):                         # three caps
            p += 1 + 4 * w[p]
        vecs = []
        for width in (2, 2, 2, 2, 1, 2):           # local, next, aux, aux_next, ctl_zs_first, quotient
            k = w[p]; p += 1
            vecs.append(w[p:p + width * k]); p += width * k
        if t == table_index:
            return ctl, np.array(sum(vecs, []), dtype=np.uint64)
        # skip the FriProof
        ncaps = w[p]; p += 1
        for _ in range(ncaps):
            p += 1 + 4 * w[p]
        nq = w[p]; p += 1
        for _ in range(nq):
            no = w[p]; p += 1
            for _ in range(no):
                p += 1 + w[p]
                p += 1 + 4 * w[p]
            ns = w[p]; p += 1
            for _ in range(ns):
                p += 1 + 2 * w[p]
                p += 1 + 4 * w[p]
        p += 1 + 2 * w[p]
        p += 1
    raise AssertionError


def test_oracle_stage_openings_are_the_proofs_openings(orc):
    """With zeta chosen freely the openings differ from the proof's; aux columns and the opening AT 1 (ctl_zs_first) do not
    depend on alpha / zeta -- they must equal the proof's, under the proof's CTL challenges."""
    traces = tr.cpu_system_traces()
    proof = binding.prove_system(orc, tr.SYSTEM_CPU, traces)
    for t, cols in enumerate(traces):
        ctl, opn = _openings_from_proof(proof, t, None)
        aux, quot, got = binding.stage_table(orc, tr.SYSTEM_CPU, t, cols, ctl, CH["alphas"], CH["zeta"])
        C, A = cols.shape[0], aux.shape[0]
        nz = got.size - 4 * C - 4 * A - 8
        assert nz > 0 and (got[4 * C + 4 * A:4 * C + 4 * A + nz] == opn[4 * C + 4 * A:4 * C + 4 * A + nz]).all()
        # ctl_zs_first is also row 0 of the Z columns (value at 1 = first row of the values on H)
        assert (aux[A - nz:, 0] == got[4 * C + 4 * A:4 * C + 4 * A + nz]).all()
        assert quot.shape == (2, 2 * cols.shape[1])


@pytest.mark.gpu
@pytest.mark.parametrize("sid", [tr.SYSTEM_LOGIC, tr.SYSTEM_ARITH, tr.SYSTEM_MINI3, tr.SYSTEM_KECCAK, tr.SYSTEM_SHA_COMPRESS, tr.SYSTEM_CPU])
def test_stage_outputs_match_oracle(zkm, orc, sid):
    from test_gpu_prove import _traces
    from zkm_b200 import lib as zl
    traces = _traces(orc, sid, 0)
    for t, cols in enumerate(traces):
        a_g, q_g, o_g = zl.stage_table(zkm, sid, t, cols, CH["ctl"], CH["alphas"], CH["zeta"])
        a_o, q_o, o_o = binding.stage_table(orc, sid, t, cols, CH["ctl"], CH["alphas"], CH["zeta"], max_aux=a_g.shape[0] + 4)
        assert a_g.shape == a_o.shape and (a_g == a_o).all(), (sid, t, "auxiliary columns")
        assert (q_g == q_o).all(), (sid, t, "quotient coefficients")
        assert o_g.size == o_o.size and (o_g == o_o).all(), (sid, t, "openings")


@pytest.mark.gpu
def test_stage_errors(zkm):
    from zkm_b200 import lib as zl
    cols = tr.logic_trace(6)
    with pytest.raises(zl.ZkmError, match="canonical"):
        zl.stage_table(zkm, tr.SYSTEM_LOGIC, 0, cols, [[tr.P, 1], [2, 3]], CH["alphas"], CH["zeta"])
    with pytest.raises(zl.ZkmError, match="no such table"):
        zl.stage_table(zkm, tr.SYSTEM_LOGIC, 3, cols, CH["ctl"], CH["alphas"], CH["zeta"])
    with pytest.raises(zl.ZkmError, match="subgroup"):
        zl.stage_table(zkm, tr.SYSTEM_LOGIC, 0, cols, CH["ctl"], CH["alphas"], [1, 0])
