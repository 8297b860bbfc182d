//! `prover/src/b200_ops.rs` -- proving straight from `Traces`: every table but Cpu crosses the FFI as its
//! operation log and is generated on the device (`zkm_b200_prove_with_ops`, include/zkm_b200.h); the Cpu rows go row-major.
//! This replaces `Traces::into_tables` (witness/traces.rs:230-318) + `prove_with_traces` for the call chain
//! `prove_with_outputs` -> `generate_traces` -> `prove_with_traces` (prover.rs:58-128).
//!
//! SOURCE ONLY (no cargo in this image), written against zkMIPS/zkm @ 04117ce3.  Needs, besides `b200.rs`:
//!  * `pub(crate)` on the three private fields of `logic::Operation` (logic.rs:100-107: operator, input0, input1) -- part of
//!    `shim/prover_b200.patch`;
//!  * `generate_traces` to hand back the `Traces` instead of calling `into_tables` (generation/mod.rs:169-186).
//! The log formats are the ones documented next to `zkm_op_log_t`; `tests/test_gpu_tracegen.py` checks each of them against the
//! restated reference generators and the all-logs proof against the proof over host-built tables.
use std::ffi::{c_char, c_int, c_void};

use anyhow::Result;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::PrimeField64;

use crate::all_stark::{Table, NUM_TABLES};
use crate::arithmetic::Operation as ArithOp;
use crate::b200::{decode_all_proof, stark_config_to_c, take_error, ZkmStarkConfig, ZkmTable};
use crate::config::StarkConfig;
use crate::cpu::columns::NUM_CPU_COLUMNS;
use crate::logic;
use crate::proof::{AllProof, PublicValues};
use crate::witness::memory::{MemoryAddress, MemoryOpKind};
use crate::witness::traces::Traces;
use plonky2::plonk::config::PoseidonGoldilocksConfig;

type F = GoldilocksField;

#[repr(C)]
pub struct ZkmTableRows {
    pub rows: *const u64,
    pub ncols: u32,
    pub log_n: u32,
}
#[repr(C)]
pub struct ZkmOpLog {
    pub ops: *const u64,
    pub n_ops: usize,
}

extern "C" {
    fn zkm_b200_prove_with_ops(
        tables: *const ZkmTable, row_tables: *const ZkmTableRows, op_logs: *const ZkmOpLog, roots_before: *const u32,
        roots_after: *const u32, userdata: *const u8, userdata_len: u32, cfg: *const ZkmStarkConfig, proof_out: *mut *mut u64,
        proof_words: *mut usize, err: *mut *mut c_char,
    ) -> c_int;
    fn zkm_b200_free(p: *mut c_void);
}

fn le_u32(b: &[u8]) -> u64 {
    u32::from_le_bytes(b.try_into().unwrap()) as u64
}

/// The variable-width log of the two byte sponges: word 0 = total words, then per operation
/// context, segment, timestamp, len, n_addr, virt[n_addr], the input bytes 8 per word (little endian).
fn sponge_log(ops: impl Iterator<Item = (Vec<MemoryAddress>, usize, Vec<u8>)>) -> (Vec<u64>, usize) {
    let mut log = vec![0u64];
    let mut n = 0;
    for (addr, timestamp, input) in ops {
        // both generators read context from base_address[0] and segment from base_address[Segment::Code as usize] = [0]
        // (keccak_sponge_stark.rs:367-368, poseidon_sponge_stark.rs:318-319)
        log.extend([addr[0].context as u64, addr[0].segment as u64, timestamp as u64, input.len() as u64, addr.len() as u64]);
        log.extend(addr.iter().map(|a| a.virt as u64));
        for chunk in input.chunks(8) {
            let mut w = [0u8; 8];
            w[..chunk.len()].copy_from_slice(chunk);
            log.push(u64::from_le_bytes(w));
        }
        n += 1;
    }
    log[0] = log.len() as u64;
    (log, n)
}

/// `Traces` -> the eleven operation logs (flat u64 words, number of operations), indexed by `Table`.
pub fn op_logs(traces: &Traces<F>) -> [(Vec<u64>, usize); NUM_TABLES] {
    let mut logs: [(Vec<u64>, usize); NUM_TABLES] = Default::default();
    let mut put = |t: Table, words: Vec<u64>, n: usize| logs[t as usize] = (words, n);
    // Arithmetic: the IS_* column index (BinaryOperator::row_filter, arithmetic/mod.rs:135-165), input0, input1
    put(
        Table::Arithmetic,
        traces.arithmetic_ops.iter().flat_map(|ArithOp::BinaryOperation { operator, input0, input1, .. }| {
            [operator.row_filter() as u64, *input0 as u64, *input1 as u64]
        }).collect(),
        traces.arithmetic_ops.len(),
    );
    put(
        Table::Logic,
        traces.logic_ops.iter().flat_map(|op| {
            let k = match op.operator { logic::Op::And => 0u64, logic::Op::Or => 1, logic::Op::Xor => 2, logic::Op::Nor => 3 };
            [k, op.input0 as u64, op.input1 as u64]
        }).collect(),
        traces.logic_ops.len(),
    );
    // Memory: context, segment, virt, timestamp, is_read, value, filter in push order (memory_stark.rs:44-72)
    put(
        Table::Memory,
        traces.memory_ops.iter().flat_map(|m| {
            [m.address.context as u64, m.address.segment as u64, m.address.virt as u64, m.timestamp as u64,
             (m.kind == MemoryOpKind::Read) as u64, m.value as u64, m.filter as u64]
        }).collect(),
        traces.memory_ops.len(),
    );
    put(
        Table::Poseidon,
        traces.poseidon_inputs.iter().flat_map(|(st, ts)| st.iter().map(|x| x.to_canonical_u64()).chain([*ts as u64]).collect::<Vec<_>>()).collect(),
        traces.poseidon_inputs.len(),
    );
    put(
        Table::Keccak,
        traces.keccak_inputs.iter().flat_map(|(lanes, ts)| lanes.iter().copied().chain([*ts as u64]).collect::<Vec<_>>()).collect(),
        traces.keccak_inputs.len(),
    );
    let (w, n) = sponge_log(traces.poseidon_sponge_ops.iter().map(|o| (o.base_address.clone(), o.timestamp, o.input.clone())));
    put(Table::PoseidonSponge, w, n);
    let (w, n) = sponge_log(traces.keccak_sponge_ops.iter().map(|o| (o.base_address.clone(), o.timestamp, o.input.clone())));
    put(Table::KeccakSponge, w, n);
    // ShaExtend: ([u8; 16], timestamp) = w[i-15], w[i-2], w[i-16], w[i-7] little endian (sha_extend_stark.rs:156-176)
    put(
        Table::ShaExtend,
        traces.sha_extend_inputs.iter().flat_map(|(b, ts)| (0..4).map(|k| le_u32(&b[4 * k..4 * k + 4])).chain([*ts as u64]).collect::<Vec<_>>()).collect(),
        traces.sha_extend_inputs.len(),
    );
    put(
        Table::ShaExtendSponge,
        traces.sha_extend_sponge_ops.iter().flat_map(|o| {
            let mut w = vec![o.i as u64];
            w.extend((0..4).map(|k| le_u32(&o.input[4 * k..4 * k + 4])));
            w.extend((0..4).map(|k| o.base_address[k].virt as u64));
            w.extend([o.output_address.virt as u64, o.base_address[0].context as u64, o.base_address[0].segment as u64, o.timestamp as u64]);
            w
        }).collect(),
        traces.sha_extend_sponge_ops.len(),
    );
    // ShaCompress: one entry per ROW: ([u8; 41] = a..h, w_i, k_i (little endian), round), w_i address, timestamp
    put(
        Table::ShaCompress,
        traces.sha_compress_inputs.iter().flat_map(|(b, addr, ts)| {
            let mut w: Vec<u64> = (0..10).map(|k| le_u32(&b[4 * k..4 * k + 4])).collect();
            w.extend([b[40] as u64, addr.virt as u64, addr.segment as u64, addr.context as u64, *ts as u64]);
            w
        }).collect(),
        traces.sha_compress_inputs.len(),
    );
    put(
        Table::ShaCompressSponge,
        traces.sha_compress_sponge_ops.iter().flat_map(|o| {
            let mut w: Vec<u64> = (0..8).map(|k| le_u32(&o.input[4 * k..4 * k + 4])).collect();
            w.extend(o.w_i_s.iter().map(|b| u32::from_le_bytes(*b) as u64));
            w.extend((0..8).map(|k| o.base_address[k].virt as u64));
            let ws = &o.base_address[8];
            w.extend([ws.virt as u64, ws.segment as u64, ws.context as u64, o.base_address[0].context as u64,
                      o.base_address[0].segment as u64, o.timestamp as u64]);
            w
        }).collect(),
        traces.sha_compress_sponge_ops.len(),
    );
    logs
}

/// `Traces::into_tables` + `prove_with_traces` in one device-side call.
pub fn prove_with_ops_b200(config: &StarkConfig, traces: Traces<F>, public_values: PublicValues) -> Result<AllProof<F, PoseidonGoldilocksConfig, 2>> {
    let logs = op_logs(&traces);
    // the Cpu table: `Vec<CpuColumnsView<F>>` is `Vec<[F; NUM_CPU_COLUMNS]>` in memory (#[repr(C)], cpu/columns/mod.rs:68-118),
    // padded to a power of two >= 64 with the reference's padding rows before it gets here (generation/mod.rs pad_cpu)
    let n = traces.cpu.len();
    anyhow::ensure!(n.is_power_of_two() && n >= 64, "the Cpu rows must be padded to a power of two");
    let cpu_words: Vec<u64> = traces.cpu.iter().flat_map(|r| {
        let row: &[F; NUM_CPU_COLUMNS] = std::borrow::Borrow::borrow(r);
        row.iter().map(|x| x.to_canonical_u64()).collect::<Vec<_>>()
    }).collect();
    let empty = || ZkmTable { cols: core::ptr::null(), ncols: 0, log_n: 0 };
    let tables: Vec<ZkmTable> = (0..NUM_TABLES).map(|_| empty()).collect();
    let mut rows: Vec<ZkmTableRows> = (0..NUM_TABLES).map(|_| ZkmTableRows { rows: core::ptr::null(), ncols: 0, log_n: 0 }).collect();
    rows[Table::Cpu as usize] = ZkmTableRows { rows: cpu_words.as_ptr(), ncols: NUM_CPU_COLUMNS as u32, log_n: n.trailing_zeros() };
    let c_logs: Vec<ZkmOpLog> = logs.iter().enumerate().map(|(t, (w, k))| {
        if t == Table::Cpu as usize { ZkmOpLog { ops: core::ptr::null(), n_ops: 0 } } else { ZkmOpLog { ops: w.as_ptr(), n_ops: *k } }
    }).collect();
    let cfg = stark_config_to_c(config)?;
    let (mut out, mut words, mut err): (*mut u64, usize, *mut c_char) = (core::ptr::null_mut(), 0, core::ptr::null_mut());
    let rc = unsafe {
        zkm_b200_prove_with_ops(
            tables.as_ptr(), rows.as_ptr(), c_logs.as_ptr(), public_values.roots_before.root.as_ptr(), public_values.roots_after.root.as_ptr(),
            public_values.userdata.as_ptr(), public_values.userdata.len() as u32, &cfg, &mut out, &mut words, &mut err,
        )
    };
    if rc != 0 {
        return Err(take_error(err));
    }
    let proof = decode_all_proof(unsafe { std::slice::from_raw_parts(out, words) });
    unsafe { zkm_b200_free(out as *mut c_void) };
    proof
}
