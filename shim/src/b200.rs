//! `prover/src/b200.rs` -- the Rust half of the zkm_b200 drop-in.
//!
//! SOURCE ONLY: this image has no cargo/rustc and plonky2 is un-vendored, so this file has never been compiled here.
//! It is written against zkMIPS/zkm @ 04117ce3 and plonky2 0.1.4 (zkMIPS/plonky2 zkm_dev f1e28a6d) and is meant to be
//! copied to `prover/src/b200.rs`, declared in `prover/src/lib.rs` as `#[cfg(feature = "b200")] pub mod b200;`, with the
//! `prove_with_traces` patch of INTEGRATION.md section 2.  Everything it touches in the reference is cited by file:line.
//!
//! Contents
//!  * `extern "C"` declarations of include/zkm_b200.h (the same convention as the tree's only FFI,
//!    recursion/src/snark/snarks.rs:7-20: c_int status + malloc'ed error string);
//!  * `layout_pairs()` -- where THIS compiler put the fields of `CpuColumnsView` and of the non-`repr(C)` general views
//!    (cpu/columns/mod.rs:68-118,184-189; cpu/columns/general.rs:8-18,143-201), handed to `zkm_b200_layout_check`;
//!  * `prove_with_traces_b200` -- marshals `[Vec<PolynomialValues<F>>; NUM_TABLES]` into `zkm_table_t[12]`, calls the library,
//!    and `decode_all_proof` rebuilds `AllProof<GoldilocksField, PoseidonGoldilocksConfig, 2>` from the flat buffer
//!    (proof.rs:25-29,178-201,283-296; plonky2 fri/proof.rs, hash/merkle_proofs.rs, hash/merkle_tree.rs);
//!  * `log_last_timing` -- the device-time scopes keyed by the reference's TimingTree strings.
#![allow(clippy::needless_range_loop)]

use std::any::Any;
use std::ffi::{c_char, c_int, c_void, CStr};

use anyhow::{anyhow, bail, ensure, Result};
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::extension::{Extendable, FieldExtension};
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::proof::{FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep};
use plonky2::hash::hash_types::{HashOut, RichField};
use plonky2::hash::merkle_proofs::MerkleProof;
use plonky2::hash::merkle_tree::MerkleCap;
use plonky2::hash::poseidon::{PoseidonHash, PoseidonPermutation};
use plonky2::hash::hashing::PlonkyPermutation;
use plonky2::plonk::config::{GenericConfig, PoseidonGoldilocksConfig};
use plonky2_util::log2_strict; // as prover.rs:19 imports it

use crate::all_stark::NUM_TABLES;
use crate::config::StarkConfig;
use crate::cpu::columns::COL_MAP;
use crate::cross_table_lookup::{GrandProductChallenge, GrandProductChallengeSet};
use crate::proof::{AllProof, MemRoots, PublicValues, StarkOpeningSet, StarkProof, StarkProofWithMetadata};

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
const D: usize = 2;
type FE = QuadraticExtension<F>;
type H = PoseidonHash;

// ------------------------------------------------------------------------------------------- C ABI (include/zkm_b200.h)

#[repr(C)]
pub struct ZkmTable {
    pub cols: *const *const u64,
    pub ncols: u32,
    pub log_n: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ZkmStarkConfig {
    pub rate_bits: u32,
    pub cap_height: u32,
    pub pow_bits: u32,
    pub num_queries: u32,
    pub num_challenges: u32,
    pub arity_bits: u32,
    pub final_poly_bits: u32,
}

extern "C" {
    fn zkm_b200_init(device: c_int, err: *mut *mut c_char) -> c_int;
    fn zkm_b200_prove_with_traces(
        tables: *const ZkmTable, roots_before: *const u32, roots_after: *const u32, userdata: *const u8, userdata_len: u32,
        cfg: *const ZkmStarkConfig, proof_out: *mut *mut u64, proof_words: *mut usize, err: *mut *mut c_char,
    ) -> c_int;
    fn zkm_b200_layout_check(pairs: *const u32, n_pairs: usize, err: *mut *mut c_char) -> c_int;
    fn zkm_b200_timing_enable(on: c_int);
    fn zkm_b200_last_timing() -> *mut c_char;
    fn zkm_b200_free(p: *mut c_void);
    fn zkm_b200_free_string(s: *mut c_char);
}

pub(crate) fn take_error(err: *mut c_char) -> anyhow::Error {
    if err.is_null() {
        return anyhow!("zkm_b200: unknown error");
    }
    let m = unsafe { CStr::from_ptr(err) }.to_string_lossy().into_owned();
    unsafe { zkm_b200_free_string(err) };
    anyhow!(m) // the library returns the reference's own panic / ensure! texts (prover.rs:461-464,509,596-599)
}

// ------------------------------------------------------------------------------------------- layout handshake

/// Keys of `zkm_layout_key_t` (include/zkm_b200.h).  Kept as plain constants: the header is the single numbering source.
mod lk {
    pub const NUM_COLUMNS: u32 = 0;
    pub const CPU_IS_BOOTSTRAP_KERNEL: u32 = 100;
    pub const CPU_IS_EXIT_KERNEL: u32 = 101;
    pub const CPU_CONTEXT: u32 = 102;
    pub const CPU_CODE_CONTEXT: u32 = 103;
    pub const CPU_PROGRAM_COUNTER: u32 = 104;
    pub const CPU_NEXT_PROGRAM_COUNTER: u32 = 105;
    pub const CPU_IS_KERNEL_MODE: u32 = 106;
    pub const CPU_OP_BINARY_OP: u32 = 107;
    pub const CPU_OP_SYSCALL: u32 = 108;
    pub const CPU_BRANCH_SHOULD_JUMP: u32 = 109;
    pub const CPU_BRANCH_IS_NE: u32 = 110;
    pub const CPU_OPCODE_BITS: u32 = 111;
    pub const CPU_RS_BITS: u32 = 112;
    pub const CPU_RT_BITS: u32 = 113;
    pub const CPU_RD_BITS: u32 = 114;
    pub const CPU_SHAMT_BITS: u32 = 115;
    pub const CPU_FUNC_BITS: u32 = 116;
    pub const CPU_IS_POSEIDON_SPONGE: u32 = 117;
    pub const CPU_IS_KECCAK_SPONGE: u32 = 118;
    pub const CPU_IS_SHA_EXTEND_SPONGE: u32 = 119;
    pub const CPU_IS_SHA_COMPRESS_SPONGE: u32 = 120;
    pub const CPU_GENERAL: u32 = 121;
    pub const CPU_MEMIO_IS_LH: u32 = 122;
    pub const CPU_MEMIO_AUX_FILTER: u32 = 123;
    pub const CPU_CLOCK: u32 = 124;
    pub const CPU_MEM_CHANNELS: u32 = 125;
    pub const CPU_MEM_CHANNEL_STRIDE: u32 = 126;
    pub const CPU_CH_USED_REL: u32 = 127;
    pub const CPU_CH_IS_READ_REL: u32 = 128;
    pub const CPU_CH_ADDR_CONTEXT_REL: u32 = 129;
    pub const CPU_CH_ADDR_SEGMENT_REL: u32 = 130;
    pub const CPU_CH_ADDR_VIRTUAL_REL: u32 = 131;
    pub const CPU_CH_VALUE_REL: u32 = 132;
    pub const CPU_G_SYSCALL_COND_REL: u32 = 200;
    pub const CPU_G_SYSCALL_SYSNUM_REL: u32 = 201;
    pub const CPU_G_SYSCALL_A0_REL: u32 = 202;
    pub const CPU_G_SYSCALL_A1_REL: u32 = 203;
    pub const CPU_G_MISC_RS_BITS_REL: u32 = 204;
    pub const CPU_G_MISC_IS_MSB_REL: u32 = 205;
    pub const CPU_G_MISC_IS_LSB_REL: u32 = 206;
    pub const CPU_G_MISC_AUXM_REL: u32 = 207;
    pub const CPU_G_MISC_AUXL_REL: u32 = 208;
    pub const CPU_G_MISC_AUXS_REL: u32 = 209;
    pub const CPU_G_MISC_RD_INDEX_REL: u32 = 210;
    pub const CPU_G_MISC_RD_INDEX_EQ_0_REL: u32 = 211;
    pub const CPU_G_MISC_RD_INDEX_EQ_29_REL: u32 = 212;
    pub const CPU_G_IO_RS_LE_REL: u32 = 213;
    pub const CPU_G_IO_RT_LE_REL: u32 = 214;
    pub const CPU_G_IO_MEM_LE_REL: u32 = 215;
    pub const CPU_G_IO_AUX_RS0_MUL_RS1_REL: u32 = 216;
    pub const CPU_G_LOGIC_DIFF_PINV_REL: u32 = 217;
    pub const CPU_G_HASH_VALUE_REL: u32 = 218;
    pub const CPU_G_KHASH_VALUE_REL: u32 = 219;
    pub const CPU_G_SHASH_VALUE_REL: u32 = 220;
    pub const CPU_G_ELEMENT_VALUE_REL: u32 = 221;
}

/// (key, value) pairs describing the column layout as compiled by THIS rustc.  `COL_MAP` (cpu/columns/mod.rs:184-189) is
/// a `CpuColumnsView<usize>` whose every field holds its own column index, including through the `general` union's
/// accessors, so reading a field yields the offset rustc chose for it.
pub fn layout_pairs(num_columns: [usize; NUM_TABLES]) -> Vec<u32> {
    let m = &COL_MAP;
    let g = m.general.syscall().cond[0]; // every view starts at the union's offset 0 ...
    let mut v: Vec<(u32, usize)> = Vec::new();
    for (t, n) in num_columns.iter().enumerate() {
        v.push((lk::NUM_COLUMNS + t as u32, *n));
    }
    let misc = m.general.misc();
    let io = m.general.io();
    let sys = m.general.syscall();
    // ... unless rustc reordered a view's fields: the smallest offset of each view is the union base
    let base = *[sys.cond[0], sys.sysnum[0], sys.a0[0], sys.a1, misc.rs_bits[0], misc.auxm, io.rs_le[0], g].iter().min().unwrap();
    v.extend_from_slice(&[
        (lk::CPU_IS_BOOTSTRAP_KERNEL, m.is_bootstrap_kernel), (lk::CPU_IS_EXIT_KERNEL, m.is_exit_kernel),
        (lk::CPU_CONTEXT, m.context), (lk::CPU_CODE_CONTEXT, m.code_context), (lk::CPU_PROGRAM_COUNTER, m.program_counter),
        (lk::CPU_NEXT_PROGRAM_COUNTER, m.next_program_counter), (lk::CPU_IS_KERNEL_MODE, m.is_kernel_mode),
        (lk::CPU_OP_BINARY_OP, m.op.binary_op), (lk::CPU_OP_SYSCALL, m.op.syscall),
        (lk::CPU_BRANCH_SHOULD_JUMP, m.branch.should_jump), (lk::CPU_BRANCH_IS_NE, m.branch.is_ne),
        (lk::CPU_OPCODE_BITS, m.opcode_bits[0]), (lk::CPU_RS_BITS, m.rs_bits[0]), (lk::CPU_RT_BITS, m.rt_bits[0]),
        (lk::CPU_RD_BITS, m.rd_bits[0]), (lk::CPU_SHAMT_BITS, m.shamt_bits[0]), (lk::CPU_FUNC_BITS, m.func_bits[0]),
        (lk::CPU_IS_POSEIDON_SPONGE, m.is_poseidon_sponge), (lk::CPU_IS_KECCAK_SPONGE, m.is_keccak_sponge),
        (lk::CPU_IS_SHA_EXTEND_SPONGE, m.is_sha_extend_sponge), (lk::CPU_IS_SHA_COMPRESS_SPONGE, m.is_sha_compress_sponge),
        (lk::CPU_GENERAL, base), (lk::CPU_MEMIO_IS_LH, m.memio.is_lh), (lk::CPU_MEMIO_AUX_FILTER, m.memio.aux_filter),
        (lk::CPU_CLOCK, m.clock), (lk::CPU_MEM_CHANNELS, m.mem_channels[0].used),
        (lk::CPU_MEM_CHANNEL_STRIDE, m.mem_channels[1].used - m.mem_channels[0].used),
        (lk::CPU_CH_USED_REL, 0), (lk::CPU_CH_IS_READ_REL, m.mem_channels[0].is_read - m.mem_channels[0].used),
        (lk::CPU_CH_ADDR_CONTEXT_REL, m.mem_channels[0].addr_context - m.mem_channels[0].used),
        (lk::CPU_CH_ADDR_SEGMENT_REL, m.mem_channels[0].addr_segment - m.mem_channels[0].used),
        (lk::CPU_CH_ADDR_VIRTUAL_REL, m.mem_channels[0].addr_virtual - m.mem_channels[0].used),
        (lk::CPU_CH_VALUE_REL, m.mem_channels[0].value - m.mem_channels[0].used),
        (lk::CPU_G_SYSCALL_COND_REL, sys.cond[0] - base), (lk::CPU_G_SYSCALL_SYSNUM_REL, sys.sysnum[0] - base),
        (lk::CPU_G_SYSCALL_A0_REL, sys.a0[0] - base), (lk::CPU_G_SYSCALL_A1_REL, sys.a1 - base),
        (lk::CPU_G_MISC_RS_BITS_REL, misc.rs_bits[0] - base), (lk::CPU_G_MISC_IS_MSB_REL, misc.is_msb[0] - base),
        (lk::CPU_G_MISC_IS_LSB_REL, misc.is_lsb[0] - base), (lk::CPU_G_MISC_AUXM_REL, misc.auxm - base),
        (lk::CPU_G_MISC_AUXL_REL, misc.auxl - base), (lk::CPU_G_MISC_AUXS_REL, misc.auxs - base),
        (lk::CPU_G_MISC_RD_INDEX_REL, misc.rd_index - base), (lk::CPU_G_MISC_RD_INDEX_EQ_0_REL, misc.rd_index_eq_0 - base),
        (lk::CPU_G_MISC_RD_INDEX_EQ_29_REL, misc.rd_index_eq_29 - base),
        (lk::CPU_G_IO_RS_LE_REL, io.rs_le[0] - base), (lk::CPU_G_IO_RT_LE_REL, io.rt_le[0] - base),
        (lk::CPU_G_IO_MEM_LE_REL, io.mem_le[0] - base), (lk::CPU_G_IO_AUX_RS0_MUL_RS1_REL, io.aux_rs0_mul_rs1 - base),
        (lk::CPU_G_LOGIC_DIFF_PINV_REL, m.general.logic().diff_pinv - base),
        (lk::CPU_G_HASH_VALUE_REL, m.general.hash().value[0] - base), (lk::CPU_G_KHASH_VALUE_REL, m.general.khash().value[0] - base),
        (lk::CPU_G_SHASH_VALUE_REL, m.general.shash().value[0] - base),
        (lk::CPU_G_ELEMENT_VALUE_REL, m.general.element().value - base),
    ]);
    v.into_iter().flat_map(|(k, x)| [k, x as u32]).collect()
}

/// One-time set-up: create the device context and check the column layout.  Call before the first proof.
pub fn init(device: i32, num_columns: [usize; NUM_TABLES]) -> Result<()> {
    let mut err: *mut c_char = core::ptr::null_mut();
    if unsafe { zkm_b200_init(device as c_int, &mut err) } != 0 {
        return Err(take_error(err));
    }
    let pairs = layout_pairs(num_columns);
    if unsafe { zkm_b200_layout_check(pairs.as_ptr(), pairs.len() / 2, &mut err) } != 0 {
        return Err(take_error(err));
    }
    Ok(())
}

// ------------------------------------------------------------------------------------------- proof decoding

const PROOF_MAGIC: u64 = 0x464F_4F52_504D_4B5A; // "ZKMPROOF"

struct Reader<'a> {
    buf: &'a [u64],
    pos: usize,
}
impl<'a> Reader<'a> {
    fn u(&mut self) -> Result<u64> {
        let v = *self.buf.get(self.pos).ok_or_else(|| anyhow!("proof buffer truncated"))?;
        self.pos += 1;
        Ok(v)
    }
    fn words(&mut self, n: usize) -> Result<&'a [u64]> {
        // lengths come from the buffer: neither pos + n nor n * unit may be trusted not to wrap
        let end = self.pos.checked_add(n).filter(|&e| e <= self.buf.len()).ok_or_else(|| anyhow!("proof buffer truncated"))?;
        let s = &self.buf[self.pos..end];
        self.pos = end;
        Ok(s)
    }
    /// "vec X": a length word, then `unit` words per item.
    fn vec(&mut self, unit: usize) -> Result<&'a [u64]> {
        let n = self.u()? as usize;
        self.words(n.checked_mul(unit).ok_or_else(|| anyhow!("proof buffer truncated"))?)
    }
    /// A number of items that take at least one word each: bounded by what is left of the buffer, hence safe as a capacity.
    fn count(&mut self) -> Result<usize> {
        let n = self.u()? as usize;
        ensure!(n <= self.buf.len() - self.pos, "proof buffer truncated");
        Ok(n)
    }
}

fn f(x: u64) -> F {
    F::from_canonical_u64(x) // the library emits canonical residues only
}
fn fs(w: &[u64]) -> Vec<F> {
    w.iter().map(|&x| f(x)).collect()
}
fn exts(w: &[u64]) -> Vec<FE> {
    w.chunks_exact(2).map(|c| FE::from_basefield_array([f(c[0]), f(c[1])])).collect()
}
fn hashes(w: &[u64]) -> Vec<HashOut<F>> {
    w.chunks_exact(4).map(|c| HashOut { elements: [f(c[0]), f(c[1]), f(c[2]), f(c[3])] }).collect()
}
fn cap(w: &[u64]) -> MerkleCap<F, H> {
    MerkleCap(hashes(w))
}
fn path(w: &[u64]) -> MerkleProof<F, H> {
    MerkleProof { siblings: hashes(w) }
}

fn decode_stark_proof(r: &mut Reader) -> Result<StarkProofWithMetadata<F, C, D>> {
    // layout: include/zkm_b200.h "Proof buffer layout"; struct fields: proof.rs:178-201,283-296
    let init_challenger_state = PoseidonPermutation::<F>::new(fs(r.words(12)?));
    let trace_cap = cap(r.vec(4)?);
    let auxiliary_polys_cap = cap(r.vec(4)?);
    let quotient_polys_cap = cap(r.vec(4)?);
    let openings = StarkOpeningSet {
        local_values: exts(r.vec(2)?),
        next_values: exts(r.vec(2)?),
        auxiliary_polys: exts(r.vec(2)?),
        auxiliary_polys_next: exts(r.vec(2)?),
        ctl_zs_first: fs(r.vec(1)?),
        quotient_polys: exts(r.vec(2)?),
    };
    let ncaps = r.count()?;
    let mut commit_phase_merkle_caps = Vec::with_capacity(ncaps);
    for _ in 0..ncaps {
        commit_phase_merkle_caps.push(cap(r.vec(4)?));
    }
    let nq = r.count()?;
    let mut query_round_proofs = Vec::with_capacity(nq);
    for _ in 0..nq {
        let noracles = r.count()?;
        let mut evals_proofs = Vec::with_capacity(noracles);
        for _ in 0..noracles {
            let leaf = fs(r.vec(1)?);
            evals_proofs.push((leaf, path(r.vec(4)?)));
        }
        let nsteps = r.count()?;
        let mut steps = Vec::with_capacity(nsteps);
        for _ in 0..nsteps {
            let evals = exts(r.vec(2)?);
            steps.push(FriQueryStep { evals, merkle_proof: path(r.vec(4)?) });
        }
        query_round_proofs.push(FriQueryRound { initial_trees_proof: FriInitialTreeProof { evals_proofs }, steps });
    }
    let final_poly = PolynomialCoeffs::new(exts(r.vec(2)?));
    let pow_witness = f(r.u()?);
    Ok(StarkProofWithMetadata {
        init_challenger_state,
        proof: StarkProof {
            trace_cap,
            auxiliary_polys_cap,
            quotient_polys_cap,
            openings,
            opening_proof: FriProof { commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness },
        },
    })
}

/// Rebuilds `AllProof` (proof.rs:25-29) from the library's flat little-endian u64 buffer.
pub fn decode_all_proof(buf: &[u64]) -> Result<AllProof<F, C, D>> {
    let mut r = Reader { buf, pos: 0 };
    ensure!(r.u()? == PROOF_MAGIC, "bad proof magic");
    ensure!(r.u()? == 1, "bad proof version");
    ensure!(r.u()? as usize == NUM_TABLES, "proof is not an AllStark proof");
    let nch = r.count()?;
    let mut challenges = Vec::with_capacity(nch);
    for _ in 0..nch {
        let beta = f(r.u()?);
        let gamma = f(r.u()?);
        challenges.push(GrandProductChallenge { beta, gamma }); // cross_table_lookup.rs:486-491
    }
    let mut roots_before = [0u32; 8];
    let mut roots_after = [0u32; 8];
    for x in roots_before.iter_mut() {
        *x = r.u()? as u32;
    }
    for x in roots_after.iter_mut() {
        *x = r.u()? as u32;
    }
    let userdata: Vec<u8> = r.vec(1)?.iter().map(|&b| b as u8).collect();
    let mut proofs = Vec::with_capacity(NUM_TABLES);
    for _ in 0..NUM_TABLES {
        proofs.push(decode_stark_proof(&mut r)?);
    }
    ensure!(r.pos == buf.len(), "trailing data after proof");
    let stark_proofs: [StarkProofWithMetadata<F, C, D>; NUM_TABLES] =
        proofs.try_into().map_err(|_| anyhow!("wrong number of table proofs"))?;
    Ok(AllProof {
        stark_proofs,
        ctl_challenges: GrandProductChallengeSet { challenges },
        public_values: PublicValues {
            roots_before: MemRoots { root: roots_before },
            roots_after: MemRoots { root: roots_after },
            userdata,
        },
    })
}

// ------------------------------------------------------------------------------------------- the replacement call

pub(crate) fn stark_config_to_c(config: &StarkConfig) -> Result<ZkmStarkConfig> {
    use plonky2::fri::reduction_strategies::FriReductionStrategy;
    let fc = &config.fri_config; // config.rs:4-29
    let (arity_bits, final_poly_bits) = match &fc.reduction_strategy {
        FriReductionStrategy::ConstantArityBits(a, b) => (*a as u32, *b as u32),
        other => bail!("zkm_b200 supports ConstantArityBits only, got {:?}", other),
    };
    Ok(ZkmStarkConfig {
        rate_bits: fc.rate_bits as u32,
        cap_height: fc.cap_height as u32,
        pow_bits: fc.proof_of_work_bits,
        num_queries: fc.num_query_rounds as u32,
        num_challenges: config.num_challenges as u32,
        arity_bits,
        final_poly_bits,
    })
}

/// `prove_with_traces` (prover.rs:130-140) for F = GoldilocksField, C = PoseidonGoldilocksConfig, D = 2 on the B200 library.
pub fn prove_with_traces_b200(
    config: &StarkConfig,
    mut trace_poly_values: [Vec<PolynomialValues<F>>; NUM_TABLES],
    public_values: PublicValues,
) -> Result<AllProof<F, C, D>> {
    // `GoldilocksField(u64)` may hold non-canonical representatives; the ABI takes canonical words
    for table in trace_poly_values.iter_mut() {
        for col in table.iter_mut() {
            for x in col.values.iter_mut() {
                *x = F::from_canonical_u64(x.to_canonical_u64());
            }
        }
    }
    let col_ptrs: Vec<Vec<*const u64>> =
        trace_poly_values.iter().map(|t| t.iter().map(|p| p.values.as_ptr() as *const u64).collect()).collect();
    let tables: Vec<ZkmTable> = col_ptrs
        .iter()
        .zip(trace_poly_values.iter())
        .map(|(c, t)| ZkmTable { cols: c.as_ptr(), ncols: c.len() as u32, log_n: log2_strict(t[0].len()) as u32 })
        .collect();
    let cfg = stark_config_to_c(config)?;
    let (mut out, mut words, mut err): (*mut u64, usize, *mut c_char) = (core::ptr::null_mut(), 0, core::ptr::null_mut());
    let rc = unsafe {
        zkm_b200_prove_with_traces(
            tables.as_ptr(), public_values.roots_before.root.as_ptr(), public_values.roots_after.root.as_ptr(),
            public_values.userdata.as_ptr(), public_values.userdata.len() as u32, &cfg, &mut out, &mut words, &mut err,
        )
    };
    if rc != 0 {
        return Err(take_error(err));
    }
    let proof = decode_all_proof(unsafe { std::slice::from_raw_parts(out, words) });
    unsafe { zkm_b200_free(out as *mut c_void) };
    let proof = proof?;
    ensure!(proof.public_values.userdata == public_values.userdata, "public values were not echoed back");
    Ok(proof)
}

/// Generic front used by the cfg(feature = "b200") body of `prove_with_traces<F, C, D>`: the only instantiation in the tree is
/// (GoldilocksField, PoseidonGoldilocksConfig, 2) (prover/examples/utils/src/utils.rs:34-36, recursion/src/lib.rs:25-27); any
/// other one is refused.  The casts go through `dyn Any`, i.e. they are checked at run time and need no `unsafe`.
pub fn prove_with_traces_generic<F2, C2, const D2: usize>(
    config: &StarkConfig,
    traces: [Vec<PolynomialValues<F2>>; NUM_TABLES],
    public_values: PublicValues,
) -> Result<AllProof<F2, C2, D2>>
where
    F2: RichField + Extendable<D2>,
    C2: GenericConfig<D2, F = F2> + 'static,
{
    let traces: Box<dyn Any> = Box::new(traces);
    let traces = traces
        .downcast::<[Vec<PolynomialValues<F>>; NUM_TABLES]>()
        .map_err(|_| anyhow!("zkm_b200 proves over GoldilocksField / PoseidonGoldilocksConfig / D = 2 only"))?;
    let proof: Box<dyn Any> = Box::new(prove_with_traces_b200(config, *traces, public_values)?);
    proof
        .downcast::<AllProof<F2, C2, D2>>()
        .map(|b| *b)
        .map_err(|_| anyhow!("zkm_b200 proves over GoldilocksField / PoseidonGoldilocksConfig / D = 2 only"))
}

/// Device-time scopes of the last proof, keyed by the reference's TimingTree strings (a `TimingTree` measures wall clock
/// itself and cannot be handed durations, so they are logged next to it).  Lines: "<depth>\t<ms>\t<scope>".
pub fn log_last_timing() {
    let p = unsafe { zkm_b200_last_timing() };
    if p.is_null() {
        return;
    }
    for line in unsafe { CStr::from_ptr(p) }.to_string_lossy().lines() {
        let mut it = line.splitn(3, '\t');
        if let (Some(depth), Some(ms), Some(name)) = (it.next(), it.next(), it.next()) {
            let indent = "| ".repeat(depth.parse::<usize>().unwrap_or(0));
            log::debug!("{}{:.4}s to {} [B200 device time]", indent, ms.parse::<f64>().unwrap_or(0.0) / 1e3, name);
        }
    }
    unsafe { zkm_b200_free_string(p) };
}
pub fn enable_timing(on: bool) {
    unsafe { zkm_b200_timing_enable(on as c_int) }
}
