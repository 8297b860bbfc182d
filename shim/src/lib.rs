//! Rust shim for zkm_b200 (SOURCE ONLY: this image has no cargo/rustc).
//! Decodes the flat proof buffer of include/zkm_b200.h into the reference's proof structs.
//! Field/struct names follow reference prover/src/proof.rs:25-29,178-201,283-296 and plonky2 0.1.4.
#![allow(dead_code)]

pub struct Reader<'a> { buf: &'a [u64], pos: usize }

impl<'a> Reader<'a> {
    pub fn new(buf: &'a [u64]) -> Self { Self { buf, pos: 0 } }
    pub fn u(&mut self) -> Result<u64, String> {
        let v = *self.buf.get(self.pos).ok_or("proof buffer truncated")?;
        self.pos += 1;
        Ok(v)
    }
    pub fn words(&mut self, n: usize) -> Result<&'a [u64], String> {
        let s = self.buf.get(self.pos..self.pos + n).ok_or("proof buffer truncated")?;
        self.pos += n;
        Ok(s)
    }
    /// vec of `unit`-word items: length word, then the items
    pub fn vec(&mut self, unit: usize) -> Result<&'a [u64], String> { let n = self.u()? as usize; self.words(n * unit) }
}

pub const PROOF_MAGIC: u64 = 0x464F4F52504D4B5A; // "ZKMPROOF"

/// Plain-data mirror of one StarkProofWithMetadata; the zkm-side glue converts these slices with
/// `GoldilocksField::from_canonical_u64`, `QuadraticExtension([a, b])`, `HashOut { elements }`,
/// `MerkleCap(Vec<HashOut>)`, `MerkleProof { siblings }`, `FriQueryStep { evals, merkle_proof }`,
/// `FriInitialTreeProof { evals_proofs }`, `FriQueryRound { initial_trees_proof, steps }`.
pub struct RawQueryRound<'a> { pub initial: Vec<(&'a [u64], &'a [u64])>, pub steps: Vec<(&'a [u64], &'a [u64])> }
pub struct RawStarkProof<'a> {
    pub init_challenger_state: &'a [u64],
    pub trace_cap: &'a [u64], pub auxiliary_polys_cap: &'a [u64], pub quotient_polys_cap: &'a [u64],
    pub local_values: &'a [u64], pub next_values: &'a [u64], pub auxiliary_polys: &'a [u64],
    pub auxiliary_polys_next: &'a [u64], pub ctl_zs_first: &'a [u64], pub quotient_polys: &'a [u64],
    pub commit_phase_merkle_caps: Vec<&'a [u64]>, pub query_round_proofs: Vec<RawQueryRound<'a>>,
    pub final_poly: &'a [u64], pub pow_witness: u64,
}
pub struct RawAllProof<'a> {
    pub ctl_challenges: Vec<(u64, u64)>, pub roots_before: [u32; 8], pub roots_after: [u32; 8], pub userdata: Vec<u8>,
    pub stark_proofs: Vec<RawStarkProof<'a>>,
}

pub fn decode(buf: &[u64]) -> Result<RawAllProof<'_>, String> {
    let mut r = Reader::new(buf);
    if r.u()? != PROOF_MAGIC { return Err("bad proof magic".into()); }
    if r.u()? != 1 { return Err("bad proof version".into()); }
    let num_tables = r.u()? as usize;
    let nc = r.u()? as usize;
    let mut ctl_challenges = Vec::new();
    for _ in 0..nc { let b = r.u()?; let g = r.u()?; ctl_challenges.push((b, g)); }
    let mut roots_before = [0u32; 8];
    let mut roots_after = [0u32; 8];
    for x in roots_before.iter_mut() { *x = r.u()? as u32; }
    for x in roots_after.iter_mut() { *x = r.u()? as u32; }
    let userdata = r.vec(1)?.iter().map(|&b| b as u8).collect();
    let mut stark_proofs = Vec::new();
    for _ in 0..num_tables {
        let init_challenger_state = r.words(12)?;
        let trace_cap = r.vec(4)?; let auxiliary_polys_cap = r.vec(4)?; let quotient_polys_cap = r.vec(4)?;
        let local_values = r.vec(2)?; let next_values = r.vec(2)?; let auxiliary_polys = r.vec(2)?;
        let auxiliary_polys_next = r.vec(2)?; let ctl_zs_first = r.vec(1)?; let quotient_polys = r.vec(2)?;
        let ncaps = r.u()? as usize;
        let mut commit_phase_merkle_caps = Vec::new();
        for _ in 0..ncaps { commit_phase_merkle_caps.push(r.vec(4)?); }
        let nq = r.u()? as usize;
        let mut query_round_proofs = Vec::new();
        for _ in 0..nq {
            let no = r.u()? as usize;
            let mut initial = Vec::new();
            for _ in 0..no { let leaf = r.vec(1)?; let path = r.vec(4)?; initial.push((leaf, path)); }
            let ns = r.u()? as usize;
            let mut steps = Vec::new();
            for _ in 0..ns { let evals = r.vec(2)?; let path = r.vec(4)?; steps.push((evals, path)); }
            query_round_proofs.push(RawQueryRound { initial, steps });
        }
        let final_poly = r.vec(2)?;
        let pow_witness = r.u()?;
        stark_proofs.push(RawStarkProof { init_challenger_state, trace_cap, auxiliary_polys_cap, quotient_polys_cap,
            local_values, next_values, auxiliary_polys, auxiliary_polys_next, ctl_zs_first, quotient_polys,
            commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness });
    }
    if r.pos != buf.len() { return Err("trailing data after proof".into()); }
    Ok(RawAllProof { ctl_challenges, roots_before, roots_after, userdata, stark_proofs })
}
