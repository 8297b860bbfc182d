// Transcription check on the REFERENCE side (SOURCE ONLY -- never compiled here: no cargo in this image).
//
// Append to `prover/src/cross_table_lookup.rs` (it reads the private fields of `TableWithColumns`) and run
//     cargo test --release -p zkm-prover b200_fingerprints -- --nocapture
// It prints, in the format of tests/golden/constraint_fingerprints_v1.json of the zkm_b200 repository: per table the
// alpha-fold of ALL constraints of `eval_packed_generic` on one fixed pseudo-random frame (sensitive to every constraint, its
// coefficients and its POSITION in the emission order), per in-table lookup and per cross-table lookup a fold of its column and
// filter evaluations.  `python tools/compare_fingerprints.py <printed file>` in that repository diffs the two.
#[cfg(test)]
mod b200_fingerprints {
    use plonky2::field::goldilocks_field::GoldilocksField;
    use plonky2::field::types::{Field, PrimeField64};

    use super::*;
    use crate::all_stark::{all_cross_table_lookups, AllStark, Table};
    use crate::constraint_consumer::ConstraintConsumer;
    use crate::evaluation_frame::StarkEvaluationFrame;
    use crate::stark::Stark;

    type F = GoldilocksField;
    const D: usize = 2;
    const SEED: u64 = 0x5EED_F1A6_0000_0000;
    const A0: u64 = 0x9E37_79B9_7F4A_7C15 % 0xFFFF_FFFF_0000_0001;
    const A1: u64 = 0xC2B2_AE3D_27D4_EB4F % 0xFFFF_FFFF_0000_0001;

    fn cell(x: u64) -> F {
        let mut z = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^= z >> 31;
        F::from_noncanonical_u64(z) // = z % p
    }
    fn frame(seed: u64, ncols: usize) -> (Vec<F>, Vec<F>) {
        ((0..ncols).map(|c| cell(seed + 2 * c as u64)).collect(), (0..ncols).map(|c| cell(seed + 2 * c as u64 + 1)).collect())
    }
    fn table_fp<S: Stark<F, D>>(stark: &S, t: usize) -> [u64; 2] {
        let n = S::COLUMNS;
        let (lv, nv) = frame(SEED + 0x10000 * t as u64, n);
        let vars = S::EvaluationFrame::<F, F, 1>::from_values(&lv, &nv);
        let mut cc = ConstraintConsumer::<F>::new(
            vec![F::from_canonical_u64(A0), F::from_canonical_u64(A1)], F::from_canonical_u64(3), F::from_canonical_u64(5),
            F::from_canonical_u64(7));
        stark.eval_packed_generic(&vars, &mut cc);
        let acc = cc.accumulators();
        [acc[0].to_canonical_u64(), acc[1].to_canonical_u64()]
    }
    fn lookup_fps<S: Stark<F, D>>(stark: &S, t: usize) -> Vec<(usize, u64)> {
        let (lv, nv) = frame(SEED + 0x10000 * t as u64, S::COLUMNS);
        stark.lookups().iter().map(|l| {
            let a0 = F::from_canonical_u64(A0);
            let mut acc = F::ZERO;
            for c in &l.columns { acc = acc * a0 + c.eval_with_next::<F, F, 1>(&lv, &nv); }
            acc = acc * a0 + l.table_column.eval_with_next::<F, F, 1>(&lv, &nv);
            acc = acc * a0 + l.frequencies_column.eval_with_next::<F, F, 1>(&lv, &nv);
            for f in &l.filter_columns {
                acc = acc * a0 + f.as_ref().map(|f| f.eval_filter::<F, F, 1>(&lv, &nv)).unwrap_or(F::ONE);
            }
            (l.columns.len(), acc.to_canonical_u64())
        }).collect()
    }
    fn twc_fp(t: &TableWithColumns<F>, ncols: &[usize; 12]) -> (usize, usize, u64) {
        let ti = t.table as usize;
        let (lv, nv) = frame(SEED + 0x10000 * ti as u64, ncols[ti]);
        let a0 = F::from_canonical_u64(A0);
        let mut acc = t.filter.as_ref().map(|f| f.eval_filter::<F, F, 1>(&lv, &nv)).unwrap_or(F::ONE);
        for c in &t.columns { acc = acc * a0 + c.eval_with_next::<F, F, 1>(&lv, &nv); }
        (ti, t.columns.len(), acc.to_canonical_u64())
    }

    #[test]
    fn b200_fingerprints() {
        let s = AllStark::<F, D>::default();
        macro_rules! each { ($f:ident) => { vec![
            $f(&s.arithmetic_stark, 0), $f(&s.cpu_stark, 1), $f(&s.poseidon_stark, 2), $f(&s.poseidon_sponge_stark, 3),
            $f(&s.keccak_stark, 4), $f(&s.keccak_sponge_stark, 5), $f(&s.sha_extend_stark, 6), $f(&s.sha_extend_sponge_stark, 7),
            $f(&s.sha_compress_stark, 8), $f(&s.sha_compress_sponge_stark, 9), $f(&s.logic_stark, 10), $f(&s.memory_stark, 11)] } }
        let names = ["Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge", "ShaExtend", "ShaExtendSponge",
                     "ShaCompress", "ShaCompressSponge", "Logic", "Memory"];
        for (t, acc) in each!(table_fp).iter().enumerate() {
            println!("table {} acc {} {}", names[t], acc[0], acc[1]);
        }
        for (t, ls) in each!(lookup_fps).iter().enumerate() {
            for (i, (n, fp)) in ls.iter().enumerate() { println!("lookup {} {} num_columns {} fp {}", names[t], i, n, fp); }
        }
        let ncols: [usize; 12] = [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13];
        for (i, ctl) in all_cross_table_lookups::<F>().iter().enumerate() {
            for (e, t) in ctl.looking_tables.iter().chain(std::iter::once(&ctl.looked_table)).enumerate() {
                let (ti, n, fp) = twc_fp(t, &ncols);
                let role = if e < ctl.looking_tables.len() { "looking" } else { "looked" };
                println!("ctl {} entry {} {} {} num_columns {} fp {}", i, e, role, names[ti], n, fp);
            }
        }
        let _ = Table::all();
    }
}
