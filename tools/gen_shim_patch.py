#!/usr/bin/env python
"""Regenerates shim/prover_b200.patch and shim/emulator_b200.patch as real unified diffs against the reference checkout
(/root/reference, zkMIPS/zkm @ 04117ce3): the reference files are copied to a scratch directory, the edits below are applied to the
copies (every edit is anchored on a line of the reference and fails loudly when the anchor is gone), and `diff -u` writes the
patches.  tests/test_shim_patch.py applies the committed patches to a fresh copy with patch(1) and checks the result.

    python tools/gen_shim_patch.py [--reference /root/reference]

The new source files (shim/src/*.rs) are not part of the patches; they are copied next to the patched files:
    shim/src/b200.rs, shim/src/b200_ops.rs -> prover/src/        shim/src/b200_split.rs -> emulator/src/
"""
import argparse
import pathlib
import shutil
import subprocess
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent

PROVE_TWIN = '''/// Compute all STARK proofs.
#[cfg(feature = "b200")]
pub(crate) fn prove_with_traces<F, C, const D: usize>(
    _all_stark: &AllStark<F, D>,
    config: &StarkConfig,
    trace_poly_values: [Vec<PolynomialValues<F>>; NUM_TABLES],
    public_values: PublicValues,
    timing: &mut TimingTree,
) -> Result<AllProof<F, C, D>>
where
    F: RichField + Extendable<D>,
    C: GenericConfig<D, F = F> + 'static,
{
    // everything the function below does (trace commitments, CTL data, the 12 table proofs) runs on the GPU
    let proof = timed!(
        timing,
        "compute all proofs on B200",
        crate::b200::prove_with_traces_generic::<F, C, D>(config, trace_poly_values, public_values)?
    );
    crate::b200::log_last_timing();
    Ok(proof)
}

/// Compute all STARK proofs.
#[cfg(not(feature = "b200"))]
'''


def replace_once(text, old, new, what):
    assert text.count(old) >= 1, f"anchor not found: {what}"
    return text.replace(old, new, 1)


def edit_prover(d):
    p = d / "prover/Cargo.toml"
    p.write_text(replace_once(p.read_text(), "[features]\ntest = []\n",
                              "[features]\ntest = []\n# prove_with_traces served by libzkm_b200.so (build.rs: cargo:rustc-link-lib=dylib=zkm_b200, dylib=cudart)\nb200 = []\n",
                              "prover/Cargo.toml [features]"))
    p = d / "prover/src/lib.rs"
    p.write_text(replace_once(p.read_text(), "pub mod arithmetic;\n",
                              "pub mod arithmetic;\n#[cfg(feature = \"b200\")]\npub mod b200;\n#[cfg(feature = \"b200\")]\npub mod b200_ops;\n", "prover/src/lib.rs mod list"))
    p = d / "prover/src/prover.rs"
    s = p.read_text()
    # the dyn-Any casts of b200::prove_with_traces_generic need C: 'static; the callers' bounds follow (F: Field is 'static already,
    # every GenericConfig in the tree is a marker type, and fixed_recursive_verifier.rs:310 asks for the same bound)
    head, sep, tail = s.partition("fn prove_with_commitments<")
    assert sep, "anchor not found: prove_with_commitments"
    n = head.count("    C: GenericConfig<D, F = F>,\n")
    assert n == 5, f"expected the five entry points above prove_with_commitments, found {n}"
    head = head.replace("    C: GenericConfig<D, F = F>,\n", "    C: GenericConfig<D, F = F> + 'static,\n")
    s = head + sep + tail
    s = replace_once(s, "/// Compute all STARK proofs.\npub(crate) fn prove_with_traces<", PROVE_TWIN + "pub(crate) fn prove_with_traces<",
                     "prover/src/prover.rs prove_with_traces")
    p.write_text(s)
    # b200_ops.rs reads the three private fields of logic::Operation
    p = d / "prover/src/logic.rs"
    s = p.read_text()
    for f in ("operator: Op,", "input0: u32,", "input1: u32,"):
        s = replace_once(s, "    " + f + "\n", "    pub(crate) " + f + "\n", "logic::Operation." + f)
    p.write_text(s)


def edit_emulator(d):
    p = d / "emulator/src/lib.rs"
    p.write_text(replace_once(p.read_text(), "pub mod memory;\n", "#[cfg(feature = \"b200\")]\npub mod b200_split;\npub mod memory;\n", "emulator/src/lib.rs"))
    p = d / "emulator/Cargo.toml"
    p.write_text(replace_once(p.read_text(), "[features]\ntest = []\n", "[features]\ntest = []\nb200 = []    # page hashing of split_segment on the GPU (libzkm_b200.so)\n",
                              "emulator/Cargo.toml [features]"))
    p = d / "emulator/src/memory.rs"
    s = p.read_text()
    s = replace_once(s, "    pages: BTreeMap<u32, Rc<RefCell<CachedPage>>>,\n", "    pub(crate) pages: BTreeMap<u32, Rc<RefCell<CachedPage>>>,\n", "Memory.pages")
    s = replace_once(s, "    rtrace: BTreeMap<u32, [u8; PAGE_SIZE]>,\n", "    pub(crate) rtrace: BTreeMap<u32, [u8; PAGE_SIZE]>,\n", "Memory.rtrace")
    s = replace_once(s, "    wtrace: [BTreeMap<u32, Rc<RefCell<CachedPage>>>; 3],\n", "    pub(crate) wtrace: [BTreeMap<u32, Rc<RefCell<CachedPage>>>; 3],\n", "Memory.wtrace")
    p.write_text(s)
    p = d / "emulator/src/state.rs"
    s = p.read_text()
    for f in ("pre_pc: u32,", "pre_image_id: [u8; 32],", "pre_hash_root: [u8; 32],", "pre_input: Vec<Vec<u8>>,", "pre_input_ptr: usize,",
              "pre_public_values: Vec<u8>,"):
        s = replace_once(s, "    " + f + "\n", "    pub(crate) " + f + "\n", "InstrumentedState." + f)
    s = replace_once(s, "    pre_public_values_ptr: usize,\n}\n",
                     "    pub(crate) pre_public_values_ptr: usize,\n    /// zkm_splitter_t of libzkm_b200 (b200_split.rs): hash pages and image ids on the GPU\n"
                     "    #[cfg(feature = \"b200\")]\n    pub(crate) splitter: *mut std::ffi::c_void,\n}\n", "InstrumentedState.pre_public_values_ptr")
    s = replace_once(s, "            pre_public_values_ptr: 0,\n        })\n",
                     "            pre_public_values_ptr: 0,\n            #[cfg(feature = \"b200\")]\n            splitter: crate::b200_split::splitter_create(),\n        })\n",
                     "InstrumentedState::new")
    # split_segment keeps its name and signature: the original body moves under cfg(not(b200)), the twin forwards
    s = replace_once(s, "    /// the caller should provide a write to write segment if proof is true\n    pub fn split_segment<W: Write>(\n",
                     "    /// the caller should provide a write to write segment if proof is true\n    #[cfg(feature = \"b200\")]\n"
                     "    pub fn split_segment<W: Write>(&mut self, proof: bool, output: &str, new_writer: fn(&str) -> Option<W>) {\n"
                     "        self.split_segment_b200(proof, output, new_writer)\n    }\n\n"
                     "    /// the caller should provide a write to write segment if proof is true\n    #[cfg(not(feature = \"b200\"))]\n    pub fn split_segment<W: Write>(\n",
                     "InstrumentedState::split_segment")
    p.write_text(s)


def diff(a, b, files, header):
    out = [header]
    for f in files:
        r = subprocess.run(["diff", "-u", "--label", f"a/{f}", "--label", f"b/{f}", str(a / f), str(b / f)], capture_output=True, text=True)
        assert r.returncode == 1, f"{f}: no difference or diff failed\n{r.stderr}"
        out.append(r.stdout)
    return "".join(out)


PROVER_FILES = ["prover/Cargo.toml", "prover/src/lib.rs", "prover/src/logic.rs", "prover/src/prover.rs"]
EMULATOR_FILES = ["emulator/Cargo.toml", "emulator/src/lib.rs", "emulator/src/memory.rs", "emulator/src/state.rs"]


def generate(reference):
    with tempfile.TemporaryDirectory() as tmp:
        a, b = pathlib.Path(tmp) / "a", pathlib.Path(tmp) / "b"
        for f in PROVER_FILES + EMULATOR_FILES:
            for side in (a, b):
                (side / f).parent.mkdir(parents=True, exist_ok=True)
                shutil.copy(reference / f, side / f)
        edit_prover(b)
        edit_emulator(b)
        prover = diff(a, b, PROVER_FILES,
                      "# zkMIPS/zkm @ 04117ce3, `patch -p1` from the repository root (generated by tools/gen_shim_patch.py; applied by\n"
                      "# tests/test_shim_patch.py, never compiled in this image: no cargo).  Adds the `b200` feature: prove_with_traces is served by\n"
                      "# libzkm_b200.so; copy shim/src/b200.rs and shim/src/b200_ops.rs to prover/src/.  verifier.rs is untouched; the five proving\n"
                      "# entry points gain `C: 'static` (needed by the checked casts of b200::prove_with_traces_generic).\n"
                      "# One-time set-up in the host binary (e.g. prover/examples/utils/src/utils.rs:51, next to AllStark::default()):\n"
                      "#     zkm_prover::b200::init(0, [NUM_ARITH_COLUMNS, NUM_CPU_COLUMNS, poseidon NUM_COLUMNS, NUM_POSEIDON_SPONGE_COLUMNS, keccak NUM_COLUMNS,\n"
                      "#                                NUM_KECCAK_SPONGE_COLUMNS, NUM_SHA_EXTEND_COLUMNS, NUM_SHA_EXTEND_SPONGE_COLUMNS, NUM_SHA_COMPRESS_COLUMNS,\n"
                      "#                                NUM_SHA_COMPRESS_SPONGE_COLUMNS, logic NUM_COLUMNS, memory NUM_COLUMNS])?;\n")
        emulator = diff(a, b, EMULATOR_FILES,
                        "# zkMIPS/zkm @ 04117ce3, `patch -p1` from the repository root (generated by tools/gen_shim_patch.py).  Adds the `b200` feature to\n"
                        "# the emulator: InstrumentedState::split_segment hashes the segment's pages on the GPU; copy shim/src/b200_split.rs to\n"
                        "# emulator/src/.  A state loaded from a segment file also calls b200_split::splitter_seed(self.splitter, &self.state.memory)\n"
                        "# before its first split.\n")
    return prover, emulator


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    prover, emulator = generate(pathlib.Path(args.reference))
    (ROOT / "shim/prover_b200.patch").write_text(prover)
    (ROOT / "shim/emulator_b200.patch").write_text(emulator)
    print(f"prover_b200.patch {len(prover.splitlines())} lines, emulator_b200.patch {len(emulator.splitlines())} lines")
