"""shim/*.patch are real unified diffs: applied with patch(1) to a scratch copy of the reference checkout they must go in without
fuzz or rejects and leave the edits the Rust sources in shim/src/ rely on.  Also: the committed patches are what
tools/gen_shim_patch.py generates.  Needs /root/reference (present where the CPU suite runs; skipped elsewhere); cargo is not
available, so compiling the result remains the maintainer's step."""
import pathlib
import shutil
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF = pathlib.Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REF / "prover/src/prover.rs").exists(), reason="reference checkout not present")


def test_patches_apply_to_the_reference(tmp_path):
    for sub in ("prover", "emulator"):
        (tmp_path / sub).mkdir()
        shutil.copy(REF / sub / "Cargo.toml", tmp_path / sub / "Cargo.toml")
        shutil.copytree(REF / sub / "src", tmp_path / sub / "src")
    for name in ("prover_b200.patch", "emulator_b200.patch"):
        r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", str(ROOT / "shim" / name)], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "fuzz" not in r.stdout and "offset" not in r.stdout and not list(tmp_path.rglob("*.rej")), r.stdout
    for src, dst in (("b200.rs", "prover/src"), ("b200_ops.rs", "prover/src"), ("b200_split.rs", "emulator/src")):
        shutil.copy(ROOT / "shim/src" / src, tmp_path / dst / src)
    prover = (tmp_path / "prover/src/prover.rs").read_text()
    assert prover.count('#[cfg(feature = "b200")]\npub(crate) fn prove_with_traces<') == 1
    assert prover.count('#[cfg(not(feature = "b200"))]\npub(crate) fn prove_with_traces<') == 1
    assert "crate::b200::prove_with_traces_generic::<F, C, D>(config, trace_poly_values, public_values)?" in prover
    assert prover.count("C: GenericConfig<D, F = F> + 'static,") == 6          # five entry points + the twin
    lib = (tmp_path / "prover/src/lib.rs").read_text()
    assert '#[cfg(feature = "b200")]\npub mod b200;' in lib and '#[cfg(feature = "b200")]\npub mod b200_ops;' in lib
    assert "b200 = []" in (tmp_path / "prover/Cargo.toml").read_text() and "b200 = []" in (tmp_path / "emulator/Cargo.toml").read_text()
    # what the shim sources reach into exists with the visibility they need
    logic = (tmp_path / "prover/src/logic.rs").read_text()
    for f in ("operator: Op", "input0: u32", "input1: u32"):
        assert f"pub(crate) {f}," in logic
    memory, state = (tmp_path / "emulator/src/memory.rs").read_text(), (tmp_path / "emulator/src/state.rs").read_text()
    for f in ("pages:", "rtrace:", "wtrace:"):
        assert f"pub(crate) {f}" in memory
    split = (ROOT / "shim/src/b200_split.rs").read_text()
    for f in ("pre_pc", "pre_image_id", "pre_hash_root", "pre_input", "pre_input_ptr", "pre_public_values", "pre_public_values_ptr", "splitter"):
        assert f"pub(crate) {f}:" in state and f"self.{f}" in split, f
    assert "splitter: crate::b200_split::splitter_create()," in state and "self.split_segment_b200(proof, output, new_writer)" in state
    assert state.count("pub fn split_segment<W: Write>(") == 2 and '#[cfg(not(feature = "b200"))]\n    pub fn split_segment<W: Write>(' in state
    # every `use crate::...` of the shim sources names a module / item that exists in the patched tree
    for src, crate_dir in (("b200.rs", "prover/src"), ("b200_ops.rs", "prover/src"), ("b200_split.rs", "emulator/src")):
        import re
        text = (ROOT / "shim/src" / src).read_text()
        for mod in set(re.findall(r"use crate::(\w+)", text)) | set(re.findall(r"\bcrate::(\w+)::", text)):
            d = tmp_path / crate_dir
            assert (d / f"{mod}.rs").exists() or (d / mod / "mod.rs").exists(), f"{src}: crate::{mod}"


def test_committed_patches_are_the_generated_ones():
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_shim_patch
    prover, emulator = gen_shim_patch.generate(REF)
    assert prover == (ROOT / "shim/prover_b200.patch").read_text()
    assert emulator == (ROOT / "shim/emulator_b200.patch").read_text()


def _crate_uses(text):
    """`use crate::a::b::{X, Y as Z};` / `use crate::a::X;` -> [(['a', 'b'], 'X'), ...]"""
    import re
    out = []
    for path, group, single in re.findall(r"use crate::((?:\w+::)*)(?:\{([^}]*)\}|(\w+)(?:\s+as\s+\w+)?)\s*;", text):
        mods = [m for m in path.split("::") if m]
        names = [n.split(" as ")[0].strip() for n in group.split(",")] if group else [single]
        out += [(mods, n) for n in names if n and n != "self"]
    return out


def test_shim_imports_resolve_in_the_reference():
    """Poor man's name resolution (no rustc here): every item the shim sources import from the host crate is defined, or re-exported,
    in the module file the path names, and is not private."""
    import re
    checked = 0
    for src, crate_dir in (("b200.rs", "prover/src"), ("b200_ops.rs", "prover/src"), ("b200_split.rs", "emulator/src")):
        text = (ROOT / "shim/src" / src).read_text()
        for mods, name in _crate_uses(text):
            if mods and mods[0] in ("b200", "b200_ops", "b200_split"):
                target = (ROOT / "shim/src" / f"{mods[0]}.rs").read_text()
            else:
                base = REF / crate_dir
                # the last path element may itself be the item's module (use crate::logic;)
                cands = [base.joinpath(*mods).with_suffix(".rs"), base.joinpath(*mods) / "mod.rs"] if mods else []
                cands += [base / f"{name}.rs", base / name / "mod.rs"] if not mods else []
                files = [c for c in cands if c.exists()]
                assert files, f"{src}: no module file for crate::{'::'.join(mods + [name])}"
                if not mods:
                    checked += 1
                    continue                                  # a module import: the file exists
                target = files[0].read_text()
            pat = rf"^\s*pub(?:\(crate\))?\s+(?:unsafe\s+)?(?:const\s+fn|fn|struct|enum|trait|type|const|static|union)\s+{name}\b|^\s*pub(?:\(crate\))?\s+use\s+[^;]*\b{name}\b"
            assert re.search(pat, target, flags=re.M), f"{src}: crate::{'::'.join(mods)}::{name} is not a visible item of its module"
            checked += 1
    assert checked >= 25


def test_shim_reads_fields_the_reference_structs_have():
    """b200_ops.rs walks `Traces` and its operation records field by field: each field it names exists upstream."""
    import re
    ops = re.sub(r"//[^\n]*", "", (ROOT / "shim/src/b200_ops.rs").read_text())
    traces_rs = (REF / "prover/src/witness/traces.rs").read_text()
    fields = set(re.findall(r"\btraces\.(\w+)", ops))
    assert {"arithmetic_ops", "logic_ops", "memory_ops", "cpu", "poseidon_inputs", "keccak_inputs", "keccak_sponge_ops", "poseidon_sponge_ops",
            "sha_extend_inputs", "sha_extend_sponge_ops", "sha_compress_inputs", "sha_compress_sponge_ops"} <= fields
    for f in fields:
        assert re.search(rf"^\s*pub(?:\(crate\))?\s+{f}\s*:", traces_rs, flags=re.M), f"Traces has no field {f}"
    defs = {"MemoryOp": "prover/src/witness/memory.rs", "MemoryAddress": "prover/src/witness/memory.rs",
            "KeccakSpongeOp": "prover/src/keccak_sponge/keccak_sponge_stark.rs", "PoseidonSpongeOp": "prover/src/poseidon_sponge/poseidon_sponge_stark.rs",
            "ShaExtendSpongeOp": "prover/src/sha_extend_sponge/sha_extend_sponge_stark.rs",
            "ShaCompressSpongeOp": "prover/src/sha_compress_sponge/sha_compress_sponge_stark.rs"}
    used = {"MemoryOp": ["address", "timestamp", "kind", "value", "filter"], "MemoryAddress": ["context", "segment", "virt"],
            "KeccakSpongeOp": ["base_address", "timestamp", "input"], "PoseidonSpongeOp": ["base_address", "timestamp", "input"],
            "ShaExtendSpongeOp": ["base_address", "timestamp", "input", "i", "output_address"],
            "ShaCompressSpongeOp": ["base_address", "timestamp", "input", "w_i_s"]}
    for struct, path in defs.items():
        text = (REF / path).read_text()
        m = re.search(rf"pub(?:\(crate\))?\s+struct\s+{struct}\s*\{{(.*?)\n\}}", text, flags=re.S)
        assert m, struct
        for f in used[struct]:
            assert re.search(rf"\bpub(?:\(crate\))?\s+{f}\s*:", m.group(1)), f"{struct}.{f}"
            assert re.search(rf"\.{f}\b", ops), f"b200_ops.rs no longer reads .{f}"


def _literal_fields(text, name):
    """Field names of every struct literal `name { ... }` in `text` (brace matching, top-level commas)."""
    import re
    out = []
    for m in re.finditer(rf"\b{name}\s*\{{", text):
        depth, i, start = 1, m.end(), m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        body, fields, level, cur = text[start:i - 1], [], 0, ""
        for ch in body:
            if ch in "{([":
                level += 1
            elif ch in "})]":
                level -= 1
            if ch == "," and level == 0:
                fields.append(cur)
                cur = ""
            else:
                cur += ch
        fields.append(cur)
        out.append(sorted(re.match(r"\s*(\w+)", f).group(1) for f in fields if f.strip()))
    return out


def test_shim_struct_literals_name_exactly_the_upstream_fields():
    """A struct literal must list every field: the typed proof structs decode_all_proof builds are compared with their definitions
    in the reference (proof.rs, cross_table_lookup.rs).  The plonky2 structs (FriProof, ...) are not in the tree and stay unchecked."""
    import re
    shim = re.sub(r"//[^\n]*", "", (ROOT / "shim/src/b200.rs").read_text())
    where = {"AllProof": "proof.rs", "StarkProofWithMetadata": "proof.rs", "StarkProof": "proof.rs", "StarkOpeningSet": "proof.rs",
             "PublicValues": "proof.rs", "MemRoots": "proof.rs", "GrandProductChallenge": "cross_table_lookup.rs",
             "GrandProductChallengeSet": "cross_table_lookup.rs"}
    for name, path in where.items():
        text = (REF / "prover/src" / path).read_text()
        m = re.search(rf"pub(?:\(crate\))?\s+struct\s+{name}\b[^{{;]*\{{(.*?)\n\}}", text, flags=re.S)
        assert m, name
        want = sorted(re.findall(r"^\s*pub(?:\(crate\))?\s+(\w+)\s*:", m.group(1), flags=re.M))
        assert want, name
        lits = [f for f in _literal_fields(shim, name)]
        assert lits, f"b200.rs builds no {name}"
        for got in lits:
            assert got == want, f"{name}: shim {got} != reference {want}"


def test_layout_pairs_reads_fields_that_exist_upstream():
    """Every field or view accessor layout_pairs() of shim/src/b200.rs touches is declared in the reference's column structs
    (the files the layout fixture was derived from)."""
    import json
    import re
    shim = re.sub(r"//[^\n]*", "", (ROOT / "shim/src/b200.rs").read_text())
    body = re.search(r"pub fn layout_pairs.*?\n\}", shim, flags=re.S).group(0)
    files = json.loads((ROOT / "tests/golden/column_layout_v1.json").read_text())["reference_files"]
    decls = "\n".join((REF / f).read_text() for f in files)
    rust_methods = {"iter", "min", "unwrap", "enumerate", "push", "extend_from_slice", "into_iter", "flat_map", "collect"}
    names = set(re.findall(r"\.(\w+)", body)) - rust_methods
    assert len(names) >= 55
    for n in sorted(names):
        field = re.search(rf"^\s*pub(?:\(crate\))?\s+{n}\s*:", decls, flags=re.M)
        accessor = re.search(rf"\bfn\s+{n}\s*\(\s*&self", decls)
        assert field or accessor, f"layout_pairs reads .{n}, which the reference's column structs do not declare"


def test_fingerprint_test_source_names_upstream_items():
    """shim/fingerprint_test.rs (the cargo test of INTEGRATION.md section 3) is appended to cross_table_lookup.rs upstream: the struct
    fields and functions it uses are declared there / in the modules it imports."""
    import re
    src = re.sub(r"//[^\n]*", "", (ROOT / "shim/fingerprint_test.rs").read_text())
    P = REF / "prover/src"

    def fields(path, struct):
        m = re.search(rf"struct\s+{struct}\b[^{{;]*\{{(.*?)\n\}}", (P / path).read_text(), flags=re.S)
        assert m, struct
        return set(re.findall(r"^\s*(?:pub(?:\([a-z]+\))?\s+)?(\w+)\s*:", m.group(1), flags=re.M))

    assert set(re.findall(r"\bs\.(\w+)", src)) == fields("all_stark.rs", "AllStark") - {"cross_table_lookups"}
    assert set(re.findall(r"\bl\.(\w+)", src)) == fields("lookup.rs", "Lookup")
    assert set(re.findall(r"\bt\.(\w+)", src)) - {"iter"} == fields("cross_table_lookup.rs", "TableWithColumns")
    assert set(re.findall(r"\bctl\.(\w+)", src)) == fields("cross_table_lookup.rs", "CrossTableLookup")
    ctl_rs, cc_rs, all_rs = (P / "cross_table_lookup.rs").read_text(), (P / "constraint_consumer.rs").read_text(), (P / "all_stark.rs").read_text()
    assert re.search(r"pub fn eval_with_next<FE, P, const D: usize>\(&self, v: &\[P\], next_v: &\[P\]\) -> P", ctl_rs)
    assert re.search(r"pub\(crate\) fn eval_filter<FE, P, const D: usize>\(&self, v: &\[P\], next_v: &\[P\]\) -> P", ctl_rs)
    new = re.search(r"pub fn new\(\s*alphas: Vec<P::Scalar>,\s*z_last: P,\s*lagrange_basis_first: P,\s*lagrange_basis_last: P,\s*\) -> Self", cc_rs)
    assert new and "pub fn accumulators(self) -> Vec<P>" in cc_rs
    assert len(re.search(r"ConstraintConsumer::<F>::new\((.*?)\);", src, flags=re.S).group(1).split("F::from_canonical_u64(")) - 1 == 5   # 2 alphas + 3
    assert "pub(crate) fn all_cross_table_lookups<F: Field>() -> Vec<CrossTableLookup<F>>" in all_rs and "pub(crate) fn all() -> [Self; NUM_TABLES]" in all_rs
    # the column counts it hard-codes are the reference's (and bench.py's)
    sys.path.insert(0, str(ROOT))
    import bench
    assert [int(x) for x in re.search(r"ncols: \[usize; 12\] = \[([^\]]*)\]", src).group(1).split(",")] == bench.NCOLS


def _ext_uses(text, crate):
    import re
    out = set()
    for path, group, single in re.findall(rf"use {crate}::((?:\w+::)*)(?:\{{([^}}]*)\}}|(\w+)(?:\s+as\s+\w+)?)\s*;", text):
        names = [n.split(" as ")[0].strip() for n in group.split(",")] if group else [single]
        out |= {(path, n) for n in names if n}
    return out


def test_plonky2_paths_of_the_shim_are_paths_the_reference_uses():
    """plonky2 is not in the tree, but the reference imports from it everywhere: each plonky2 item the shim imports must be imported
    from the SAME module path somewhere in the reference (a wrong path would be the first compile error under cargo)."""
    import re
    ref_uses = set()
    for f in (REF / "prover/src").rglob("*.rs"):
        ref_uses |= _ext_uses(re.sub(r"//[^\n]*", "", f.read_text()), "plonky2")
    assert len(ref_uses) > 60
    missing = []
    for src in ("b200.rs", "b200_ops.rs"):
        for use in sorted(_ext_uses(re.sub(r"//[^\n]*", "", (ROOT / "shim/src" / src).read_text()), "plonky2")):
            if use not in ref_uses:
                missing.append((src, "plonky2::" + use[0] + use[1]))
    # items plonky2 0.1.4 exports that the reference happens not to import by that path (checked by hand against the fork's layout)
    known = {"plonky2::field::extension::quadratic::QuadraticExtension", "plonky2::hash::poseidon::PoseidonPermutation",
             "plonky2::fri::proof::FriInitialTreeProof", "plonky2::fri::proof::FriQueryRound", "plonky2::fri::proof::FriQueryStep",
             "plonky2::hash::merkle_proofs::MerkleProof", "plonky2::hash::hashing::PlonkyPermutation",
             "plonky2::hash::hash_types::HashOut", "plonky2::hash::poseidon::PoseidonHash"}
    unexpected = [m for m in missing if m[1] not in known]
    assert not unexpected, unexpected
    # the other crates the shim names are dependencies of the prover crate
    cargo = (REF / "prover/Cargo.toml").read_text()
    for src in ("b200.rs", "b200_ops.rs"):
        text = re.sub(r"//[^\n]*", "", (ROOT / "shim/src" / src).read_text())
        for crate in set(re.findall(r"^use (\w+)::", text, flags=re.M)) - {"std", "core", "crate"}:
            assert re.search(rf"^{crate.replace('_', '[-_]')}\s*=", cargo, flags=re.M), f"{src}: crate {crate} is not a dependency of zkm-prover"
