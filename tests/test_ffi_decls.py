"""The Rust side of the drop-in cannot be compiled here (no cargo), so its FFI declarations are checked textually: every
`fn zkm_b200_*` inside an `extern "C"` block of shim/src/*.rs and of INTEGRATION.md's Rust snippets must name a function of
include/zkm_b200.h with the same parameters (count, order, pointer depth, constness, integer width) and return type, and the
`#[repr(C)]` structs of the shim must list the header structs' fields in order.  A drifted prototype would otherwise only show
as memory corruption on the maintainer's machine."""
import pathlib
import re

ROOT = pathlib.Path(__file__).resolve().parent.parent

RUST_SCALARS = {"c_int": "int", "u8": "uint8_t", "u32": "uint32_t", "u64": "uint64_t", "usize": "size_t", "c_char": "char", "c_void": "void",
                "f64": "double", "ZkmTable": "zkm_table_t", "ZkmStarkConfig": "zkm_stark_config_t", "ZkmTableRows": "zkm_table_rows_t",
                "ZkmOpLog": "zkm_op_log_t", "ZkmSplitState": "zkm_split_state_t"}
OPAQUE = {"zkm_pagetree_t", "zkm_splitter_t", "zkm_worker_t", "zkm_batch_t"}      # bound as *mut c_void on the Rust side


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def c_type(decl):
    """'const uint32_t roots_before[8]' -> ('uint32_t', ['const*']); pointer levels listed from the outermost inwards."""
    decl = decl.strip()
    array = decl.endswith("]")
    decl = re.sub(r"\[[^\]]*\]$", "", decl).strip()
    m = re.match(r"^(const\s+)?(\w+)\s*((?:\*\s*(?:const\s*)?)*)\s*(\w+)?$", decl)
    assert m, decl
    const_base, base, stars = bool(m.group(1)), m.group(2), m.group(3)
    levels = re.findall(r"\*\s*(const)?", stars)                 # innermost first; 'const' here qualifies the pointer itself
    if array:
        levels.append("")
    # constness of what each pointer level points to: level 0 points to the base type
    pointee_const = [const_base] + [bool(c) for c in levels[:-1]]
    if base in OPAQUE:
        base, pointee_const = "void", [False] + pointee_const[1:]
    return base, ["const*" if c else "mut*" for c in pointee_const[:len(levels)]]


def rust_type(t):
    t = t.strip()
    levels = []
    while True:
        m = re.match(r"^\*(const|mut)\s+(.*)$", t)
        if not m:
            break
        levels.append("const*" if m.group(1) == "const" else "mut*")
        t = m.group(2).strip()
    assert t in RUST_SCALARS, t
    base = RUST_SCALARS[t]
    levels = levels[::-1]                                        # innermost first, as in c_type
    if base == "void":
        levels[0] = "mut*"
    return base, levels


def header_functions():
    text = _strip_comments((ROOT / "include/zkm_b200.h").read_text())
    out = {}
    for ret, name, params in re.findall(r"^\s*([\w\s\*]+?)\s*\b(zkm_b200_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ps = [] if params.strip() in ("", "void") else [c_type(p) for p in params.split(",")]
        out[name] = (c_type(ret + " _")[0:2] if "*" in ret else (ret.strip(), []), ps)
    return out


def rust_functions(text):
    text = _strip_comments(text)
    out = []
    for block in re.findall(r'extern\s+"C"\s*\{(.*?)\n\s*\}|extern\s+"C"\s*\{(.*?)\}', text, flags=re.S):
        body = block[0] or block[1]
        for name, params, ret in re.findall(r"fn\s+(zkm_b200_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", body, flags=re.S):
            ps = [rust_type(p.split(":", 1)[1]) for p in params.split(",") if p.strip()]
            out.append((name, (rust_type(ret) if ret else ("void", [])), ps))
    return out


def _sources():
    for p in sorted((ROOT / "shim/src").glob("*.rs")):
        yield p.name, p.read_text()
    md = (ROOT / "INTEGRATION.md").read_text()
    yield "INTEGRATION.md", "\n".join(re.findall(r"```rust(.*?)```", md, flags=re.S))


def test_rust_extern_declarations_match_the_header():
    header = header_functions()
    assert len(header) >= 50 and "zkm_b200_prove_with_traces" in header
    seen = set()
    for src, text in _sources():
        for name, ret, params in rust_functions(text):
            assert name in header, f"{src}: {name} is not in include/zkm_b200.h"
            h_ret, h_params = header[name]
            assert tuple(ret) == tuple(h_ret), f"{src}: {name} returns {ret}, the header says {h_ret}"
            assert len(params) == len(h_params), f"{src}: {name} takes {len(params)} arguments, the header {len(h_params)}"
            for k, (r, h) in enumerate(zip(params, h_params)):
                assert r == h, f"{src}: {name} argument {k}: {r} != {h}"
            seen.add(name)
    assert {"zkm_b200_init", "zkm_b200_prove_with_traces", "zkm_b200_prove_with_ops", "zkm_b200_layout_check", "zkm_b200_last_timing",
            "zkm_b200_splitter_split", "zkm_b200_pagetree_split", "zkm_b200_worker_bind"} <= seen, seen


def test_rust_repr_c_structs_match_the_header():
    header = _strip_comments((ROOT / "include/zkm_b200.h").read_text())
    c_structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", header, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *more = [d.strip() for d in decl.split(",")]
            base, levels = c_type(re.sub(r"\[[^\]]*\]$", "", first))
            fname = re.search(r"(\w+)\s*(\[[^\]]*\])?$", first).group(1)
            arr = re.search(r"\[(\d+)\]$", first)
            fields.append((fname, base, levels, int(arr.group(1)) if arr else 0))
            for m in more:                                   # `uint32_t pc, segment_id;` / `uint8_t a[32], b[32];`
                arr = re.search(r"\[(\d+)\]$", m)
                fields.append((re.match(r"\**\s*(\w+)", m).group(1), base, levels, int(arr.group(1)) if arr else 0))
        c_structs[name] = fields
    assert {"zkm_table_t", "zkm_stark_config_t", "zkm_table_rows_t", "zkm_op_log_t", "zkm_split_state_t"} <= set(c_structs)
    checked = set()
    for src, text in _sources():
        for name, body in re.findall(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub\s+struct\s+(\w+)\s*\{(.*?)\}", _strip_comments(text), flags=re.S):
            want = c_structs[RUST_SCALARS[name]]
            got = []
            for f in body.split(","):
                if not f.strip():
                    continue
                fname, ftype = f.replace("pub ", "").split(":", 1)
                arr = re.match(r"^\s*\[\s*(\w+)\s*;\s*(\d+)\s*\]\s*$", ftype)
                base, levels = rust_type(arr.group(1) if arr else ftype)
                got.append((fname.strip(), base, levels, int(arr.group(2)) if arr else 0))
            assert got == want, f"{src}: {name}\n rust {got}\n C    {want}"
            checked.add(name)
    assert {"ZkmTable", "ZkmStarkConfig", "ZkmTableRows", "ZkmOpLog", "ZkmSplitState"} <= checked, checked


def test_rust_layout_keys_match_the_header_and_are_all_reported():
    """The layout handshake is keyed by numbers: the `lk::*` constants of shim/src/b200.rs must be the ZKM_LK_* values of the header
    (and of the reference-derived fixture), and layout_pairs() must report every CPU key exactly once -- a key left out would
    simply not be checked."""
    import json
    enum = re.search(r"enum\s*\w*\s*\{([^}]*ZKM_LK_NUM_COLUMNS[^}]*)\}", _strip_comments((ROOT / "include/zkm_b200.h").read_text())).group(1)
    header, nxt = {}, 0
    for item in enum.split(","):
        m = re.match(r"\s*ZKM_LK_(\w+)\s*(?:=\s*(\d+))?\s*$", item)
        if not m:
            assert not item.strip(), item
            continue
        nxt = int(m.group(2)) if m.group(2) else nxt
        header[m.group(1)] = nxt
        nxt += 1
    shim = _strip_comments((ROOT / "shim/src/b200.rs").read_text())
    lk = dict((k, int(v)) for k, v in re.findall(r"pub const (\w+): u32 = (\d+);", re.search(r"mod lk \{(.*?)\n\}", shim, flags=re.S).group(1)))
    assert lk == header and len(lk) >= 50
    fixture = json.loads((ROOT / "tests/golden/column_layout_v1.json").read_text())["keys"]
    assert {k[len("ZKM_LK_"):]: v for k, v in fixture.items()} == {k: v for k, v in header.items() if k != "NUM_COLUMNS"}
    body = re.search(r"pub fn layout_pairs.*?\n\}", shim, flags=re.S).group(0)
    used = re.findall(r"\(lk::(\w+)", body)
    assert sorted(used) == sorted(set(used)), "a key is reported twice"
    assert set(used) | {"NUM_COLUMNS"} == set(lk) | {"NUM_COLUMNS"} and "NUM_COLUMNS" in re.findall(r"lk::(\w+)", body)


def test_three_decoders_read_the_buffer_in_the_same_order():
    """The flat proof buffer has three readers: the Python walk of tests/test_wire_format.py (run against real proofs), the C++
    decoder of include/zkm_b200.hpp (compiled, round-trip tested) and the Rust decoder of shim/src/b200.rs (never compiled here).
    Their per-table read sequences -- written in source order = execution order in all three -- must be the same list of
    (12 fixed words | vec of unit k | count), and each must assign the reads to the same field names."""
    rust = _strip_comments((ROOT / "shim/src/b200.rs").read_text())
    rust = re.search(r"fn decode_stark_proof.*?\n\}", rust, flags=re.S).group(0)
    r_seq = [("w12" if a.startswith("words") else "n" if a in ("u()", "count()") else "v" + a[4]) for a in
             re.findall(r"r\.(words\(12\)|vec\(\d\)|count\(\)|u\(\))", rust)]
    py = (ROOT / "tests/test_wire_format.py").read_text()
    py = py[py.index("for _ in range(nt):"):py.index("out.append(d)")]
    p_seq = [("w12" if a.startswith("words") else "n" if a == "u()" else "v" + a[4]) for a in re.findall(r"r\.(words\(12\)|vec\(\d\)|u\(\))", py)]
    cpp = _strip_comments((ROOT / "include/zkm_b200.hpp").read_text())
    cpp = re.search(r"inline StarkProofWithMetadata decode_table\(Reader& r\) \{.*?\n\}", cpp, flags=re.S).group(0)
    unit = {"hashes()": "v4", "exts()": "v2", "fs()": "v1", "u()": "n"}
    c_seq = [("w12" if a.startswith("words") else unit[a]) for a in re.findall(r"r\.(words\(12\)|hashes\(\)|exts\(\)|fs\(\)|u\(\))", cpp)]
    want = ["w12", "v4", "v4", "v4", "v2", "v2", "v2", "v2", "v1", "v2", "n", "v4", "n", "n", "v1", "v4", "n", "v2", "v4", "v2", "n"]
    assert r_seq == want and p_seq == want and c_seq == want, (r_seq, p_seq, c_seq)
    # field order inside the openings (all six are vectors of the same shape but for ctl_zs_first: a swap would not change the sequence)
    order = ["local_values", "next_values", "auxiliary_polys", "auxiliary_polys_next", "ctl_zs_first", "quotient_polys"]
    assert re.findall(r"(\w+): (?:exts|fs)\(r\.vec", rust) == order
    assert re.findall(r'"(\w+)": r\.vec', py[py.index('d["openings"]'):py.index("ncaps = r.u()")]) == order
    assert re.findall(r"p\.openings\.(\w+) = r\.", cpp) == order
    caps = ["trace_cap", "auxiliary_polys_cap", "quotient_polys_cap"]
    assert re.findall(r"let (\w+) = cap\(r\.vec\(4\)", rust) == caps and re.findall(r"p\.(\w+) = r\.hashes\(\)", cpp) == caps
    assert re.findall(r'"(\w+)": H\(r\.vec\(4\)\)', py)[:3] == caps
