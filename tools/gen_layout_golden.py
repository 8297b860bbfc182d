#!/usr/bin/env python
"""Derives the CPU-table column layout from the REFERENCE sources (struct declarations in prover/src/cpu/columns/{mod,ops,
general}.rs, field order = declaration order) and the 12 tables' column counts, and writes them as the (key, value) pairs of
include/zkm_b200.h's layout handshake to tests/golden/column_layout_v1.json.  The test suite feeds that file to
zkm_b200_layout_check, which pins the constants the kernels were compiled with (zkm_b200/csrc/tables/*.h) to an independent
reading of the reference.  Run in the container that has /root/reference; the fixture travels, the reference does not.

    python tools/gen_layout_golden.py [/root/reference]
"""
import json
import pathlib
import re
import sys

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
ROOT = pathlib.Path(__file__).resolve().parent.parent
COLS = REF / "prover/src/cpu/columns"


def parse_structs(text):
    """name -> [(field, type)] for every `struct Name<T: Copy> { ... }` / `union Name<T: Copy> { ... }`."""
    out = {}
    for m in re.finditer(r"(?:struct|union)\s+(\w+)<T: Copy>\s*\{(.*?)\n\}", text, re.S):
        fields = []
        for line in m.group(2).splitlines():
            line = line.split("//")[0].strip()
            fm = re.match(r"(?:pub(?:\(crate\))?\s+)?(\w+)\s*:\s*(.+?),?$", line)
            if fm:
                fields.append((fm.group(1), fm.group(2).rstrip(",").strip()))
        out[m.group(1)] = fields
    return out


src = "".join((COLS / f).read_text() for f in ("mod.rs", "ops.rs", "general.rs"))
S = parse_structs(src)
NUM_GP_CHANNELS = int(re.search(r"NUM_GP_CHANNELS: usize = (\d+)", (REF / "prover/src/cpu/membus.rs").read_text()).group(1))
UNIONS = {"CpuGeneralColumnsView"}


def size_of(ty):
    ty = ty.strip()
    if ty == "T":
        return 1
    m = re.match(r"\[(.+);\s*(\w+)\]$", ty)
    if m:
        n = NUM_GP_CHANNELS if m.group(2) == "NUM_GP_CHANNELS" else int(m.group(2))
        return n * size_of(m.group(1))
    m = re.match(r"(\w+)<T>$", ty)
    name = m.group(1)
    sizes = [size_of(t) for _, t in S[name]]
    return max(sizes) if name in UNIONS else sum(sizes)


def offsets(name):
    """field -> offset inside struct `name` (all fields of a union start at 0)."""
    off, out = 0, {}
    for f, t in S[name]:
        out[f] = 0 if name in UNIONS else off
        off += size_of(t)
    return out


cpu = offsets("CpuColumnsView")
ops = offsets("OpsColumnsView")
br = offsets("CpuBranchView")
memio = offsets("MemIOView")
chan = offsets("MemoryChannelView")
sysc, misc, io = offsets("CpuSyscallView"), offsets("CpuMiscView"), offsets("CpuIOAuxView")

# keys of include/zkm_b200.h zkm_layout_key_t, parsed from the header so that the numbering has one source
hdr = (ROOT / "include/zkm_b200.h").read_text()
enum = re.search(r"typedef enum \{(.*?)\} zkm_layout_key_t;", hdr, re.S).group(1)
enum = re.sub(r"/\*.*?\*/", "", enum, flags=re.S)
KEY, nxt = {}, 0
for item in enum.split(","):
    item = item.strip()
    if not item:
        continue
    if "=" in item:
        name, val = [x.strip() for x in item.split("=")]
        nxt = int(val)
    else:
        name = item
    KEY[name] = nxt
    nxt += 1


def table_ncols():
    """Column counts of the 12 tables in Table enum order (all_stark.rs:97-110), each from its own column module."""
    def const(path, name):
        m = re.search(rf"{name}: usize = ([^;]+);", (REF / "prover/src" / path).read_text())
        return m.group(1).strip()
    n_limbs = 2                                               # arithmetic/columns.rs: N_LIMBS = 32 / LIMB_BITS(16)
    arith = 26 + 9 * n_limbs + 10                             # IS_* flags .. + shared + range-check columns (columns.rs:51-127)
    general = size_of("CpuGeneralColumnsView<T>")
    assert general == 102
    cpu_cols = size_of("CpuColumnsView<T>")
    # the other ten: SURVEY Appendix B (derived there from each table's column module; constants, not struct layouts)
    return [arith, cpu_cols, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]


pairs = []
for t, n in enumerate(table_ncols()):
    pairs.append((KEY["ZKM_LK_NUM_COLUMNS"] + t, n))
G = cpu["general"]
MC = cpu["mem_channels"]
abs_keys = {
    "ZKM_LK_CPU_IS_BOOTSTRAP_KERNEL": cpu["is_bootstrap_kernel"], "ZKM_LK_CPU_IS_EXIT_KERNEL": cpu["is_exit_kernel"],
    "ZKM_LK_CPU_CONTEXT": cpu["context"], "ZKM_LK_CPU_CODE_CONTEXT": cpu["code_context"],
    "ZKM_LK_CPU_PROGRAM_COUNTER": cpu["program_counter"], "ZKM_LK_CPU_NEXT_PROGRAM_COUNTER": cpu["next_program_counter"],
    "ZKM_LK_CPU_IS_KERNEL_MODE": cpu["is_kernel_mode"],
    "ZKM_LK_CPU_OP_BINARY_OP": cpu["op"] + ops["binary_op"], "ZKM_LK_CPU_OP_SYSCALL": cpu["op"] + ops["syscall"],
    "ZKM_LK_CPU_BRANCH_SHOULD_JUMP": cpu["branch"] + br["should_jump"], "ZKM_LK_CPU_BRANCH_IS_NE": cpu["branch"] + br["is_ne"],
    "ZKM_LK_CPU_OPCODE_BITS": cpu["opcode_bits"], "ZKM_LK_CPU_RS_BITS": cpu["rs_bits"], "ZKM_LK_CPU_RT_BITS": cpu["rt_bits"],
    "ZKM_LK_CPU_RD_BITS": cpu["rd_bits"], "ZKM_LK_CPU_SHAMT_BITS": cpu["shamt_bits"], "ZKM_LK_CPU_FUNC_BITS": cpu["func_bits"],
    "ZKM_LK_CPU_IS_POSEIDON_SPONGE": cpu["is_poseidon_sponge"], "ZKM_LK_CPU_IS_KECCAK_SPONGE": cpu["is_keccak_sponge"],
    "ZKM_LK_CPU_IS_SHA_EXTEND_SPONGE": cpu["is_sha_extend_sponge"], "ZKM_LK_CPU_IS_SHA_COMPRESS_SPONGE": cpu["is_sha_compress_sponge"],
    "ZKM_LK_CPU_GENERAL": G, "ZKM_LK_CPU_MEMIO_IS_LH": cpu["memio"] + memio["is_lh"],
    "ZKM_LK_CPU_MEMIO_AUX_FILTER": cpu["memio"] + memio["aux_filter"], "ZKM_LK_CPU_CLOCK": cpu["clock"], "ZKM_LK_CPU_MEM_CHANNELS": MC,
    "ZKM_LK_CPU_MEM_CHANNEL_STRIDE": size_of("MemoryChannelView<T>"),
    "ZKM_LK_CPU_CH_USED_REL": chan["used"], "ZKM_LK_CPU_CH_IS_READ_REL": chan["is_read"],
    "ZKM_LK_CPU_CH_ADDR_CONTEXT_REL": chan["addr_context"], "ZKM_LK_CPU_CH_ADDR_SEGMENT_REL": chan["addr_segment"],
    "ZKM_LK_CPU_CH_ADDR_VIRTUAL_REL": chan["addr_virtual"], "ZKM_LK_CPU_CH_VALUE_REL": chan["value"],
    "ZKM_LK_CPU_G_SYSCALL_COND_REL": sysc["cond"], "ZKM_LK_CPU_G_SYSCALL_SYSNUM_REL": sysc["sysnum"],
    "ZKM_LK_CPU_G_SYSCALL_A0_REL": sysc["a0"], "ZKM_LK_CPU_G_SYSCALL_A1_REL": sysc["a1"],
    "ZKM_LK_CPU_G_MISC_RS_BITS_REL": misc["rs_bits"], "ZKM_LK_CPU_G_MISC_IS_MSB_REL": misc["is_msb"],
    "ZKM_LK_CPU_G_MISC_IS_LSB_REL": misc["is_lsb"], "ZKM_LK_CPU_G_MISC_AUXM_REL": misc["auxm"], "ZKM_LK_CPU_G_MISC_AUXL_REL": misc["auxl"],
    "ZKM_LK_CPU_G_MISC_AUXS_REL": misc["auxs"], "ZKM_LK_CPU_G_MISC_RD_INDEX_REL": misc["rd_index"],
    "ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_0_REL": misc["rd_index_eq_0"], "ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_29_REL": misc["rd_index_eq_29"],
    "ZKM_LK_CPU_G_IO_RS_LE_REL": io["rs_le"], "ZKM_LK_CPU_G_IO_RT_LE_REL": io["rt_le"], "ZKM_LK_CPU_G_IO_MEM_LE_REL": io["mem_le"],
    "ZKM_LK_CPU_G_IO_AUX_RS0_MUL_RS1_REL": io["aux_rs0_mul_rs1"],
    "ZKM_LK_CPU_G_LOGIC_DIFF_PINV_REL": offsets("CpuLogicView")["diff_pinv"], "ZKM_LK_CPU_G_HASH_VALUE_REL": offsets("CpuHashView")["value"],
    "ZKM_LK_CPU_G_KHASH_VALUE_REL": offsets("CpuKHashView")["value"], "ZKM_LK_CPU_G_SHASH_VALUE_REL": offsets("CpuSHashView")["value"],
    "ZKM_LK_CPU_G_ELEMENT_VALUE_REL": offsets("CpuElementView")["value"],
}
for k, v in abs_keys.items():
    pairs.append((KEY[k], v))
out = {"generated_by": "tools/gen_layout_golden.py from the reference's struct declarations (declaration order)",
       "reference_files": ["prover/src/cpu/columns/mod.rs", "prover/src/cpu/columns/ops.rs", "prover/src/cpu/columns/general.rs",
                           "prover/src/cpu/membus.rs"],
       "keys": {k: KEY[k] for k in abs_keys}, "pairs": pairs}
dst = ROOT / "tests/golden/column_layout_v1.json"
dst.write_text(json.dumps(out, indent=1))
print(f"{len(pairs)} pairs -> {dst}")
