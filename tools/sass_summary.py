#!/usr/bin/env python
"""SASS evidence for the hot kernels (needs no GPU: cuobjdump on the built objects).  For every kernel matching one of the
patterns: instruction count, opcode histogram by issue pipe (as measured in profiles/r1b_pipe_bench.txt: IMAD* on the FMA pipe,
IADD3/LOP3/SHF/SEL/ISETP/... on the ALU pipe, D* on FP64, I2F/F2I on XU) and the memory instructions; the complete listing of the
two smallest hot kernels is written next to it.     python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
OBJ = ROOT / "zkm_b200" / "build"
PATTERNS = [("merkle.o", r"lde_leaf_hash_kernel|merkle_level_kernel"), ("ntt.o", r"ntt_pass_kernelILi1[01]E"), ("quotient.o", r"quotient_kernelILi1ELb0"),
            ("openings.o", r"open_segments_kernel"), ("fri.o", r"fri_reduce_kernel")]
PIPE = [("FMA", r"^(IMAD|FFMA|FMUL|FADD)"), ("FP64", r"^D(ADD|MUL|FMA|SETP)"), ("XU", r"^(I2F|F2I|MUFU|F2F)"),
        ("LSU", r"^(LD|ST|ATOM|RED|LDG|STG|LDS|STS|LDL|STL|LDC)"), ("TMA", r"^(UTMA|UBLKCP|LDGSTS)"), ("CTRL", r"^(BRA|EXIT|BAR|CALL|RET|NOP|WARPSYNC|BSSY|BSYNC|DEPBAR)"),
        ("UNIFORM", r"^U[A-Z]")]


def pipe_of(op):
    for name, rx in PIPE:
        if re.match(rx, op):
            return name
    return "ALU"


def main():
    full_dir = ROOT / "profiles"
    for obj, rx in PATTERNS:
        names = subprocess.run(["cuobjdump", "-elf", str(OBJ / obj)], capture_output=True, text=True).stdout
        funcs = sorted(set(re.findall(r"\.text\.(_Z\w+)", names)))
        for fn in funcs:
            if not re.search(rx, fn):
                continue
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, str(OBJ / obj)], capture_output=True, text=True).stdout
            ops = []
            for line in sass.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
                if m:
                    ops.append(m.group(1))
            by_pipe = collections.Counter(pipe_of(o.split(".")[0]) for o in ops)
            top = collections.Counter(o.split(".")[0] for o in ops).most_common(12)
            demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
            print(f"{demangled}\n  {len(ops)} instructions (static)  pipes: " + ", ".join(f"{k} {v}" for k, v in by_pipe.most_common()))
            print("  top opcodes: " + ", ".join(f"{k} {v}" for k, v in top))
            print(f"  TMA / bulk-copy instructions: {by_pipe.get('TMA', 0)} (tiles are staged by LDG + STS: the coset scale rides on the load in registers)"
                  if "ntt_pass" in fn else "")
            if re.search(r"lde_leaf_hash_kernel|ntt_pass_kernelILi10E", fn):
                short = "leaf_hash" if "leaf" in fn else "ntt_pass_10"
                (full_dir / f"r2_sass_{short}.txt").write_text(sass)


if __name__ == "__main__":
    main()
