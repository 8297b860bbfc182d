"""Builds zkm_b200/libzkm_b200.so (hand-written CUDA for sm_100a + C++ host layer) with nvcc.

Run `python -m zkm_b200.build`.  The .so is built in-tree so it travels with the repo snapshot
to the GPU box; there is no JIT and no CPU fallback.
"""
import hashlib
import os
import pathlib
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
# A/B builds for tuning runs: ZKM_BUILD_TAG=x ZKM_EXTRA_NVCC="-DFOO=1" python -m zkm_b200.build  ->  libzkm_b200_x.so, which
# zkm_b200.lib loads instead of the product library when ZKM_B200_LIB_TAG=x is set.  The product build uses neither.
TAG = os.environ.get("ZKM_BUILD_TAG", "")
OBJ = ROOT / ("build_" + TAG if TAG else "build")
LIB = ROOT / (f"libzkm_b200_{TAG}.so" if TAG else "libzkm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("ZKM_EXTRA_NVCC", "").split()


def _deps_hash(src: pathlib.Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    # quotient_p1..3.cu #include "quotient.cu": sources that include other sources depend on every .cu as well
    extra = list(CSRC.glob("*.cu")) if b'.cu"' in src.read_bytes() else []
    for p in sorted(list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.h")) + [ROOT.parent / "include/zkm_b200.h"] + extra):
        h.update(p.read_bytes())
    return h.hexdigest()


def _compile(src: pathlib.Path) -> pathlib.Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".hash")
    want = _deps_hash(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src}")
    stamp.write_text(want)
    return obj


def build(verbose: bool = True) -> pathlib.Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = ["nvcc", "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.run(cmd, check=True)
    if verbose:
        print(f"built {LIB} from {len(srcs)} sources")
    return LIB


if __name__ == "__main__":
    build()
