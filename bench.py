#!/usr/bin/env python
"""bench.py — measures the zkm STARK proving hot path on B200 (contract: task prompt ④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload U20|U16|...]

A "step" = one pass of the hot path over one synthetic segment (workload U20: the
12 MIPS STARK tables with Arithmetic/Cpu/Memory at 2^20 rows, Logic at 2^18, the rest at 2^6).
  value  : whole-job steps/s with the trace columns already resident in HBM (device timing, CUDA
           events on the library's stream, max over ranks)
  e2e    : the same through the C-ABI entry point with HOST (pinned) column buffers: H2D of every
           trace column and D2H of the caps/proof inside the timed region
  roofline / cpu_baseline : see DESIGN.md §Measurement
One process per GPU; ranks prove independent segments (no data-path collective): weak scaling.
PyTorch is used only for device/pinned memory, torch.distributed barriers and the max-over-ranks.
"""
import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# Table enum order (reference prover/src/all_stark.rs:97-110) and column counts.
TABLES = ["Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge", "ShaExtend",
          "ShaExtendSponge", "ShaCompress", "ShaCompressSponge", "Logic", "Memory"]
NCOLS = [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]


def workload_log_heights(name):
    if name.startswith("U"):
        big = int(name[1:])
        h = [6] * 12
        h[0] = h[1] = h[11] = big
        h[10] = max(6, big - 2)
        h[0] = max(h[0], 16)           # Arithmetic >= 2^16 rows (arithmetic_stark.rs:123,178-181)
        return h
    raise SystemExit(f"unknown workload {name}")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


class Segment:
    """One synthetic segment (all 12 AllStark tables) and the hot-path step over it:
    a full `prove_with_traces` (trace commitments, CTL/logUp columns, quotients, openings, FRI)."""
    SYSTEM_ALL_STARK = 0

    def __init__(self, lib, workload, seed_offset=0, rank=0, world=1):
        import torch
        from zkm_b200 import lib as zl
        self.zl, self.lib, self.torch = zl, lib, torch
        self.rank, self.world = rank, world
        self.heights = workload_log_heights(workload)
        self.stages = "full prove_with_traces: 12 tables x (trace commit, CTL/logUp aux, quotient, openings, FRI incl. PoW + 37 queries)"
        self.metric = f"MIPS-segment proofs/sec ({workload}: 2^{max(self.heights)}-row synthetic segment, full STARK prove)"
        self.unit = "proofs/s"
        self.cfg = zl.standard_fast_config(lib)
        self.dev = []
        err = C.c_void_p()
        for t, (nc, lg) in enumerate(zip(NCOLS, self.heights)):
            buf = torch.empty(nc << lg, dtype=torch.int64, device="cuda")
            seed = (0x5EED000000000000 | (t << 16)) + (seed_offset << 32)
            zl.check(lib, lib.zkm_b200_synth_trace_device(self.SYSTEM_ALL_STARK, t, lg, seed, buf.data_ptr(), C.byref(err)), err)
            self.dev.append(buf)
        self.input_bytes = sum(8 * (nc << lg) for nc, lg in zip(NCOLS, self.heights))
        self.output_bytes = 0
        self.host = None
        self.shapes = (zl.Table * 12)(*[zl.Table(None, nc, lg) for nc, lg in zip(NCOLS, self.heights)])
        self.dptrs = (C.c_void_p * 12)(*[b.data_ptr() for b in self.dev])
        self.rb = (C.c_uint32 * 8)(*range(1, 9))
        self.ra = (C.c_uint32 * 8)(*range(11, 19))
        self.userdata = bytes(32)
        self.last_proof_words = 0
        self.kept = []

    def sync(self):
        err = C.c_void_p()
        self.zl.check(self.lib, self.lib.zkm_b200_sync(C.byref(err)), err)

    def timed(self, fn):
        err = C.c_void_p()
        ms = C.c_double()
        self.zl.check(self.lib, self.lib.zkm_b200_timer_start(C.byref(err)), err)
        fn()
        self.zl.check(self.lib, self.lib.zkm_b200_timer_stop(C.byref(ms), C.byref(err)), err)
        return ms.value

    def _finish(self, rc, err, out, words, keep=False):
        self.zl.check(self.lib, rc, err)
        self.last_proof_words = words.value
        self.output_bytes = 8 * words.value
        proof = np.ctypeslib.as_array(out, shape=(words.value,)).copy() if keep else None
        self.lib.zkm_b200_free(out)
        return proof

    def step_device(self):
        lib = self.lib
        out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
        rc = lib.zkm_b200_prove_system_device(self.SYSTEM_ALL_STARK, self.shapes, self.dptrs, 12, self.rb, self.ra, self.userdata, 32,
                                              C.byref(self.cfg), C.byref(out), C.byref(words), C.byref(err))
        self._finish(rc, err, out, words)

    def step_device_keep(self):
        lib = self.lib
        out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
        rc = lib.zkm_b200_prove_system_device(self.SYSTEM_ALL_STARK, self.shapes, self.dptrs, 12, self.rb, self.ra, self.userdata, 32,
                                              C.byref(self.cfg), C.byref(out), C.byref(words), C.byref(err))
        return self._finish(rc, err, out, words, keep=True)

    def prepare_host(self, pinned=True):
        torch = self.torch
        self.host, made = [], []
        for t, (nc, lg) in enumerate(zip(NCOLS, self.heights)):
            hbuf = torch.empty(nc << lg, dtype=torch.int64, pin_memory=pinned)
            hbuf.copy_(self.dev[t])
            arr = hbuf.numpy().view(np.uint64).reshape(nc, 1 << lg)
            self.host.append(hbuf)
            made.append(self.zl.make_table(arr))
        torch.cuda.synchronize()
        self._keep = made
        self.tables = (self.zl.Table * 12)(*[m[0] for m in made])

    def step_e2e(self, gather=True):
        """The reference-facing call: zkm_b200_prove_with_traces with HOST column pointers (H2D of every trace
        column inside, D2H of the finished proof buffer)."""
        lib = self.lib
        out, words, err = C.POINTER(C.c_uint64)(), C.c_size_t(), C.c_void_p()
        rc = lib.zkm_b200_prove_with_traces(self.tables, self.rb, self.ra, self.userdata, 32, C.byref(self.cfg), C.byref(out),
                                            C.byref(words), C.byref(err))
        proof = self._finish(rc, err, out, words, keep=self.world > 1)
        if self.world > 1 and not gather:
            self.kept.append(proof)              # gathered by the caller once every worker has finished
        elif self.world > 1:
            # the path's only exchange: finished proofs gathered on rank 0 (NCCL)
            from zkm_b200 import multi
            multi.gather_proofs([proof], [self.rank], self.world)

    def profile_reset(self):
        err = C.c_void_p()
        self.zl.check(self.lib, self.lib.zkm_b200_profile_reset(C.byref(err)), err)

    def profile_families(self):
        lib = self.lib
        ptr = lib.zkm_b200_profile_families()
        names = C.cast(ptr, C.c_char_p).value.decode().split("\n") if ptr else []
        lib.zkm_b200_free_string(ptr)
        out = {}
        for n in filter(None, names):
            ms, la, by, err = C.c_double(), C.c_uint64(), C.c_double(), C.c_void_p()
            self.zl.check(lib, lib.zkm_b200_profile_get(n.encode(), C.byref(ms), C.byref(la), C.byref(by), C.byref(err)), err)
            out[n] = {"ms": ms.value, "launches": la.value, "bytes": by.value}
            tb = C.c_double()
            if lib.zkm_b200_profile_get_traffic(n.encode(), C.byref(tb), C.byref(err)) == 0:
                out[n]["traffic_bytes"] = tb.value
            else:
                lib.zkm_b200_free_string(err)
        return out

    @staticmethod
    def ncu_traffic(family):
        """dram bytes per launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json), or None."""
        f = ROOT / "profiles" / "ncu_traffic.json"
        if f.exists():
            return json.loads(f.read_text()).get(family)
        return None


def host_threads():
    """Host threads this process may use (the CPU arm runs on all of them)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(workload):
    """The `config` object of the JSON line: identical in the B200 arm and the reference arm (same workload, same stages)."""
    heights = workload_log_heights(workload)
    in_bytes = sum(8 * (nc << lg) for nc, lg in zip(NCOLS, heights))
    return {"workload": workload, "log_heights": heights, "tables": TABLES,
            "stages": "full prove_with_traces: 12 tables x (trace commit, CTL/logUp aux, quotient, openings, FRI incl. PoW + 37 queries)",
            "stark_config": "standard_fast_config: rate_bits 2, cap_height 4, pow_bits 16, 37 queries, 2 challenges, arity 16",
            "l2": (f"inputs {in_bytes / 1e9:.2f} GB per proof > 126 MB L2 (no flush needed)" if in_bytes > 126e6 else
                   f"inputs {in_bytes / 1e6:.1f} MB per proof FIT in the 126 MB L2 and nothing flushes it: a test-sized workload, not a bench line")}


def workload_metric(workload):
    return f"MIPS-segment proofs/sec ({workload}: 2^{max(workload_log_heights(workload))}-row synthetic segment, full STARK prove)"


CPU_NOTE = "restated C++ oracle, threaded like the reference's Rayon loops (the Rust reference cannot be built here: no cargo, plonky2 un-vendored)"


def _cpu_cache_path(workload, ncores):
    import socket
    return pathlib.Path("/tmp") / f"zkm_b200_cpu_arm_{socket.gethostname()}_{workload}_{ncores}.json"


def cpu_pass(workload, ncores, tiny=False, use_cache=False):
    """CPU arm: the restated oracle (oracle/liborc.so, `kind: port`) running the same full prove_with_traces on the FULL
    workload, once, on `ncores` host threads -- no sampling and no extrapolation.  `tiny` proves the same 12 tables at 2^8
    rows (untimed warm-up: pages the library and its tables in).  With `use_cache` a figure measured earlier on this host
    (same workload and thread count; written by every full run) is reused instead of spending minutes of host time again."""
    import hashlib
    from oracle import binding
    cache = _cpu_cache_path(workload, ncores)
    if use_cache and not tiny and cache.exists():
        try:
            c = json.loads(cache.read_text())
            c["sample"] += " [figure reused from this host's earlier run: " + str(cache) + "]"
            return c
        except Exception:
            pass
    orc = binding.load()
    orc.orc_set_threads(ncores)
    full = workload_log_heights(workload)
    heights = [min(h, 8) for h in full] if tiny else full
    traces = synthetic_traces_host(heights)
    t0 = time.perf_counter()
    proof = binding.prove_system(orc, 0, traces)
    dt = time.perf_counter() - t0
    out = {"value": 1.0 / dt, "unit": "proofs/s", "seconds": dt, "metric": workload_metric(workload),
           "sample": f"full: one complete proof at heights {heights} in {dt:.1f} s on {ncores} threads, nothing extrapolated",
           "proof_sha256": hashlib.sha256(proof.tobytes()).hexdigest(), "note": CPU_NOTE}
    if not tiny:
        try:
            cache.write_text(json.dumps(out))
        except Exception:
            pass
    return out


def synthetic_traces_host(log_heights, seed=0x5EED000000000000):
    """numpy twin of the device generator's contract: uniform cells, CTL-filter columns one-hot (at most one flag per
    row).  Used by the CPU arm only (it has no GPU); the flag columns are the union of the filter columns of
    every AllStark CTL, listed per table below (derived from zkm_b200/csrc/tables/all_stark.h)."""
    rng = np.random.default_rng(seed & 0xFFFFFFFF)
    P = np.uint64(0xFFFFFFFF00000001)
    flags = {
        0: list(range(0, 26)),                                     # Arithmetic op flags
        1: [7, 8, 10, 16, 17] + [82, 83, 84, 85] + [205 + 6 * c for c in range(9)],   # Cpu: binary/imm/logic/shift flags, sponge flags, channel.used
        2: [0],                                                    # Poseidon FILTER
        3: [0] + list(range(14, 46)),                              # PoseidonSponge is_full_input_block, is_final_input_len
        4: [0, 23],                                                # Keccak reg_step(0), reg_step(23)
        5: [0] + list(range(40, 176)),                             # KeccakSponge
        6: [77],                                                   # ShaExtend is_real_round
        7: list(range(0, 48)),                                     # ShaExtendSponge round flags
        8: list(range(159, 224)),                                  # ShaCompress round flags
        9: [126],                                                  # ShaCompressSponge is_real_round
        10: [0, 1, 2, 3],                                          # Logic op flags
        11: [0],                                                   # Memory FILTER
    }
    out = []
    for t, (nc, lg) in enumerate(zip(NCOLS, log_heights)):
        n = 1 << lg
        a = rng.integers(0, 2**63, size=(nc, n), dtype=np.uint64) % P
        f = flags[t]
        hot = rng.integers(-len(f), len(f), size=n)
        for k, c in enumerate(f):
            a[c] = (hot == k).astype(np.uint64)
        out.append(np.ascontiguousarray(a))
    return out


def run_n22(args):
    """BASELINE config #3: 2^22-row columns, NTT + Merkle only (PolynomialBatch::from_values), C in {1, 13, 54, 259}.
    Prints one JSON line per C with the NTT figure BASELINE asks for: 48*n*C algorithmic bytes (read values,
    write coefficients, write the 4n LDE) over the NTT kernels' device time, against the measured HBM peak."""
    import torch
    from zkm_b200 import lib as zl
    lib = zl.init(0)
    peak, peak_kind = peaks()
    log_n = int(args.workload[1:])
    n = 1 << log_n
    for ncols in (1, 13, 54, 259):
        buf = torch.empty(ncols * n, dtype=torch.int64, device="cuda")
        err = C.c_void_p()
        zl.check(lib, lib.zkm_b200_synth_columns_device(buf.data_ptr(), ncols, log_n, 0x5EED000000000000 | ncols, C.byref(err)), err)
        cap = np.zeros(64, dtype=np.uint64)

        def step():
            h = C.c_void_p()
            zl.check(lib, lib.zkm_b200_commit_values_device(buf.data_ptr(), ncols, log_n, 2, 4, C.byref(h), zl.u64ptr(cap), C.byref(err)), err)
            lib.zkm_b200_batch_free(h)

        for _ in range(max(3, args.warmup)):
            step()
        lib.zkm_b200_profile_enable(1)
        zl.check(lib, lib.zkm_b200_profile_reset(C.byref(err)), err)
        ms = C.c_double()
        zl.check(lib, lib.zkm_b200_timer_start(C.byref(err)), err)
        for _ in range(args.steps):
            step()
        zl.check(lib, lib.zkm_b200_timer_stop(C.byref(ms), C.byref(err)), err)
        fam = {}
        for name in ("ntt_pass", "leaf_hash", "merkle_levels"):
            m, la, by = C.c_double(), C.c_uint64(), C.c_double()
            rc = lib.zkm_b200_profile_get(name.encode(), C.byref(m), C.byref(la), C.byref(by), C.byref(err))
            if rc == 0:
                fam[name] = {"ms_per_step": m.value / args.steps, "launches_per_step": la.value / args.steps}
            else:
                lib.zkm_b200_free_string(err)
        lib.zkm_b200_profile_enable(0)
        t_ntt = fam.get("ntt_pass", {"ms_per_step": 0})["ms_per_step"]
        alg = 48.0 * n * ncols
        ach = alg / (t_ntt * 1e-3) / 1e9 if t_ntt else 0.0
        perms = 4 * n * ((ncols + 7) // 8 if ncols > 4 else 0) + 4 * n - 16
        t_hash = sum(fam.get(k, {"ms_per_step": 0})["ms_per_step"] for k in ("leaf_hash", "merkle_levels"))
        print(json.dumps({"metric": f"NTT GB/s (iNTT + coset LDE x4 of {ncols} columns x 2^{log_n}; 48*n*C algorithmic bytes)", "value": ach,
                          "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
                          "ms_per_step": ms.value / args.steps, "higher_is_better": True, "dtype": "u64 (Goldilocks)", "data": "synthetic",
                          "config": {"workload": args.workload, "columns": ncols, "stages": "a1+a2 (from_values: NTT + Poseidon Merkle cap)"},
                          "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                                       "frac": ach / peak, "peak_kind": peak_kind, "traffic": None,
                                       "pass_traffic_GBps": 160.0 * n * ncols / (t_ntt * 1e-3) / 1e9 if t_ntt else 0.0},
                          "poseidon": {"permutations": perms, "Gperm_per_s": perms / (t_hash * 1e-3) / 1e9 if t_hash else 0.0},
                          "kernel_families": fam}), flush=True)
        del buf
        torch.cuda.empty_cache()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust reference cannot be built in this image
    (no cargo; plonky2 un-vendored -- DESIGN.md), so this times the restated C++ oracle on all host cores, on the SAME
    workload as the B200 arm: one full proof (a U20 proof is minutes of host time, so exactly one is timed whatever --steps
    says; the line reports the steps actually run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = host_threads()
    if args.warmup:
        cpu_pass(args.workload, ncores, tiny=True)
    r = cpu_pass(args.workload, ncores)
    line = {"impl": "reference", "metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1000.0 * r["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks)", "data": "synthetic", "config": workload_config(args.workload),
            "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": ncores, "kind": "port", "sample": r["sample"],
                             "note": r["note"], "proof_sha256": r["proof_sha256"]},
            "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this rank's host threads, and first-touch its pinned trace buffers, on the CPU socket the GPU
    hangs off (sysfs local_cpulist of the GPU's PCI function).  Best effort: returns the CPU list used, or None."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                       # sysfs uses a 4-digit PCI domain
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus)"
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="U20")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workers", type=int, default=0, help="proofs in flight per GPU (worker contexts); 0 = 3")
    ap.add_argument("--no-pageable", action="store_true", help="skip the e2e leg from pageable host memory")
    ap.add_argument("--no-in-segment", action="store_true", help="N > 1: skip the in-segment sharding measurement")
    ap.add_argument("--host-memory", default="pinned", choices=["pinned", "pageable"],
                    help="e2e leg: where the caller's trace columns live (the reference's Vec<F> columns are pageable)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload.startswith("N"):
        return run_n22(args)

    import torch
    import torch.distributed as dist
    from zkm_b200 import lib as zl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = zl.init(local)
    # measured on one B200 (profiles/r1c_workers_sweep.txt): 1 worker 2.88, 2: 2.99, 3: 3.16, 4: 3.20 proofs/s
    # the same number of proofs in flight per GPU at every N (each worker is one mostly-blocked host thread + its uploader)
    NW = args.workers if args.workers > 0 else (3 if host_threads() >= 4 else 1)
    # one synthetic segment per worker context: NW proofs are in flight on this GPU at any time (one host thread, one pair of
    # streams and one arena each, include/zkm_b200.h "Worker contexts"); a step = NW segments, one per worker
    segs = [Segment(lib, args.workload, seed_offset=rank * NW + i, rank=rank, world=world) for i in range(NW)]
    seg = segs[0]
    workers = [zl.Worker(lib) for _ in range(NW)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        seg.sync()

    def timed_concurrent(step_name, warm, steps, after_join=None):
        """Every worker thread runs `warm` untimed and then `steps` timed calls of Segment.<step_name>; the timed region is
        bracketed by CUDA events on the (idle) main stream, recorded after a device-wide barrier and after every worker has
        drained its stream."""
        import threading
        ready, go = threading.Barrier(NW + 1), threading.Barrier(NW + 1)
        errors = []

        def body(i):
            try:
                with workers[i]:
                    for _ in range(warm):
                        getattr(segs[i], step_name)()
                    segs[i].sync()
                    ready.wait(timeout=600)
                    go.wait(timeout=600)
                    for _ in range(steps):
                        getattr(segs[i], step_name)()
                    segs[i].sync()
            except BaseException as e:           # never leave the other threads waiting on a barrier
                errors.append(e)
                ready.abort()
                go.abort()
        threads = [threading.Thread(target=body, args=(i,)) for i in range(NW)]
        for t in threads:
            t.start()
        try:
            ready.wait(timeout=600)
            barrier()
            err = C.c_void_p()
            zl.check(lib, lib.zkm_b200_timer_start(C.byref(err)), err)
            timed_concurrent.l0 = lib.zkm_b200_launch_count()
            go.wait(timeout=600)
        except threading.BrokenBarrierError:
            pass
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        if after_join:
            after_join()
        ms, err = C.c_double(), C.c_void_p()
        zl.check(lib, lib.zkm_b200_timer_stop(C.byref(ms), C.byref(err)), err)
        timed_concurrent.launches = lib.zkm_b200_launch_count() - timed_concurrent.l0
        return ms.value

    W = max(3, args.warmup)
    # ---- device-resident timing: K steps with no instrumentation between launches ----
    lib.zkm_b200_profile_enable(0)
    with ClockSampler(local) as clk:
        t_dev = timed_concurrent("step_device", W, args.steps)
        barrier()
    launches = timed_concurrent.launches
    # ---- K single-context steps with a CUDA-event pair around every launch (per-family device times, roofline); the events
    # serialise the launch-bound phases of the 8 small tables, so this pass is reported next to the clean one, not instead ----
    for _ in range(2):
        seg.step_device()
    lib.zkm_b200_profile_enable(1)
    seg.profile_reset()
    barrier()
    t_prof = seg.timed(lambda: [seg.step_device() for _ in range(args.steps)])
    barrier()
    fam = seg.profile_families()
    lib.zkm_b200_profile_enable(0)
    # ---- end to end through the C ABI with host buffers ----
    for sg in segs:
        sg.prepare_host(pinned=args.host_memory == "pinned")

    def gather_all():
        if world > 1:                            # the path's only exchange: finished proofs gathered on rank 0 (NCCL)
            from zkm_b200 import multi
            proofs = [p for sg in segs for p in sg.kept]
            multi.gather_proofs(proofs, [rank * len(proofs) + k for k in range(len(proofs))], len(proofs) * world)
            for sg in segs:
                sg.kept = []
    for sg in segs:
        sg.step_e2e_nogather = lambda sg=sg: sg.step_e2e(gather=False)
    t_e2e = timed_concurrent("step_e2e_nogather", 1, args.steps, after_join=gather_all)
    barrier()
    # the same call from PAGEABLE host columns -- what the reference's caller holds (one Vec<F> per column) -- at N = 1
    t_e2e_pageable, pageable_steps = None, 0
    if world == 1 and args.host_memory == "pinned" and not args.no_pageable:
        for sg in segs:
            sg.prepare_host(pinned=False)
        pageable_steps = min(args.steps, 3)
        t_e2e_pageable = timed_concurrent("step_e2e_nogather", 1, pageable_steps)
        barrier()
    for w in workers:
        w.close()
    # ---- latency of ONE proof on one GPU (one context, nothing else in flight) ----
    barrier()
    t_single = seg.timed(lambda: [seg.step_device() for _ in range(args.steps)]) / args.steps
    # ---- in-segment sharding (north_star): all N (<= 8) GPUs prove ONE segment together ----
    in_segment = None
    if world > 1 and not args.no_in_segment:
        import hashlib
        from zkm_b200 import multi
        gi, ri, g = multi.shard_group_init(lib)
        del segs[1:]
        torch.cuda.empty_cache()
        sseg = Segment(lib, args.workload, seed_offset=1000 + gi, rank=rank, world=world)      # same traces on every rank of a group
        for _ in range(2):
            sseg.step_device()
        barrier()
        t_shard = sseg.timed(lambda: [sseg.step_device() for _ in range(args.steps)]) / args.steps
        proof = sseg.step_device_keep()
        barrier()
        multi.shard_group_shutdown(lib)
        ref = sseg.step_device_keep() if ri == 0 else None       # the same segment proved by one GPU alone
        digest = hashlib.sha256(proof.tobytes()).digest()
        same_as_single = (hashlib.sha256(ref.tobytes()).digest() == digest) if ref is not None else True
        dg = torch.tensor(list(digest) + [int(same_as_single)], dtype=torch.uint8, device="cuda")
        alld = [torch.zeros_like(dg) for _ in range(world)]
        dist.all_gather(alld, dg)
        ts = torch.tensor([t_shard], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        groups = world // g
        ok = all(bool((alld[r][:32] == alld[(r // g) * g][:32]).all()) for r in range(world)) and all(bool(a[32]) for a in alld)
        in_segment = {"group_size": g, "groups": groups, "ms_per_proof": float(ts.item()), "proofs_per_s": groups / (float(ts.item()) * 1e-3),
                      "exchange": "NCCL: cap all-gather per commitment (512 B), quotient halves broadcast, FRI query answers all-gather",
                      "proofs_identical_across_ranks_and_to_single_gpu": ok}
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, t_single], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_single = t.tolist()
    if rank == 0:
        peak, peak_kind = peaks()
        total_ms = sum(v["ms"] for v in fam.values()) or 1.0
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        value = args.steps * NW * world / (t_dev * 1e-3)
        e2e = args.steps * NW * world / (t_e2e * 1e-3)

        def roof(name, v, note):
            """achieved = algorithmic bytes of the family's launches / their device time (CUDA events on
            the launching stream), against the measured copy peak; `traffic` = DRAM bytes per launch from the committed
            `ncu --set full` capture."""
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else 0.0
            r = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "peak_kind": peak_kind, "traffic": seg.ncu_traffic(name), "launches": v["launches"],
                 "avg_launch_ms": v["ms"] / max(1, v["launches"]), "share_of_step": v["ms"] / total_ms,
                 "algorithmic_bytes_per_launch": v["bytes"] / max(1, v["launches"]), "note": note}
            if name == "ntt_pass" and v.get("traffic_bytes"):      # the second counter is pass traffic for the NTT family only
                r["pass_traffic_GBps"] = v["traffic_bytes"] / (v["ms"] * 1e-3) / 1e9
                r["pass_traffic_frac"] = r["pass_traffic_GBps"] / peak
            return r
        notes = {"leaf_hash": "Poseidon leaf hashing: 8 B per column per leaf against ~18 000 instructions per permutation -- "
                              "issue-bound integer/FP64 kernel, its HBM fraction is tiny by construction (DESIGN.md section 3 "
                              "gives the pipe-level bound; `poseidon` below gives permutations/s)",
                 "ntt_pass": "algorithmic bytes per SURVEY 8(d): 16*N*C per plain transform, 48*n*C per iNTT + 4x coset LDE "
                             "(the LDE's re-read of the fresh coefficients is not counted); pass_traffic counts the 16 B per "
                             "element every shared-memory pass moves"}
        line = {"metric": workload_metric(args.workload), "value": value, "unit": seg.unit, "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": t_dev / args.steps, "ms_per_step_instrumented": t_prof / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 (Goldilocks)", "data": "synthetic",
                "config": workload_config(args.workload),
                "launch": {"segments_per_step": NW, "numa_binding": numa,
                           "parallelism": f"{world} GPU(s) x {NW} worker contexts, one independent segment each (a step = "
                                          f"{NW * world} proofs)"},
                "e2e": {"value": e2e, "unit": seg.unit, "ms_per_step": t_e2e / args.steps,
                        "h2d_bytes_per_step": seg.input_bytes * NW, "d2h_bytes_per_step": seg.output_bytes * NW,
                        "host_memory": args.host_memory,
                        "pageable_host_memory": ({"value": pageable_steps * NW / (t_e2e_pageable * 1e-3), "unit": seg.unit, "steps": pageable_steps}
                                                 if t_e2e_pageable else None)},
                "gpu_launches": int(launches),
                "roofline": roof(top[0], top[1], notes.get(top[0], "")),
                "dominant_family": top[0],
                "kernel_families": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                                        "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else 0.0),
                                        "share": v["ms"] / total_ms} for k, v in fam.items()},
                "single_proof_latency_ms": t_single,
                "clocks": clk.summary()}
        if in_segment:
            in_segment["speedup_vs_one_gpu"] = t_single / in_segment["ms_per_proof"]
            line["in_segment"] = in_segment
        if "ntt_pass" in fam:
            line["roofline_ntt"] = roof("ntt_pass", fam["ntt_pass"], notes["ntt_pass"])
        hashing = [fam[k] for k in ("leaf_hash", "merkle_levels", "leaf_hash_rows") if k in fam]
        if hashing:
            perms = sum(v.get("traffic_bytes", 0.0) for v in hashing)          # hashing families report permutations there
            t_hash = sum(v["ms"] for v in hashing)
            line["poseidon"] = {"permutations_per_step": perms / args.steps, "Gperm_per_s": perms / (t_hash * 1e-3) / 1e9 if t_hash else 0.0,
                                "share_of_step": t_hash / total_ms}
        if not args.no_cpu_baseline and world == 1:
            nc = host_threads()
            cb = cpu_pass(args.workload, nc, use_cache=True)
            line["cpu_baseline"] = {"value": cb["value"], "unit": cb["unit"], "cores": nc, "kind": "port",
                                    "sample": cb["sample"], "note": cb["note"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
