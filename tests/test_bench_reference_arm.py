"""bench.py --impl reference end to end on a test-sized workload (CPU: the arm runs the oracle on the host cores and needs no
device): the JSON line carries the contract's keys, describes the run it actually made and is reproducible (same proof digest)."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _line(*args):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    import bench
    a = _line("--workload", "U10", "--steps", "1", "--warmup", "0")
    assert a["impl"] == "reference" and a["n_gpus"] == 1 and a["higher_is_better"] is True and a["unit"] == "proofs/s"
    assert a["config"]["workload"] == "U10" and a["config"]["log_heights"] == bench.workload_log_heights("U10")
    assert a["metric"] == bench.workload_metric("U10")
    assert a["steps"] == 1 and abs(a["value"] * a["ms_per_step"] / 1e3 - 1.0) < 1e-6          # value = proofs / measured time
    cb = a["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == bench.host_threads() and cb["value"] == a["value"] and cb["sample"].startswith("full")
    assert a["e2e"] == {"value": a["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert a["gpu_launches"] == 0 and a["vs_baseline"] is None and "FIT in the 126 MB L2" in a["config"]["l2"]
    b = _line("--workload", "U10", "--steps", "1", "--warmup", "0")
    assert b["cpu_baseline"]["proof_sha256"] == cb["proof_sha256"]                              # deterministic prover, same inputs
