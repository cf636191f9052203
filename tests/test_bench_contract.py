"""CPU test of bench.py's reference arm: exactly one JSON line on stdout with the keys the driver reads (the GPU arm
prints the same line plus roofline / clocks / gpu_launches; it is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env.pop("RANK", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scan-to-map registrations/s" and d["unit"] == "registrations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("C-3")
    cb = d["cpu_baseline"]
    # "reference" = the reference's own laserMapping node out of oracle/_ref (built where the reference tree exists, and
    # travelling with the snapshot); "port" = the oracle restatement, when no such build is at hand
    import oracle_lib
    want = "reference" if oracle_lib.ref_lib("mapping") is not None else "port"
    assert cb["kind"] == want and cb["cores"] >= 1 and cb["value"] == d["value"] and "registrations" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
