"""bench.py --impl reference: the CPU arm the driver times next to the GPU arm (no GPU involved, so it is tested here).
It runs the reference's own render loop (oracle/_ref, built from /root/reference by oracle/Makefile) -- or the C restatement
where that library is absent -- on all host threads and prints the contract's JSON line; ranks other than 0 of a torchrun
launch exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    out = _run(None, "--config", "C1", "--steps", "1", "--warmup", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1  # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" == d["unit"] and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C1: simple 1024x768")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "spp per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f64"


def test_reference_arm_runs_on_rank_zero_only():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--config", "C1", "--steps", "1", "--warmup", "3")
    assert out.returncode == 0 and out.stdout.strip() == ""
