"""Host logic of the closest-hit hierarchy: the builder of
cpu-path-tracing_b200/csrc/ptb_bvh.hpp and a CPU restatement of the device traversal, checked against a linear
scan with the same sphere test.  No GPU: tests/host/bvh_check.cpp is compiled with g++ and run."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvh_builder_and_traversal_match_the_linear_scan(tmp_path):
    exe = tmp_path / "bvh_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests", "host", "bvh_check.cpp"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "2000", "20000"], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok ") == 5
    # the point of the hierarchy: tens of sphere tests per ray instead of 10 000
    line = [l for l in out.stdout.splitlines() if l.startswith("ok spheres10k")][0]
    tests_per_ray = float(line.split("node visits and ")[1].split()[0])
    assert tests_per_ray < 40.0
