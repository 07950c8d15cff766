"""The run-time specialisation of the sorted megakernel (cpu-path-tracing_b200/csrc/ptb_jit.cpp) hands the device
headers to NVRTC.  NVRTC needs no GPU: compile the same translation unit here, so that a header change that breaks the
run-time build (a host-only include, a missing guard) is caught on the CPU."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nvrtc():
    for n in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return ctypes.CDLL(n)
        except OSError:
            pass
    return None


@pytest.mark.skipif(_nvrtc() is None, reason="libnvrtc not installed")
def test_device_headers_compile_under_nvrtc(tmp_path):
    cubin = tmp_path / "k.cubin"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "dev", "nvrtc_check.py"), str(cubin)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mega_sorted_kernel" in out.stdout
    # a run-time compiled kernel shares no __constant__ symbol with anybody: coefficients are literals, cameras ride in
    # the kernel argument (so contexts on one GPU need no lock for it)
    assert "c_scene" not in out.stdout
    assert cubin.stat().st_size > 10000
    # the point of the exercise: with literal coefficients the closest-hit scan has (almost) no constant-bank loads left
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if "/*" in l and ";" in l]
    first_sqrt = next(i for i, l in enumerate(lines) if "MUFU.SQRT" in l and i > 500)
    window = lines[first_sqrt - 50:first_sqrt + 90]
    assert sum("LDC" in l for l in window) <= 4


def test_embedded_sources_are_the_sources(tmp_path):
    """gen_jit_sources.py must embed the headers verbatim (the run-time build and the precompiled build are one source)."""
    out = tmp_path / "s.inc"
    csrc = os.path.join(ROOT, "cpu-path-tracing_b200", "csrc")
    subprocess.run([sys.executable, os.path.join(csrc, "gen_jit_sources.py"), csrc, str(out)], check=True)
    text = out.read_text()
    joined = text.replace(')PTBJIT"\n    R"PTBJIT(', "")  # long headers are split into chunks the compiler concatenates
    for h in ("ptb_types.h", "ptb_rng.cuh", "ptb_scene.cuh", "ptb_kernels.h", "ptb_path_f32.cuh", "ptb_mega_sorted.cuh"):
        assert '"%s"' % h in text
        body = open(os.path.join(csrc, h)).read()
        assert body in joined
