"""The drop-in claim, end to end: the REFERENCE'S OWN PROGRAM -- src/main.cpp with the
INTEGRATION.md patch (its taskflow row tasks replaced by calls into include/ptb200.h, every GPU of the box behind one context), compiled against the
reference's unmodified headers and translation units -- runs on libptb200.so and writes the image the library's own
host program writes.  oracle/Makefile builds it where /root/reference exists (oracle/ref/make_dropin.py applies the
patch in a scratch directory); the binary travels to the GPU box, the sources do not."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "cpu_path_tracer_b200")
SB_EXE = os.path.join(ROOT, "oracle", "_ref", "smallpt_b200")  # sandbox/main.cpp with the INTEGRATION.md section 2 patch
PTB_SMALLPT = os.path.join(ROOT, "cpu-path-tracing_b200", "ptb_smallpt")
PTB_MAIN = os.path.join(ROOT, "cpu-path-tracing_b200", "ptb_main")

needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/cpu_path_tracer_b200 not built (no /root/reference here)")


def read_ppm(path):
    tok = open(path).read().split()
    assert tok[0] == "P3"
    w, h, mx = int(tok[1]), int(tok[2]), int(tok[3])
    return np.array(tok[4:], dtype=np.int32).reshape(h, w, 3), mx


@needs_exe
def test_reference_program_binds_only_the_abi():
    dyn = subprocess.run(["readelf", "-d", EXE], capture_output=True, text=True, check=True).stdout
    assert "libptb200.so" in dyn
    und = subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True, check=True).stdout
    bound = sorted(l.split()[-1] for l in und.splitlines() if " ptb_" in l)
    assert bound == ["ptb_create_multi", "ptb_destroy", "ptb_device_count", "ptb_last_error", "ptb_render", "ptb_resolve",
                     "ptb_set_camera", "ptb_set_image", "ptb_upload_scene"]
    assert "tf::" not in und and "omp_" not in und  # no task pool left behind


@needs_exe
def test_patch_script_refuses_a_changed_reference(tmp_path):
    """make_dropin.py must fail loudly, not emit a half-patched program."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_dropin", os.path.join(ROOT, "oracle", "ref", "make_dropin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with pytest.raises(SystemExit):
        mod.patch("int main() { return 0; }\n")
    with pytest.raises(SystemExit):
        mod.patch("#include <taskflow/taskflow.hpp>\nint main() { tf::Executor executor{}; }\n")


@needs_exe
@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present: covered by the gpu test below")
def test_reference_program_fails_loudly_without_a_gpu(tmp_path):
    out = subprocess.run([EXE, "4"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 1 and "CUDA" in out.stderr
    assert not os.path.exists(tmp_path / "image.ppm")  # no CPU fallback rendered anything


@needs_exe
@pytest.mark.gpu
def test_reference_program_renders_through_the_library(tmp_path):
    """`cpu_path_tracer_b200 64` (the reference's CLI: total spp) against `ptb_main 64` on the same scene and size."""
    a = subprocess.run([EXE, "64"], cwd=tmp_path, capture_output=True, text=True)
    assert a.returncode == 0, a.stderr
    b = subprocess.run([PTB_MAIN, "64", "--scene", "box_mirror", "--size", "1024x768", "--out", str(tmp_path / "lib.ppm")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    img_a, mx = read_ppm(tmp_path / "image.ppm")
    img_b, _ = read_ppm(tmp_path / "lib.ppm")
    assert mx == 255 and img_a.shape == (768, 1024, 3) == img_b.shape
    # same library, same seed, same samples: only the order of the float additions into a slot differs
    assert np.abs(img_a - img_b).max() <= 1
    assert (img_a == img_b).mean() > 0.999
    assert img_a.mean() > 20  # and it is a picture, not a black frame


@needs_exe
@pytest.mark.gpu
def test_reference_program_on_every_gpu_writes_the_one_gpu_image(tmp_path):
    """The patched program takes every GPU of the box behind ONE context (ptb_create_multi); PTB_GPUS=1 caps it.  The
    stream is keyed by the absolute sample index: the picture must not depend on the GPU count."""
    import ctypes
    n = ctypes.CDLL(os.path.join(ROOT, "cpu-path-tracing_b200", "libptb200.so")).ptb_device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    imgs = []
    for cap in ("1", str(n)):
        d = tmp_path / cap
        d.mkdir()
        out = subprocess.run([EXE, "64"], cwd=d, capture_output=True, text=True, env=dict(os.environ, PTB_GPUS=cap))
        assert out.returncode == 0, out.stderr
        imgs.append(read_ppm(d / "image.ppm")[0])
    assert np.abs(imgs[0] - imgs[1]).max() <= 1 and (imgs[0] == imgs[1]).mean() > 0.999
    assert imgs[0].mean() > 20


needs_sb = pytest.mark.skipif(not os.path.exists(SB_EXE), reason="oracle/_ref/smallpt_b200 not built (no /root/reference here)")


@needs_sb
def test_sandbox_program_binds_only_the_abi():
    und = subprocess.run(["nm", "-D", "--undefined-only", SB_EXE], capture_output=True, text=True, check=True).stdout
    bound = sorted(l.split()[-1] for l in und.splitlines() if " ptb_" in l)
    assert bound == ["ptb_create_multi", "ptb_destroy", "ptb_device_count", "ptb_last_error", "ptb_render", "ptb_resolve",
                     "ptb_set_image", "ptb_set_smallpt_camera", "ptb_upload_scene"]
    # the program's own radiance() is still compiled (dead code now, hence erand48), its thread pool is gone
    assert "omp_" not in und and "GOMP" not in und


@needs_sb
@pytest.mark.gpu
def test_sandbox_program_renders_through_the_library(tmp_path):
    """`smallpt_b200 64` (sandbox/main.cpp's CLI) against the library's own `ptb_smallpt 64`."""
    a = subprocess.run([SB_EXE, "64"], cwd=tmp_path, capture_output=True, text=True)
    assert a.returncode == 0, a.stderr
    os.rename(tmp_path / "image.ppm", tmp_path / "ref_program.ppm")
    b = subprocess.run([PTB_SMALLPT, "64"], cwd=tmp_path, capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    img_a, mx = read_ppm(tmp_path / "ref_program.ppm")
    img_b, _ = read_ppm(tmp_path / "image.ppm")
    assert mx == 255 and img_a.shape == (768, 1024, 3) == img_b.shape
    assert np.abs(img_a - img_b).max() <= 1 and (img_a == img_b).mean() > 0.999
    assert img_a.mean() > 20
