"""The C-ABI boundary: header, exported symbols, argument/error behaviour that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(pkg):
    declared = _header_symbols()
    assert declared == sorted(pkg.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ptb200.h but not exported by libptb200.so"


def test_abi_version(pkg):
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    assert pkg.abi_version() == int(re.search(r"#define PTB_ABI_VERSION (\d+)", text).group(1))


def test_record_sizes_match_header(pkg):
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    for macro, value in (("PTB_SPHERE_BYTES", pkg.SPHERE_BYTES), ("PTB_CAMERA_BYTES", pkg.CAMERA_BYTES),
                         ("PTB_CAMERA_CONFIG_BYTES", pkg.CAMERA_CONFIG_BYTES)):
        assert int(re.search(rf"#define {macro} (\d+)", text).group(1)) == value
    assert pkg.SPHERE_DTYPE.itemsize == 88 and pkg.SPHERE_DTYPE.fields["reflection"][1] == 80
    assert pkg.CAMERA_DTYPE.fields["lens_radius"][1] == 168
    assert pkg.CAMERA_CONFIG_DTYPE.fields["focus_distance"][1] == 104


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product refuses to run; it never computes on the CPU."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PtbError) as e:
        pkg.Renderer(0)
    assert e.value.code == -2  # PTB_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "cpu-path-tracing_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("//", "#", "*", "/*")))
                assert "pt_oracle" not in code and "libptref" not in code and "import oracle" not in code, f


def test_binding_constants_are_the_header_s(pkg):
    """Every PTB_X = value of include/ptb200.h that the ctypes binding mirrors (flags, transports, sizes) has that value there."""
    import re
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    header = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"\bPTB_([A-Z0-9_]+)\s*=\s*(-?(?:0x[0-9A-Fa-f]+|\d+))", text)}
    header.update({m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+PTB_([A-Z0-9_]+)\s+(\d+)\b", text)})
    mirrored = [n for n in header if hasattr(pkg, n) and isinstance(getattr(pkg, n), int)]
    assert {"VARIANT_MEGAKERNEL_SORTED", "PRECISION_FP64", "INTEGRATOR_SMALLPT", "ACCEL_SCAN", "CODEGEN_PRECOMPILED", "TRANSPORT_PEER",
            "SPHERE_BYTES", "CAMERA_BYTES", "COMM_ID_BYTES"} <= set(mirrored)
    for n in mirrored:
        assert getattr(pkg, n) == header[n], n
    assert header["ERR_MEMORY"] == -6 and header["ERR_INTERNAL"] == -7 and header["OK"] == 0


def test_no_cxx_exception_crosses_the_boundary(pkg, tmp_path):
    """Every status-returning entry point is a function-try-block (ptb_context.hpp: PTB_CATCH): a std::length_error /
    std::bad_alloc inside the library comes back as a status, it does not unwind into a C / Go / ctypes caller."""
    lib = ctypes.CDLL(pkg.LIB_PATH)
    one = np.zeros(3, dtype=np.uint8)
    out = tmp_path / "never.ppm"
    # a 2^30 x 2^30 image as P3 text: 12 * 2^60 bytes, more than a std::vector can ever hold -- thrown before anything is read
    rc = lib.ptb_write_ppm_rgb8(os.fsencode(str(out)), one.ctypes.data_as(ctypes.c_void_p), 1 << 30, 1 << 30, 0)
    assert rc in (-6, -7)  # PTB_ERR_MEMORY / PTB_ERR_INTERNAL
    assert not out.exists()  # and no file was opened for an image that could not be formatted


def test_host_helpers_reject_bad_arguments(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    assert lib.ptb_camera_with_config(None, None) == -1
    n = ctypes.c_size_t(0)
    assert lib.ptb_builtin_scene(b"no_such_scene", 4, 4, None, 0, ctypes.byref(n), None) == -1
    assert lib.ptb_builtin_scene(b"box", 0, 4, None, 0, ctypes.byref(n), None) == -1
    buf = np.zeros(88, dtype=np.uint8)  # room for one sphere, the scene has eight
    assert lib.ptb_builtin_scene(b"box", 4, 4, buf.ctypes.data_as(ctypes.c_void_p), 1, ctypes.byref(n), None) == -1
    assert n.value == 8
    assert lib.ptb_write_ppm(None, None, 1, 1) == -1
    # every context call on a null context is an argument error, not a crash
    for name in ("ptb_clear", "ptb_synchronize"):
        assert getattr(lib, name)(None) == -1
    assert lib.ptb_render(None, ctypes.c_uint64(0), 0, 1, 0) == -1
    assert lib.ptb_resolve(None, None) == -1
