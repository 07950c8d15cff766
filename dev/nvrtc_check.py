"""Compile the sorted megakernel with NVRTC on the CPU (no GPU needed) the way ptb_jit.cpp does, with box_mirror-like
literals: checks that the device headers are NVRTC-clean and reports the compile time and SASS size.
   python dev/nvrtc_check.py [out.cubin]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cpu-path-tracing_b200", "csrc")
n = ctypes.CDLL("libnvrtc.so.12")
HEADERS = ["ptb_types.h", "ptb_rng.cuh", "ptb_scene.cuh", "ptb_kernels.h", "ptb_path_f32.cuh", "ptb_smallpt_f32.cuh",
           "ptb_mega_inplace.cuh", "ptb_mega_sorted.cuh"]
srcs = [open(os.path.join(CSRC, h), "rb").read() for h in HEADERS]


def initialiser(ns, nb):
    small = ["{0x0p+0f, 0x1.99999ap-4f, 0x1.333334p-1f, 0x1.47ae14p-5f}", "{0x1.99999ap-3f, -0x1.99999ap-3f, 0x1.99999ap-2f, 0x1.47ae14p-5f}",
             "{-0x1.99999ap-3f, -0x1.99999ap-3f, 0x1.99999ap-2f, 0x1.47ae14p-5f}", "{0.5f, 0.25f, -0.5f, 0.01f}"]
    big = "{-0.5000002f, 0.0f, 0.0f, 5e-07f, 0.40000007f, 2e6f, 0.0f, 0.0f}"
    axis = ["-0.5000002f, 0.40000007f", "0.5000002f, 0.40000007f", "-0.5000002f, 0.40000007f", "0.5000002f, 0.40000007f", "0.5f, 0.0f", "0.25f, 0.1f"]
    return ("{ { " + ", ".join(small[:max(ns, 1)]) + " }, { " + ", ".join([big] * max(nb, 1)) + " }, { "
            + ", ".join(axis[:max(nb, 1)]) + " }, 0, 0, 0, 0 }")


def compile_kernel(label, header, name, ns, nb, out=None):
    tu = ("#define PTB_JIT_SCENE_INIT " + initialiser(ns, nb) + "\n#include \"ptb_kernels.h\"\n#include \"ptb_path_f32.cuh\"\n"
          "#include \"" + header + "\"\n").encode()
    prog = ctypes.c_void_p()
    hs = (ctypes.c_char_p * len(HEADERS))(*srcs)
    hn = (ctypes.c_char_p * len(HEADERS))(*[h.encode() for h in HEADERS])
    assert n.nvrtcCreateProgram(ctypes.byref(prog), tu, b"ptb_jit_tu.cu", len(HEADERS), hs, hn) == 0
    assert n.nvrtcAddNameExpression(prog, name) == 0
    opts = (ctypes.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo")
    t0 = time.time()
    rc = n.nvrtcCompileProgram(prog, 3, opts)
    dt = time.time() - t0
    sz = ctypes.c_size_t(); n.nvrtcGetProgramLogSize(prog, ctypes.byref(sz))
    log = ctypes.create_string_buffer(sz.value); n.nvrtcGetProgramLog(prog, log)
    print(label, "compile rc", rc, "in %.2f s" % dt)
    if log.value.strip():
        print(log.value.decode()[:3000])
    if rc == 0:
        low = ctypes.c_char_p(); n.nvrtcGetLoweredName(prog, name, ctypes.byref(low)); print("  kernel:", low.value.decode())
        n.nvrtcGetCUBINSize(prog, ctypes.byref(sz)); print("  cubin bytes", sz.value)
        if out:
            buf = ctypes.create_string_buffer(sz.value); n.nvrtcGetCUBIN(prog, buf); open(out, "wb").write(buf.raw)
    return rc


rc = compile_kernel("sorted / box", "ptb_mega_sorted.cuh",
                    b"ptb::mega_sorted_kernel<ptb::SceneShape<2, 1, 5, 0, 2, 2, 1, true, true, 3>, true, 1>", 3, 5,
                    sys.argv[1] if len(sys.argv) > 1 else None)
rc |= compile_kernel("in-place / box", "ptb_mega_inplace.cuh",
                     b"ptb::mega_kernel<ptb::SceneShape<2, 1, 5, 0, 2, 2, 1, true, true, 3>, true, ptb::IntegratorPt>", 3, 5)
rc |= compile_kernel("in-place smallpt / sandbox", "ptb_mega_inplace.cuh",
                     b"ptb::mega_kernel<ptb::SceneShape<3, 1, 0, 6, 0, 0, 0, true, false, 0>, true, ptb::IntegratorSmallpt>", 4, 6)
sys.exit(rc)
