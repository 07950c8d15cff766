"""Compile the sorted megakernel with NVRTC on the CPU (no GPU needed) the way ptb_jit.cpp does, with box_mirror-like
literals: checks that the device headers are NVRTC-clean and reports the compile time and SASS size.
   python dev/nvrtc_check.py [out.cubin]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cpu-path-tracing_b200", "csrc")
HEADERS = ["ptb_types.h", "ptb_rng.cuh", "ptb_scene.cuh", "ptb_kernels.h", "ptb_path_f32.cuh", "ptb_mega_sorted.cuh"]
n = ctypes.CDLL("libnvrtc.so.12")
init = ("{ { {0x0p+0f, 0x1.99999ap-4f, 0x1.333334p-1f, 0x1.47ae14p-5f}, {0x1.99999ap-3f, -0x1.99999ap-3f, 0x1.99999ap-2f, 0x1.47ae14p-5f}, "
        "{-0x1.99999ap-3f, -0x1.99999ap-3f, 0x1.99999ap-2f, 0x1.47ae14p-5f}, }, { "
        + "{0,0,0,5e-07f,0,0,0,0}, " * 5 + "}, { -0.5000002f, 0.40000007f, 0.5000002f, 0.40000007f, -0.5000002f, 0.40000007f, 0.5000002f, 0.40000007f, 0.5f, 0.0f, }, 0, 0, 0, 0 }")
tu = ("#define PTB_JIT_SCENE_INIT " + init + "\n#include \"ptb_kernels.h\"\n#include \"ptb_path_f32.cuh\"\n"
      "namespace ptb { __constant__ ConstSceneF32 c_scene; }\n#include \"ptb_mega_sorted.cuh\"\n").encode()
srcs = [open(os.path.join(CSRC, h), "rb").read() for h in HEADERS]
prog = ctypes.c_void_p()
hs = (ctypes.c_char_p * len(HEADERS))(*srcs)
hn = (ctypes.c_char_p * len(HEADERS))(*[h.encode() for h in HEADERS])
assert n.nvrtcCreateProgram(ctypes.byref(prog), tu, b"ptb_jit_tu.cu", len(HEADERS), hs, hn) == 0
name = b"ptb::mega_sorted_kernel<ptb::SceneShape<2, 1, 5, 0, 2, 2, 1, true, true, 3>, true, 1>"
assert n.nvrtcAddNameExpression(prog, name) == 0
assert n.nvrtcAddNameExpression(prog, b"&ptb::c_scene") == 0
opts = (ctypes.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo")
t0 = time.time()
rc = n.nvrtcCompileProgram(prog, 3, opts)
dt = time.time() - t0
sz = ctypes.c_size_t(); n.nvrtcGetProgramLogSize(prog, ctypes.byref(sz))
log = ctypes.create_string_buffer(sz.value); n.nvrtcGetProgramLog(prog, log)
print("compile rc", rc, "in %.2f s" % dt)
if log.value.strip():
    print(log.value.decode()[:3000])
if rc == 0:
    low = ctypes.c_char_p(); n.nvrtcGetLoweredName(prog, name, ctypes.byref(low)); print("kernel:", low.value.decode())
    n.nvrtcGetLoweredName(prog, b"&ptb::c_scene", ctypes.byref(low)); print("c_scene:", low.value.decode())
    n.nvrtcGetCUBINSize(prog, ctypes.byref(sz)); print("cubin bytes", sz.value)
    if len(sys.argv) > 1:
        buf = ctypes.create_string_buffer(sz.value); n.nvrtcGetCUBIN(prog, buf); open(sys.argv[1], "wb").write(buf.raw)
sys.exit(rc)
