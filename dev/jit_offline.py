"""Recompile a dumped run-time translation unit (dev/jit_tu/<scene>.cu, made on a GPU box by dev/jit_dump.py) against the
WORKING TREE's headers with NVRTC -- no GPU needed -- and report registers, code size and the instruction count of the
bounce loop, optionally with extra macro definitions:

    python dev/jit_offline.py box_mirror [-DPTB_X=1 ...] [--sass out.sass]

The bounce loop = everything between the closest-hit scan's first instruction after the slow path and the loop's backward
branch; reported as the number of SASS instructions in address ranges (cheap proxy: static counts, not executed counts)."""
import ctypes, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cpu-path-tracing_b200", "csrc")
HEADERS = ["ptb_types.h", "ptb_rng.cuh", "ptb_scene.cuh", "ptb_kernels.h", "ptb_path_f32.cuh", "ptb_smallpt_f32.cuh",
           "ptb_mega_inplace.cuh", "ptb_mega_sorted.cuh"]


def compile_tu(scene, defines, cubin_path):
    n = ctypes.CDLL("libnvrtc.so.12")
    srcs = [open(os.path.join(CSRC, h), "rb").read() for h in HEADERS]
    tu = open(os.path.join(ROOT, "dev", "jit_tu", scene + ".cu"), "rb").read()
    name = open(os.path.join(ROOT, "dev", "jit_tu", scene + ".name"), "rb").read().strip()
    prog = ctypes.c_void_p()
    hs = (ctypes.c_char_p * len(HEADERS))(*srcs)
    hn = (ctypes.c_char_p * len(HEADERS))(*[h.encode() for h in HEADERS])
    assert n.nvrtcCreateProgram(ctypes.byref(prog), tu, b"ptb_jit_tu.cu", len(HEADERS), hs, hn) == 0
    assert n.nvrtcAddNameExpression(prog, name) == 0
    o = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo"] + [d.encode() for d in defines]
    opts = (ctypes.c_char_p * len(o))(*o)
    rc = n.nvrtcCompileProgram(prog, len(o), opts)
    sz = ctypes.c_size_t(); n.nvrtcGetProgramLogSize(prog, ctypes.byref(sz))
    log = ctypes.create_string_buffer(sz.value); n.nvrtcGetProgramLog(prog, log)
    if rc != 0:
        print(log.value.decode()[:4000]); sys.exit(1)
    n.nvrtcGetCUBINSize(prog, ctypes.byref(sz))
    buf = ctypes.create_string_buffer(sz.value); n.nvrtcGetCUBIN(prog, buf)
    open(cubin_path, "wb").write(buf.raw)


def sass_of(cubin):
    txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    rows = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    return rows


def res_usage(cubin):
    txt = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True, check=True).stdout
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", txt)
    return (int(m.group(1)), int(m.group(2))) if m else (None, None)


if __name__ == "__main__":
    scene = sys.argv[1]
    defines = [a for a in sys.argv[2:] if a.startswith("-D")]
    out = None
    if "--sass" in sys.argv:
        out = sys.argv[sys.argv.index("--sass") + 1]
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "k.cubin")
        compile_tu(scene, defines, cubin)
        rows = sass_of(cubin)
        regs, smem = res_usage(cubin)
    if out:
        with open(out, "w") as f:
            for a, ins in rows:
                f.write(f"{a:04x}  {ins}\n")
    ops = {}
    for _, ins in rows:
        op = ins.split()[1].split(".")[0] if ins.startswith("@") else ins.split()[0].split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    # the bounce loop: from the target of the LAST backward branch to that branch
    back = [(a, int(re.search(r"0x([0-9a-f]+)", ins).group(1), 16)) for a, ins in rows if re.match(r"(@!?U?P\d+ )?BRA", ins) and re.search(r"0x([0-9a-f]+)", ins)]
    back = [(a, t) for a, t in back if t < a]
    print(f"{scene} {' '.join(defines)}: {len(rows)} instructions, {regs} registers, {smem} B shared")
    if back:
        a, t = max(back, key=lambda p: p[0] - p[1])
        body = [ins for ad, ins in rows if t <= ad <= a]
        print(f"  widest loop 0x{t:04x}-0x{a:04x}: {len(body)} instructions")
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:18]
    if "--ops" in sys.argv:
        print("  " + ", ".join(f"{k} {v}" for k, v in top))
