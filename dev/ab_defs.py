"""A/B timing of #ifdef'ed experiments in the run-time compiled kernels with ONE library, alternating processes on the same box:
   python dev/ab_defs.py <scene> <W> <H> <samples> <flags> <rounds> "name=-DPTB_X -DPTB_Y" "base=" "old=@path/to/other/libptb200.so -DPTB_Z" ...
(a leading @path selects another build of the library for that variant)
Every process renders three times (the run-time build happens on the second call) and reports the third render's device time;
also prints rays traced so that variants that change the paths show up."""
import os, subprocess, sys, statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
from __graft_entry__ import load_package
pkg = load_package()
name, W, H, S, flags = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5], 0)
with pkg.Renderer(0) as r:
    if name == "smallpt":
        sph, cam8 = pkg.builtin_smallpt_scene()
        r.upload_scene(sph); r.set_smallpt_camera(cam8)
    else:
        sph, cfg = pkg.builtin_scene(name, W, H)
        r.upload_scene(sph); r.set_camera(pkg.camera_with_config(cfg))
    r.set_image(W, H, 2)
    for _ in range(3):
        r.clear()
        r.render(1, 0, S, flags)
    st = r.stats()
    print(st.last_render_ms, st.rays, r.jit_info()["last_launch_jit"])
''' % ROOT

scene, W, H, S, flags, rounds = sys.argv[1:7]
variants = [v.split("=", 1) for v in sys.argv[7:]]
times = {n: [] for n, _ in variants}
rays = {}
for _ in range(int(rounds)):
    for n, d in variants:
        lib, defs = None, d
        if d.startswith("@"):
            lib, _, defs = d[1:].partition(" ")
        env = dict(os.environ, PTB_JIT_DEFINES=defs)
        if lib:
            env["PTB200_LIB"] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, "-c", CHILD, scene, W, H, S, flags], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(n, "FAILED", out.stderr[-400:])
            continue
        ms, ry, jit = out.stdout.strip().splitlines()[-1].split()
        times[n].append(float(ms)); rays[n] = (int(ry), int(jit))
base = statistics.median(times[variants[0][0]]) if times[variants[0][0]] else float("nan")
for n, d in variants:
    t = times[n]
    if t:
        print(f"{n:24s} median {statistics.median(t):9.3f} ms  min {min(t):9.3f}  ({statistics.median(t) / base:.4f} of first)  rays {rays[n][0]} jit {rays[n][1]}  [{d}]")
