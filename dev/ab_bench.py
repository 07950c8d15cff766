"""A/B timing of library builds on the same GPU box, alternating processes:
   python dev/ab_bench.py <scene> <W> <H> <samples> <flags> <rounds> libA.so libB.so [...]
Each process renders once to warm up (module load, clocks) and reports the second render's device time."""
import os, subprocess, sys, statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
from __graft_entry__ import load_package
pkg = load_package()
name, W, H, S, flags = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5], 0)
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
    r.render(1, 0, S, flags)
    r.clear()
    r.render(1, 0, S, flags)
    print(r.stats().last_render_ms)
''' % ROOT

scene, W, H, S, flags, rounds = sys.argv[1:7]
libs = sys.argv[7:]
times = {l: [] for l in libs}
for _ in range(int(rounds)):
    for l in libs:
        env = dict(os.environ, PTB200_LIB=os.path.abspath(l))
        out = subprocess.run([sys.executable, "-c", CHILD, scene, W, H, S, flags], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(l, "FAILED", out.stderr[-300:])
            continue
        times[l].append(float(out.stdout.strip().splitlines()[-1]))
base = statistics.median(times[libs[0]]) if times[libs[0]] else float("nan")
for l in libs:
    t = times[l]
    if t:
        print(f"{os.path.basename(l):32s} median {statistics.median(t):8.3f} ms  min {min(t):8.3f}  ({statistics.median(t) / base:.4f} of first)  n={len(t)}")
