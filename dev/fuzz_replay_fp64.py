"""Replay one scene of dev/fuzz_scenes.py through the FP64 parity kernel against the oracle port and list the pixels that differ:
   python dev/fuzz_replay_fp64.py <seed base> <scene number>"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dev"))
import fuzz_scenes as fz
from oracle import Oracle
pkg = fz.pkg
base, it = int(sys.argv[1]), int(sys.argv[2])
s, cfg, cam, W, H, S, n, style, hostile = fz.make_scene(base, it)
ns = [2, 2, 1, 3, 5][it % 5]
if ns > 2:
    W, H = max(1, W // 2), max(1, H // 2)
ref = Oracle("port").render(s, cam, W, H, S, ns, 7 + it, 0)
with pkg.Renderer(0) as r:
    r.upload_scene(s); r.set_camera(cam); r.set_image(W, H, ns)
    r.render(7 + it, 0, S, pkg.PRECISION_FP64)
    img = r.resolve()
d = np.abs(img - ref)
rel = d / np.maximum(np.abs(ref), 1e-30)
bad = np.argwhere(d.max(axis=2) > 1e-6)
print(W, H, S, ns, "pixels off:", len(bad), "max abs", d.max(), "max rel", rel[d > 1e-6].max() if len(bad) else 0)
for y, x in bad[:12]:
    print(y, x, img[y, x], ref[y, x])
print("emissive spheres:", [(i, float(s["emission"][i].max()), float(s["radius"][i])) for i in range(n) if s["emission"][i].max() > 10][:5])
