"""Compare two ptb_render variants on the same seed: python dev/cmp_variants.py <scene> W H S flagsA flagsB"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

pkg = load_package()
name, W, H, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fa, fb = int(sys.argv[5], 0), int(sys.argv[6], 0)
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
out = []
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
    for f in (fa, fb):
        r.clear()
        r.render(23, 0, S, f)
        out.append((r.download_accum(), r.stats()))
(a, sa), (b, sb) = out
print(name, "counts ok:", bool(np.all(a[:, 3] == S)), bool(np.all(b[:, 3] == S)), "bad slots:", int((b[:, 3] != S).sum()),
      "sum n:", a[:, 3].sum(), b[:, 3].sum())
for k in ("paths", "rays", "hits_diffuse", "hits_specular", "hits_dielectric"):
    print(f"  {k}: {getattr(sa, k)} {getattr(sb, k)}")
print("  non-finite slots:", int((~np.isfinite(a)).any(axis=1).sum()), int((~np.isfinite(b)).any(axis=1).sum()))
for i in np.nonzero((~np.isfinite(a)).any(axis=1))[0][:4]:
    print("   non-finite slot", i, a[i], b[i])
d = np.abs(a[:, :3] - b[:, :3])
rel = d / np.maximum(np.abs(a[:, :3]), 1e-3)
print("  max abs diff", d.max(), "max rel", rel.max(), "frac slots > 1e-5 rel:", (rel.max(axis=1) > 1e-5).mean(),
      "frac > 1e-3:", (rel.max(axis=1) > 1e-3).mean())
bad = np.argsort(-rel.max(axis=1))[:5]
for i in bad:
    print("   slot", i, a[i], b[i])
