import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package
from oracle import Oracle
pkg = load_package(); orc = Oracle("port")
sph, cam8 = pkg.builtin_smallpt_scene()
W, H = 256, 192
rng = np.random.default_rng(5); n = 100000
xs, ys = rng.integers(0, W, n), rng.integers(0, H, n); sx, sy, ss = rng.integers(0, 2, n), rng.integers(0, 2, n), rng.integers(0, 1 << 20, n)
oh, orad, oray, od = orc.sb_samples(sph, cam8, W, H, 4, xs, ys, sx, sy, ss)
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_smallpt_camera(cam8); r.set_image(W, H, 2)
    for prec, nm in ((pkg.PRECISION_FP64, "fp64"), (pkg.PRECISION_FP32, "fp32")):
        h, rad, ray, d = r.trace_samples(4, xs, ys, sx, sy, ss, prec | pkg.INTEGRATOR_SMALLPT)
        rel = np.abs(rad - orad).max(axis=1) / np.maximum(np.abs(orad).max(axis=1), 1e-12)
        bad = np.where(rel > 1e-3)[0]
        print(nm, "hit-eq", (h == oh).mean(), "rel<=1e-4", (rel <= 1e-4).mean(), "rel<=1e-3", (rel <= 1e-3).mean(), "mean", rad.mean(axis=0), "oracle", orad.mean(axis=0),
              "ray maxabs", np.abs(ray - oray).max())
        if nm == "fp32":
            print("  bad samples: primary hit histogram", np.bincount(oh[bad] + 1, minlength=12), "all:", np.bincount(oh + 1, minlength=12))
            print("  oracle draws for bad: mean", od[bad].mean(), "all", od.mean())
        r.clear(); r.render(4, 0, 4, prec | pkg.INTEGRATOR_SMALLPT); st = r.stats()
        print("  render rays/path", st.rays / st.paths, "d/s/g per path", st.hits_diffuse / st.paths, st.hits_specular / st.paths, st.hits_dielectric / st.paths, "ms", st.last_render_ms)
