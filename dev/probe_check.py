"""FP32 probe vs oracle on box_mirror; PTB200_LIB selects the library variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package
from oracle import Oracle
pkg = load_package(); orc = Oracle("port")
rng = np.random.default_rng(1)
for name in sys.argv[1:] or ["box_mirror"]:
    W, H = 256, 192
    sph, cfg = pkg.builtin_scene(name, W, H); cam = pkg.camera_with_config(cfg)
    with pkg.Renderer(0) as r:
        r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
        n = 200000
        xs, ys = rng.integers(0, W, n), rng.integers(0, H, n); sx, sy = rng.integers(0, 2, n), rng.integers(0, 2, n); ss = rng.integers(0, 1 << 20, n)
        oh, orad, oray, od = orc.samples(sph, cam, W, H, 2, 5, xs, ys, sx, sy, ss)
        h, rad, ray, d = r.trace_samples(5, xs, ys, sx, sy, ss, pkg.PRECISION_FP32)
        rel = np.abs(rad - orad).max(axis=1) / np.maximum(np.abs(orad).max(axis=1), 1e-12)
        se = np.sqrt((rad.var(axis=0) + orad.var(axis=0)) / n)
        print(f"{os.environ.get('PTB200_LIB','default')[-24:]:24s} {name:10s} hit-eq {np.mean(h == oh):.6f} rel<=1e-4 {np.mean(rel <= 1e-4):.5f} "
              f"mean gpu {rad.mean(axis=0)} oracle {orad.mean(axis=0)} z {(rad.mean(axis=0)-orad.mean(axis=0))/se}")
