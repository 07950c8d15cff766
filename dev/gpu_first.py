"""First contact with the GPU: smoke, parity statistics, a timed box_mirror render."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package, smoke
from oracle import Oracle

pkg = load_package()
try:
    smoke()
except AssertionError as e:
    print("SMOKE FAILED:", e)
orc = Oracle("port")
rng = np.random.default_rng(1)
for name in ["simple", "box", "box_mirror", "dof_glass"]:
    W, H = 256, 192
    sph, cfg = pkg.builtin_scene(name, W, H)
    cam = pkg.camera_with_config(cfg)
    with pkg.Renderer(0) as r:
        r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
        n = 100000
        xs, ys = rng.integers(0, W, n), rng.integers(0, H, n)
        sx, sy = rng.integers(0, 2, n), rng.integers(0, 2, n)
        ss = rng.integers(0, 1 << 20, n)
        oh, orad, oray, od = orc.samples(sph, cam, W, H, 2, 5, xs, ys, sx, sy, ss)
        for prec, pname in [(pkg.PRECISION_FP64, "fp64"), (pkg.PRECISION_FP32, "fp32")]:
            h, rad, ray, d = r.trace_samples(5, xs, ys, sx, sy, ss, prec)
            rel = np.abs(rad - orad).max(axis=1) / np.maximum(np.abs(orad).max(axis=1), 1e-12)
            print(f"{name:10s} {pname}: hit-eq {np.mean(h == oh):.6f} exact-rad {np.mean((rad == orad).all(axis=1)):.4f} "
                  f"rel<=1e-4 {np.mean(rel <= 1e-4):.5f} rel<=1e-2 {np.mean(rel <= 1e-2):.5f} median-rel {np.median(rel):.2e} "
                  f"ray-maxabs {np.abs(ray - oray).max():.2e} mean-rad gpu {rad.mean():.5f} oracle {orad.mean():.5f}"
                  + (f" draws-eq {np.mean(d == od):.5f}" if prec == pkg.PRECISION_FP64 else ""))
        # image comparison, both precisions
        S = 4
        ref = orc.render(sph, cam, W, H, S, 2, 9)
        for prec, pname in [(pkg.PRECISION_FP64, "fp64"), (pkg.PRECISION_FP32, "fp32")]:
            r.clear(); r.render(9, 0, S, prec); img = r.resolve(); st = r.stats()
            print(f"{name:10s} {pname} image: mean|diff| {np.abs(img-ref).mean():.3e} max {np.abs(img-ref).max():.3e} "
                  f"mean img {img.mean():.5f} ref {ref.mean():.5f} rays/path {st.rays/st.paths:.3f} ms {st.last_render_ms:.2f}")

# timing
for name, W, H, S in [("box_mirror", 1920, 1080, 16), ("box_mirror", 1920, 1080, 64), ("box", 1024, 768, 64), ("simple", 1024, 768, 64),
                      ("dof_glass", 1920, 1080, 16)]:
    sph, cfg = pkg.builtin_scene(name, W, H)
    cam = pkg.camera_with_config(cfg)
    with pkg.Renderer(0) as r:
        r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
        for vname, flag in (("mega", pkg.VARIANT_MEGAKERNEL), ("wave", pkg.VARIANT_WAVEFRONT)):
            r.render(1, 0, 4, flag)
            r.clear()
            r.render(1, 0, S, flag); st = r.stats()
            paths = W * H * 4 * S
            print(f"TIMING {vname} {name} {W}x{H} samps/subpixel {S}: {st.last_render_ms:.2f} ms  {paths/st.last_render_ms/1e3:.1f} Mpaths/s "
                  f"{st.rays/st.last_render_ms/1e3:.1f} Mrays/s rays/path {st.rays/paths:.3f}")

# ---- sandbox (stand-alone smallpt) mode ------------------------------------------------------------------
sph, cam8 = pkg.builtin_smallpt_scene()
W, H = 1024, 768
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_smallpt_camera(cam8); r.set_image(W, H, 2)
    print("smallpt layout", r.scene_layout())
    for S in (4, 64):
        r.clear(); r.render(1, 0, S, pkg.INTEGRATOR_SMALLPT); st = r.stats()
        print(f"TIMING smallpt {W}x{H} samps/subpixel {S}: {st.last_render_ms:.2f} ms {W*H*4*S/st.last_render_ms/1e3:.1f} Mpaths/s "
              f"{st.rays/st.last_render_ms/1e3:.1f} Mrays/s rays/path {st.rays/st.paths:.3f}")
