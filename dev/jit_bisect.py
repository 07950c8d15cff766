"""Which ingredient of a fuzz scene makes the run-time build differ from the precompiled kernels?
   python dev/jit_bisect.py <base> <scene>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fuzz_scenes import make_scene, pkg
base, it = int(sys.argv[1]), int(sys.argv[2])
s0, cfg0, cam0, W, H, S, n, style, hostile = make_scene(base, it)
F32 = pkg.PRECISION_FP32


def run(label, s, cfg):
    cam = pkg.camera_with_config(cfg)
    out = []
    with pkg.Renderer(0) as r:
        r.upload_scene(s); r.set_camera(cam); r.set_image(W, H, 2)
        lay = r.scene_layout()
        for flags, reps in ((F32 | pkg.VARIANT_MEGAKERNEL_SORTED | pkg.CODEGEN_PRECOMPILED, 1), (F32 | pkg.VARIANT_MEGAKERNEL_SORTED, 2),
                            (F32 | pkg.VARIANT_MEGAKERNEL, 2)):
            for _ in range(reps):
                r.clear(); r.render(7 + it, 0, S, flags)
            out.append((r.stats().rays, float(r.resolve().mean()), r.jit_info()["last_launch_jit"]))
    print(f"{label:34s} pre {out[0][:2]}  jit-sorted {out[1]}  jit-inplace {out[2]}  layout {lay}", flush=True)


run("as generated", s0, cfg0)
for i in range(n):
    if s0["radius"][i] == 0.0 or s0["radius"][i] < 1e-5:
        s = s0.copy(); s["radius"][i] = 0.05
        run(f"sphere {i}: radius {s0['radius'][i]} -> 0.05", s, cfg0)
s = s0.copy(); s["color"] = np.minimum(s["color"], 1.0); run("colours <= 1", s, cfg0)
s = s0.copy(); s["emission"] = np.minimum(s["emission"], 8.0); run("emission <= 8", s, cfg0)
cfg = cfg0.copy(); cfg["aperture"] = 0.0; run("pinhole", s0, cfg)
cfg = cfg0.copy(); cfg["position"][0] += (0, 2.0, 0); run("camera 2 up", s0, cfg)
for i in range(n):
    s = np.delete(s0, i)
    run(f"without sphere {i} (r={s0['radius'][i]:.3g}, refl {s0['reflection'][i]})", s, cfg0)
