"""Camera inside a big sphere (both-roots big list): run-time build against the precompiled kernels and the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from oracle import Oracle
oracle = Oracle("port")
F32 = pkg.PRECISION_FP32
W, H, S = 64, 48, 4
_, cfg = pkg.builtin_scene("simple", W, H)
cfg["position"][0] = (0.0, -0.52, 0.5); cfg["direction"][0] = (0.0, -0.6, -1.0); cfg["aperture"][0] = 0.0


def scene(kind):
    rows = [(1000.0, (0, -1000.5, -1.5), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)]
    if kind >= 1:
        rows.append((0.2, (0.3, 0.0, -1.0), (4, 4, 4), (0.7, 0.7, 0.7), 0, 0))
    if kind >= 2:
        rows.append((0.25, (-0.4, -0.2, -1.2), (0, 0, 0), (0.9, 0.9, 0.9), 2, 0))
    if kind >= 3:
        rows.append((0.3, (0.0, 0.6, -1.6), (0, 0, 0), (0.9, 0.9, 0.9), 1, 0))
    s = np.zeros(len(rows), dtype=pkg.SPHERE_DTYPE)
    for i, rw in enumerate(rows):
        s[i] = rw
    return s


for kind in range(4):
    for emit in (0.0, 2.0):
        s = scene(kind)
        s["emission"][0] = emit
        cam = pkg.camera_with_config(cfg)
        ref = oracle.render(s, cam, W, H, S, 2, 5, 0)
        with pkg.Renderer(0) as r:
            r.upload_scene(s); r.set_camera(cam); r.set_image(W, H, 2)
            out = []
            for flags, reps in ((F32 | pkg.VARIANT_MEGAKERNEL_SORTED | pkg.CODEGEN_PRECOMPILED, 1), (F32 | pkg.VARIANT_MEGAKERNEL_SORTED, 2)):
                for _ in range(reps):
                    r.clear(); r.render(5, 0, S, flags)
                st = r.stats(); im = r.resolve()
                out.append((st.rays, st.hits_diffuse, st.hits_specular, st.hits_dielectric, float(np.abs(im - ref).mean()), r.jit_info()["last_launch_jit"]))
            print(f"kind {kind} ground emission {emit}: layout {r.scene_layout()}\n    pre {out[0]}\n    jit {out[1]}", flush=True)
