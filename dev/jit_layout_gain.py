"""What run-time compilation buys a scene WITHOUT a precompiled specialisation (box_mirror + 2 balls): unrolled scan
with immediates against the run-time-count scan.  python dev/jit_layout_gain.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
W, H, S = 1920, 1080, 64
sph, cfg = pkg.builtin_scene("box_mirror", W, H)
cam = pkg.camera_with_config(cfg)
extra = sph[6:8].copy(); extra["position"][:, 1] += 0.35; extra["position"][:, 2] -= 0.15
scene = np.concatenate([sph, extra])
with pkg.Renderer(0) as r:
    r.upload_scene(scene); r.set_camera(cam); r.set_image(W, H, 2)
    for name, f in (("precompiled run-time-count scan", pkg.CODEGEN_PRECOMPILED), ("run-time compiled, unrolled", pkg.CODEGEN_AUTO)):
        flags = pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL_SORTED | f
        r.render(1, 0, S, flags); r.clear(); r.render(1, 0, S, flags)
        st = r.stats()
        print(f"{name:36s} {st.last_render_ms:8.2f} ms  {W*H*4*S/st.last_render_ms/1e3:8.1f} Mpaths/s  jit={r.jit_info()['last_launch_jit']}")
