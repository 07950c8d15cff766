"""Are the run-time compiled kernels arithmetically identical to the precompiled ones?  Compare counters and slots at size.
   python dev/jit_identity.py <scene> W H S"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
name, W, H, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
    res = {}
    for vname, v in (("in-place", pkg.VARIANT_MEGAKERNEL), ("sorted", pkg.VARIANT_MEGAKERNEL_SORTED)):
        for cname, c in (("precompiled", pkg.CODEGEN_PRECOMPILED), ("jit", pkg.CODEGEN_AUTO)):
            f = pkg.PRECISION_FP32 | v | c
            r.clear(); r.render(9, 0, S, f); r.clear(); r.render(9, 0, S, f)
            st = r.stats()
            res[(vname, cname)] = (st.rays, st.hits_diffuse, st.hits_specular, st.hits_dielectric, r.download_accum(), r.jit_info()["last_launch_jit"])
    base = res[("in-place", "precompiled")]
    for k, v in res.items():
        d = np.abs(v[4][:, :3] - base[4][:, :3]) / np.maximum(np.abs(base[4][:, :3]), 1e-3)
        print(k, "jit" if v[5] else "pre", "rays", v[0], "d/s/g", v[1:4], "rays diff vs in-place precompiled", v[0] - base[0],
              "slots > 1e-4 rel:", int((d.max(axis=1) > 1e-4).sum()))
