"""Which material should the sorted megakernel scatter in place?  Times the run-time-compiled kernel with
PTB_INLINE_MATERIAL = -1 (none: every hit is parked), 0 (diffuse), 1 (specular) on the built-in scenes.
   python dev/inline_policy.py [W H S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
W, H, S = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1920, 1080, 64)
flags = pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL_SORTED
for name in ("simple", "dof_glass", "box", "box_mirror"):
    sph, cfg = pkg.builtin_scene(name, W, H)
    cam = pkg.camera_with_config(cfg)
    row = []
    for inl in (-1, 0, 1):
        os.environ["PTB_INLINE_MATERIAL"] = str(inl)
        with pkg.Renderer(0) as r:
            r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
            r.render(1, 0, 2, flags); r.render(1, 2, 2, flags)  # second sight: compiled
            r.clear()
            best = 1e30
            for _ in range(3):
                r.clear(); r.render(1, 0, S, flags)
                best = min(best, r.stats().last_render_ms)
            st = r.stats()
            acc = r.download_accum()
            ok = bool((acc[:, 3] == S).all())
            row.append((inl, best, r.jit_info()["last_launch_jit"], ok, st.rays))
    paths = W * H * 4 * S
    print(name, " ".join(f"[inline {i:2d}: {paths / t / 1e3:8.1f} Mpaths/s jit={j} slots_ok={ok} rays={rays}]" for i, t, j, ok, rays in row), flush=True)
