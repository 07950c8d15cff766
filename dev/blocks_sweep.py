"""Occupancy sweep of the sorted megakernel: PTB_BLOCKS_PER_SM = 1..6 on one scene.
   python dev/blocks_sweep.py <scene> <W> <H> <S> [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
name, W, H, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
flags = int(sys.argv[5], 0) if len(sys.argv) > 5 else (pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL_SORTED | pkg.CODEGEN_PRECOMPILED)
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
for b in (6, 5, 4, 3, 2):
    os.environ["PTB_BLOCKS_PER_SM"] = str(b)
    with pkg.Renderer(0) as r:
        r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
        r.render(1, 0, 2, flags)
        best = 1e30
        for _ in range(3):
            r.clear(); r.render(1, 0, S, flags)
            best = min(best, r.stats().last_render_ms)
    print(f"{name} blocks/SM {b}: {best:9.3f} ms  {W * H * 4 * S / best / 1e3:9.1f} Mpaths/s", flush=True)
