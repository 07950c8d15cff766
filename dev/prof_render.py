"""One short render for ncu: python dev/prof_render.py <scene> <W> <H> <samples/subpixel> [flags] [renders]
`renders` > 1 repeats the render (clearing in between): the run-time compiled kernel is built when a scene is rendered
the second time, so `ncu -k regex:mega --launch-skip 1 -c 1` with renders = 2 captures the kernel bench.py measures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

pkg = load_package()
name, W, H, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
flags = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
renders = int(sys.argv[6]) if len(sys.argv) > 6 else 1
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
    for i in range(renders):
        r.clear()
        r.render(1, 0, S, flags)
        st = r.stats()
        print(f"{name} {W}x{H} S={S} flags={flags:#x} render {i}: {st.last_render_ms:.2f} ms, {W*H*4*S/st.last_render_ms/1e3:.1f} Mpaths/s, "
              f"rays/path {st.rays/(W*H*4*S):.3f}, jit {r.jit_info()}")
