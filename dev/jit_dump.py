"""Dump the run-time generated translation units (PTB_JIT_DUMP) of the built-in scenes: python dev/jit_dump.py <outdir>
The .cu files hold the packed coefficients as literals; dev/jit_offline.py recompiles them against the working tree's
headers on a machine without a GPU and counts the instructions of the bounce loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.abspath(sys.argv[1]); os.makedirs(out, exist_ok=True)
os.environ["PTB_JIT_DUMP"] = out
from __graft_entry__ import load_package
pkg = load_package()
for name in ("box_mirror", "box", "dof_glass", "simple"):
    sph, cfg = pkg.builtin_scene(name, 64, 48)
    with pkg.Renderer(0) as r:
        r.upload_scene(sph); r.set_camera(pkg.camera_with_config(cfg)); r.set_image(64, 48, 2)
        for _ in range(3):
            r.render(1, 0, 2, pkg.VARIANT_MEGAKERNEL_SORTED)
        info = r.jit_info()
        pass
    print(name, info)
    for f in sorted(os.listdir(out)):
        if f.startswith("jit_"):  # what this scene's context wrote: jit_<n>_<kind>.{cu,name,cubin}
            os.rename(os.path.join(out, f), os.path.join(out, name + "." + f.rsplit(".", 1)[1]))
sph, cam8 = pkg.builtin_smallpt_scene()
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_smallpt_camera(cam8); r.set_image(64, 48, 2)
    for _ in range(3):
        r.render(1, 0, 2, pkg.VARIANT_MEGAKERNEL | pkg.INTEGRATOR_SMALLPT)
    print("smallpt", r.jit_info())
for f in sorted(os.listdir(out)):
    if f.startswith("jit_"):
        os.rename(os.path.join(out, f), os.path.join(out, "smallpt." + f.rsplit(".", 1)[1]))
