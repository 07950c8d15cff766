"""Soak test of the context's state machine: a long-lived context goes through random sequences of upload_scene / set_camera /
set_image / clear / render (any variant, any flags) and must always render what a FRESH context renders for the same inputs.
   python dev/fuzz_api_sequences.py [n_steps] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fuzz_scenes import make_scene, pkg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
F32 = pkg.PRECISION_FP32
variants = [("sorted", pkg.VARIANT_MEGAKERNEL_SORTED), ("inplace", pkg.VARIANT_MEGAKERNEL), ("wavefront", pkg.VARIANT_WAVEFRONT)]
bad = 0
live = pkg.Renderer(0)
state = {}


def pick_scene():
    if rng.random() < 0.4:
        name = str(rng.choice(["simple", "box", "box_mirror", "dof_glass"]))
        W, H = int(rng.integers(8, 120)), int(rng.integers(8, 90))
        s, cfg = pkg.builtin_scene(name, W, H)
        return s, pkg.camera_with_config(cfg), W, H, name
    s, cfg, cam, W, H, S, n, style, hostile = make_scene(int(rng.integers(0, 50)), int(rng.integers(0, 250)))
    return s, cam, W, H, f"random n={n} style={style} hostile={hostile}"


s, cam, W, H, what = pick_scene()
live.upload_scene(s); live.set_camera(cam); live.set_image(W, H, 2)
state.update(s=s, cam=cam, W=W, H=H, ns=2, first=0, what=what)
for step in range(N):
    op = rng.choice(["scene", "camera", "image", "clear", "render", "render", "render"])
    if op == "scene":
        s, cam, W, H, what = pick_scene()
        live.upload_scene(s)
        state.update(s=s, first=0, what=what)
        live.clear()
    elif op == "camera":
        _, cam, _, _, _ = pick_scene()
        live.set_camera(cam)
        state.update(cam=cam, first=0)
        live.clear()
    elif op == "image":
        W, H, ns = int(rng.integers(1, 150)), int(rng.integers(1, 100)), int(rng.choice([1, 2, 2, 3]))
        live.set_image(W, H, ns)
        state.update(W=W, H=H, ns=ns, first=0)
    elif op == "clear":
        live.clear()
        state.update(first=0)
    else:
        vname, vflag = variants[int(rng.integers(0, 3))]
        flags = F32 | vflag
        if rng.random() < 0.3:
            flags |= pkg.CODEGEN_PRECOMPILED
        if rng.random() < 0.2:
            flags |= pkg.ACCEL_SCAN
        S = int(rng.integers(1, 6))
        seed = int(rng.integers(0, 1000))
        # a fresh sequence of samples every time: clear, then render [0, S)
        live.clear()
        live.render(seed, 0, S, flags)
        acc_l, rays_l = live.download_accum(), live.stats().rays
        with pkg.Renderer(0) as fresh:
            fresh.upload_scene(state["s"]); fresh.set_camera(state["cam"]); fresh.set_image(state["W"], state["H"], state["ns"])
            fresh.render(seed, 0, S, F32 | pkg.VARIANT_MEGAKERNEL | pkg.CODEGEN_PRECOMPILED | (flags & pkg.ACCEL_SCAN))
            acc_f, rays_f = fresh.download_accum(), fresh.stats().rays
        jit = live.jit_info()["last_launch_jit"]
        rays_ok = rays_l == rays_f if (jit == 0) else abs(rays_l - rays_f) <= 5e-3 * max(rays_f, 1)
        close = np.isclose(acc_l[:, :3], acc_f[:, :3], rtol=1e-4, atol=1e-4).all(axis=1).mean()
        ok = rays_ok and np.all(acc_l[:, 3] == S) and np.isfinite(acc_l).all() and (close == 1.0 if jit == 0 else close > 0.97)
        if not ok:
            bad += 1
        if not ok or step % 20 == 0:
            print(f"step {step}: render {vname} flags {flags:#x} S={S} {state['W']}x{state['H']}x{state['ns']} [{state['what']}] jit={jit} "
                  f"rays {rays_l} vs {rays_f}  slots close {close:.4f} {'fine' if ok else 'PROBLEM'}", flush=True)
live.close() if hasattr(live, "close") else None
print("problems:", bad, "of", N, "steps")
