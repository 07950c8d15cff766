"""Soak test of the sandbox (smallpt) integrator: random scenes in the sandbox's units, FP32 in-place kernel (precompiled and
run-time build) and FP64 kernel against the oracle's restatement of sandbox/main.cpp.
   python dev/fuzz_smallpt.py [n_scenes] [base]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from oracle import Oracle
oracle = Oracle("port")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
BASE = int(sys.argv[2]) if len(sys.argv) > 2 else 0
base_sph, base_cam = pkg.builtin_smallpt_scene()
F = pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL | pkg.INTEGRATOR_SMALLPT
bad = 0
for it in range(N):
    rng = np.random.default_rng(7000003 * BASE + it)
    style = int(rng.integers(0, 3))
    if style == 0:      # the sandbox box with other balls in it
        s = base_sph.copy()
        k = int(rng.integers(0, 10))
        extra = np.zeros(k, dtype=pkg.SPHERE_DTYPE)
        extra["radius"] = rng.uniform(3, 18, k)
        extra["position"] = rng.uniform((10, 5, 10), (90, 75, 150), (k, 3))
        extra["color"] = rng.uniform(0.1, 0.999, (k, 3))
        extra["reflection"] = rng.integers(0, 3, k)
        extra["emission"][rng.random(k) < 0.2] = (6, 6, 6)
        s = np.concatenate([s, extra])
        if rng.random() < 0.5:
            s["reflection"][:6] = rng.integers(0, 2, 6)       # mirror walls
    elif style == 1:    # free-floating balls, black background
        k = int(rng.integers(1, 40))
        s = np.zeros(k, dtype=pkg.SPHERE_DTYPE)
        s["radius"] = rng.uniform(2, 25, k)
        s["position"] = rng.uniform((0, 0, -50), (100, 80, 160), (k, 3))
        s["color"] = rng.uniform(0.1, 0.999, (k, 3))
        s["reflection"] = rng.integers(0, 3, k)
        s["emission"][rng.random(k) < 0.3] = rng.uniform(1, 12, 3)
    else:               # many small balls (hierarchy)
        k = int(rng.choice([70, 300, 1500]))
        s = np.zeros(k + 6, dtype=pkg.SPHERE_DTYPE)
        s[:6] = base_sph[:6]
        s["radius"][6:] = rng.uniform(0.5, 3.0, k)
        s["position"][6:] = rng.uniform((3, 2, 3), (97, 79, 165), (k, 3))
        s["color"][6:] = rng.uniform(0.1, 0.999, (k, 3))
        s["reflection"][6:] = rng.integers(0, 3, k)
        s["emission"][6:][rng.random(k) < 0.1] = (8, 8, 8)
    cam8 = base_cam.copy()
    if rng.random() < 0.5:
        cam8[:3] += rng.uniform(-15, 15, 3)
        cam8[3:6] += rng.uniform(-0.2, 0.2, 3)
    W, H, S = int(rng.integers(4, 90)), int(rng.integers(4, 70)), int(rng.integers(1, 5))
    ref = oracle.sb_render(s, cam8, W, H, S, 1, 11 + it, 0)
    out = {}
    with pkg.Renderer(0) as r:
        r.upload_scene(s); r.set_smallpt_camera(cam8); r.set_image(W, H, 2)
        for label, flags, reps in (("f32", F | pkg.CODEGEN_PRECOMPILED, 1), ("f32-jit", F, 2), ("f32-scan", F | pkg.ACCEL_SCAN | pkg.CODEGEN_PRECOMPILED, 1),
                                   ("f64", pkg.PRECISION_FP64 | pkg.INTEGRATOR_SMALLPT, 1)):
            for _ in range(reps):
                r.clear(); r.render(11 + it, 0, S, flags)
            acc = r.download_accum()
            out[label] = (r.resolve(), r.stats().rays, bool(np.isfinite(acc).all()), bool(np.all(acc[:, 3] == S)))
    d64 = np.abs(out["f64"][0] - ref)
    d32 = np.abs(out["f32"][0] - ref)
    off32 = float((d32.max(axis=2) > 0.05).mean())
    rays = [out[k][1] for k in ("f32", "f32-jit", "f32-scan")]
    spread = (max(rays) - min(rays)) / max(1, max(rays))
    ok = all(v[2] for v in out.values()) and all(out[k][3] for k in ('f32', 'f32-jit', 'f32-scan'))
    problem = (not ok) or (d64.max(axis=2) > 1e-6).mean() > 0.01 or d32.mean() > 1.5e-2 or off32 > 0.08 or rays[0] != rays[2] or spread > 5e-3
    if problem:
        bad += 1
    if problem or it % 10 == 0:
        print(f"scene {it}: style {style} n={len(s)} {W}x{H}x{S} {'PROBLEM' if problem else 'fine'} finite/slots {ok}  fp64 pixels off {(d64.max(axis=2) > 1e-6).mean():.4f}"
              f"  fp32 mean|diff| {d32.mean():.2e} off {off32:.3f}  rays {rays} f64 {out['f64'][1]}", flush=True)
print("problems:", bad, "of", N)
