"""Soak test: random scenes through every FP32 render path; checks that every slot got its samples, that nothing is
non-finite, and that the sorted and the in-place megakernel (and scan vs hierarchy) trace the same number of rays.
   python dev/fuzz_scenes.py [n_scenes] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
def make_scene(BASE, it):
    rng = np.random.default_rng(1000003 * BASE + it)
    n = int(rng.choice([1, 2, 3, 5, 8, 9, 12, 16, 17, 30, 64, 65, 100, 700, 3000]))
    style = rng.integers(0, 4)
    s = np.zeros(n, dtype=pkg.SPHERE_DTYPE)
    scale = [1.0, 30.0, 0.05, 1.0][style]            # unit scenes, large scenes, tiny scenes
    s["radius"] = rng.uniform(0.02, 0.4, n) * scale * (1.0 if n < 100 else 0.3)
    s["position"] = (rng.uniform(-1.5, 1.5, (n, 3)) + (0, 0, -1.5)) * scale
    s["color"] = rng.uniform(0.0, 1.0, (n, 3))
    s["emission"][rng.random(n) < 0.15] = rng.uniform(0, 8, 3)
    s["reflection"] = rng.integers(0, 3, n)
    if style == 3 and n >= 6:                      # a closed box of huge spheres around everything, like the reference's
        R = 1e5
        for k, (ax, sg) in enumerate([(0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)]):
            c = np.array([0.0, 0.0, -1.5]); c[ax] += sg * (R + 2.5)
            s[k] = (R, tuple(c), (0, 0, 0), tuple(rng.uniform(0.3, 0.9, 3)), int(rng.integers(0, 2)), 0)
        s[0]["emission"] = (1.5, 1.5, 1.5)
    if rng.random() < 0.3:
        s[n - 1] = (1000.0 * scale, (0, -1000.5 * scale, -1.5 * scale), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)
    W, H, S = int(rng.integers(1, 200)), int(rng.integers(1, 120)), int(rng.integers(1, 12))
    cfg = np.zeros(1, dtype=pkg.CAMERA_CONFIG_DTYPE) if hasattr(pkg, "CAMERA_CONFIG_DTYPE") else None
    _, cfg = pkg.builtin_scene("simple", W, H)
    hostile = rng.random() < 0.5
    if hostile:
        # camera anywhere (often INSIDE a sphere), any aperture; degenerate radii; colours above 1; huge emission
        j = int(rng.integers(0, n))
        cfg["position"][0] = s["position"][j] + rng.uniform(-0.5, 0.5, 3) * s["radius"][j] * rng.choice([0.5, 3.0])
        cfg["direction"][0] = s["position"][int(rng.integers(0, n))] + rng.uniform(-0.1, 0.1, 3) * scale
        if np.allclose(cfg["direction"][0], cfg["position"][0]):
            cfg["direction"][0] += (0.3 * scale, 0.1 * scale, -1.0 * scale)
        cfg["aperture"][0] = rng.choice([0.0, 0.05, 1.0]) * scale
        cfg["focus_distance"][0] = rng.uniform(0.5, 5.0) * scale
        cfg["vertical_fov_radians"][0] = rng.uniform(0.1, 2.5)
        k = rng.integers(0, n, 3)
        s["radius"][k[0]] = rng.choice([0.0, 1e-6 * scale, 5.0 * scale, -0.2 * scale])
        s["color"][k[1]] = rng.uniform(1.0, 1.2, 3)
        s["emission"][k[2]] = 1e6
    cam = pkg.camera_with_config(cfg)
    return s, cfg, cam, W, H, S, n, style, hostile


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    BASE = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ONLY = [int(a) for a in sys.argv[3:]]            # scene numbers to replay (with the oracle's verdict)
    F32 = pkg.PRECISION_FP32
    bad = 0
    oracle = None
    for it in (ONLY if ONLY else range(N)):
        s, cfg, cam, W, H, S, n, style, hostile = make_scene(BASE, it)
        ns = [2, 2, 1, 3, 5][it % 5]
        if ns > 2:
            W, H = max(1, W // 2), max(1, H // 2)
        res, imgs = {}, {}
        with pkg.Renderer(0) as r:
            r.upload_scene(s); r.set_camera(cam); r.set_image(W, H, ns)
            for label, flags, reps in (("sorted", F32 | pkg.VARIANT_MEGAKERNEL_SORTED | pkg.CODEGEN_PRECOMPILED, 1),
                                       ("sorted-jit", F32 | pkg.VARIANT_MEGAKERNEL_SORTED, 2),
                                       ("inplace", F32 | pkg.VARIANT_MEGAKERNEL | pkg.CODEGEN_PRECOMPILED, 1),
                                       ("scan", F32 | pkg.VARIANT_MEGAKERNEL_SORTED | pkg.ACCEL_SCAN | pkg.CODEGEN_PRECOMPILED, 1),
                                       ("wavefront", F32 | pkg.VARIANT_WAVEFRONT, 1)):
                if label == "scan" and n > 700:
                    continue
                for _ in range(reps):
                    r.clear()
                    if label == "inplace" and S > 1:    # progressive: the same samples in two calls
                        r.render(7 + it, 0, S // 2, flags); r.render(7 + it, S // 2, S - S // 2, flags)
                    else:
                        r.render(7 + it, 0, S, flags)
                acc = r.download_accum(); st = r.stats()
                if ONLY or label == "sorted":
                    imgs[label] = r.resolve()
                res[label] = (bool(np.all(acc[:, 3] == S)), bool(np.isfinite(acc).all()), st.rays, float(acc[:, :3].mean()))
        if ONLY:
            if oracle is None:
                from oracle import Oracle
                oracle = Oracle("port")
            ref = oracle.render(s, cam, W, H, S, ns, 7 + it, 0)
            print(f"scene {it}: n={n} style={style} hostile={hostile} {W}x{H}x{S} oracle image mean {ref.mean():.6g}")
            for label, im in imgs.items():
                d = np.abs(im - ref)
                print(f"   {label:10s} rays {res[label][2]:9d}  image mean {im.mean():.6g}  mean|diff| {d.mean():.3e}  pixels off by > 1e-3: {(d.max(axis=2) > 1e-3).mean():.4f}")
            print("   spheres:", s.tolist())
            print("   camera config:", cfg.tolist())
            continue
        vs_oracle = ""
        if W * H * S * ns * ns <= 120000:
            if oracle is None:
                from oracle import Oracle
                oracle = Oracle("port")
            ref = oracle.render(s, cam, W, H, S, ns, 7 + it, 0)
            with pkg.Renderer(0) as r:
                r.upload_scene(s); r.set_camera(cam); r.set_image(W, H, ns)
                r.render(7 + it, 0, S, pkg.PRECISION_FP64)
                img64 = r.resolve()
            d32, d64 = np.abs(imgs["sorted"] - ref), np.abs(img64 - ref)
            off32 = float((d32.max(axis=2) > 0.05).mean())
            off64 = float((d64.max(axis=2) > 1e-6).mean())
            chaotic = res["sorted"][2] > 30 * W * H * S * ns * ns   # ~100 bounces per path (mirror balls with colour > 1): FP32 cannot follow
            if (not chaotic and (d32.mean() > 1e-2 or off32 > 0.06)) or off64 > 0.01:
                vs_oracle = f" ORACLE: fp32 mean|diff| {d32.mean():.2e} pixels off {off32:.3f}; fp64 mean|diff| {d64.mean():.2e} pixels off {off64:.3f}"
        ok = all(v[0] and v[1] for v in res.values()) and not vs_oracle
        rays = [v[2] for v in res.values()]
        spread = (max(rays) - min(rays)) / max(1, max(rays))
        means = [v[3] for v in res.values()]
        mspread = (max(means) - min(means)) / max(1e-9, max(means))
        exact = res["sorted"][2] == res["inplace"][2] == res["wavefront"][2] and ("scan" not in res or res["scan"][2] == res["sorted"][2])
        if not ok or not exact or spread > 5e-3:
            bad += 1
            print(f"scene {it}: n={n} style={style} hostile={hostile} {W}x{H}x{S} PROBLEM ok={ok} exact={exact} spread={spread:.2e}{vs_oracle} {res}", flush=True)
        elif it % 10 == 0:
            print(f"scene {it}: n={n} style={style} hostile={hostile} {W}x{H}x{S} fine (rays {rays[0]}, jit spread {spread:.1e}, mean spread {mspread:.1e})", flush=True)
    print("problems:", bad, "of", N)
