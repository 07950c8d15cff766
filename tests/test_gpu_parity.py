"""Parity of the CUDA path against the oracle, through the C ABI (ctypes over libptb200.so).

Bars (BASELINE.json north_star):
  (a) deterministic mode (PTB_PRECISION_FP64): primary-ray hit sphere indices BIT-EXACT, per-sample
      radiance within 1e-4 relative;
  (b) throughput mode (FP32 megakernel): images statistically matching at matched spp.
"""
import numpy as np
import pytest

from conftest import classify_outliers, make_renderer

pytestmark = pytest.mark.gpu

SCENES = ("simple", "box", "box_mirror", "dof_glass")
REL_TOL = 1e-4  # north_star: "per-sample radiance matches within 1e-4 relative"


def rel_err(a, b):
    return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-12)


def probe_inputs(rng, W, H, n, nsub=2, smax=1 << 24):
    return (rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(0, nsub, n), rng.integers(0, nsub, n),
            rng.integers(0, smax, n))


# ---- (a) deterministic mode ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["simple", "box", "box_mirror"])
def test_fp64_samples_against_reference_golden(gpu, golden, golden_scene, name):
    """The committed vectors come from the reference's own intersect()/radiance() (tests/golden/make_golden.py)."""
    z = golden(f"samples_{name}.npz")
    W, H = int(z["width"]), int(z["height"])
    sph, _, cam = golden_scene(name, W, H)
    from oracle import Oracle
    with make_renderer(gpu, sph, cam, W, H, int(z["nsub"])) as r:
        hit, rad, ray, draws = r.trace_samples(int(z["seed"]), z["x"], z["y"], z["sx"], z["sy"], z["sample"],
                                               gpu.PRECISION_FP64)
        # every sample outside 1e-4 is accounted for: same spheres as the reference's path through a diffuse / glass bounce
        n_bad, n_div = classify_outliers(gpu, Oracle("port"), r, sph, cam, W, H, int(z["nsub"]), int(z["seed"]), z["x"], z["y"],
                                         z["sx"], z["sy"], z["sample"], rad, z["radiance"], REL_TOL)
    print(f"{name}: {n_bad} of {len(rad)} samples outside {REL_TOL} (all after a diffuse/glass bounce; {n_div} change spheres later)")
    assert np.array_equal(hit, z["hit"]), "primary-hit indices must be bit-exact"
    assert np.array_equal(ray, z["ray"]), "camera rays use only + - * / sqrt: bit-exact"
    assert n_bad <= 1e-3 * len(rad)
    assert (draws == z["draws"]).mean() >= 0.999


@pytest.mark.parametrize("name", SCENES)
def test_fp64_samples_against_oracle(gpu, oracle_port, name):
    rng = np.random.default_rng(7)
    W, H, n = 333, 187, 50000  # odd sizes on purpose
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, n)
    ohit, orad, oray, odraws = oracle_port.samples(sph, cam, W, H, 2, 99, xs, ys, sx, sy, ss)
    with make_renderer(gpu, sph, cam, W, H) as r:
        hit, rad, ray, draws = r.trace_samples(99, xs, ys, sx, sy, ss, gpu.PRECISION_FP64)
        n_bad, n_div = classify_outliers(gpu, oracle_port, r, sph, cam, W, H, 2, 99, xs, ys, sx, sy, ss, rad, orad, REL_TOL)
        # the same statement at tolerance ZERO: a sample that is not bit-identical to the oracle's has met a diffuse or
        # glass surface before the first difference -- mirror-only paths (+ - * / sqrt) are exact to the last bit
        n_inexact, _ = classify_outliers(gpu, oracle_port, r, sph, cam, W, H, 2, 99, xs, ys, sx, sy, ss, rad, orad, 0.0)
    print(f"{name}: {n_bad} of {n} samples outside {REL_TOL} (all after a diffuse/glass bounce; {n_div} change spheres later); "
          f"{n_inexact} not bit-identical, every one of them after a diffuse/glass bounce")
    assert np.array_equal(hit, ohit)
    assert np.array_equal(ray, oray)
    assert n_bad <= 1e-3 * n
    # mirror-only paths touch nothing but + - * / sqrt: most samples are bit-identical
    assert (rad == orad).all(axis=1).mean() > 0.9
    assert (draws == odraws).mean() >= 0.999


@pytest.mark.parametrize("name", SCENES)
def test_fp64_image_against_oracle(gpu, oracle_port, name):
    W, H, S = 96, 54, 6
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(sph, cam, W, H, S, 2, 5, 0)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(5, 0, S, gpu.PRECISION_FP64)
        img = r.resolve()
        st = r.stats()
    assert st.paths == W * H * 4 * S
    # a handful of chaotic paths may differ (sin/cos/pow ulps): everything else agrees to rounding
    close = np.abs(img - ref) <= 1e-9
    assert close.mean() >= 0.998
    assert np.abs(img - ref).mean() < 1e-4


def test_fp64_image_against_reference_golden(gpu, golden):
    for name in ("simple", "box", "box_mirror"):
        z = golden(f"image_{name}.npz")
        w, h, S = int(z["width"]), int(z["height"]), int(z["samps"])
        sph, cfg = gpu.builtin_scene(name, w, h)
        cam = gpu.camera_with_config(cfg)
        with make_renderer(gpu, sph, cam, w, h, int(z["nsub"])) as r:
            r.render(int(z["seed"]), int(z["first_sample"]), S, gpu.PRECISION_FP64)
            img = r.resolve()
        assert (np.abs(img - z["image"]) <= 1e-9).mean() >= 0.998


def test_nsub_other_than_two(gpu, oracle_port):
    for nsub in (1, 3):
        W, H, S = 40, 30, 3
        sph, cfg = gpu.builtin_scene("box", W, H)
        cam = gpu.camera_with_config(cfg)
        ref = oracle_port.render(sph, cam, W, H, S, nsub, 8, 0)
        with make_renderer(gpu, sph, cam, W, H, nsub) as r:
            r.render(8, 0, S, gpu.PRECISION_FP64)
            img64 = r.resolve()
            r.clear()
            r.render(8, 0, S, gpu.PRECISION_FP32)
            img32 = r.resolve()
            r.clear()
            r.render(8, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
            img32s = r.resolve()
            acc = r.download_accum()
        assert (np.abs(img64 - ref) <= 1e-9).mean() >= 0.995
        assert np.abs(img32 - ref).mean() < 5e-3
        assert np.all(acc[:, 3] == S) and acc.shape[0] == W * H * nsub * nsub
        assert np.abs(img32s - img32).max() < 1e-5  # the two megakernels trace the same paths


# ---- (b) throughput mode -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_fp32_samples_against_oracle(gpu, oracle_port, name):
    """Same uniforms, FP32 arithmetic: hits agree, radiance agrees except on chaotic (deep specular) paths,
    and the MEAN agrees within Monte-Carlo error (no bias)."""
    rng = np.random.default_rng(11)
    W, H, n = 320, 180, 200000
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, n)
    ohit, orad, oray, _ = oracle_port.samples(sph, cam, W, H, 2, 3, xs, ys, sx, sy, ss)
    with make_renderer(gpu, sph, cam, W, H) as r:
        hit, rad, ray, _ = r.trace_samples(3, xs, ys, sx, sy, ss, gpu.PRECISION_FP32)
    assert (hit == ohit).mean() >= 0.9999
    assert np.abs(ray - oray).max() < 1e-5
    rel = rel_err(rad, orad)
    assert (rel <= 1e-3).mean() >= 0.985
    se = np.sqrt((rad.var(axis=0) + orad.var(axis=0)) / n)
    z = (rad.mean(axis=0) - orad.mean(axis=0)) / se
    assert np.abs(z).max() < 4.0, f"FP32 mean radiance biased: z = {z}"


@pytest.mark.parametrize("variant", ["VARIANT_MEGAKERNEL", "VARIANT_MEGAKERNEL_SORTED"])
@pytest.mark.parametrize("name", SCENES)
def test_fp32_image_against_oracle_same_seed(gpu, oracle_port, name, variant):
    W, H, S = 160, 90, 8
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(sph, cam, W, H, S, 2, 17, 0)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(17, 0, S, gpu.PRECISION_FP32 | getattr(gpu, variant))
        img = r.resolve()
        st = r.stats()
        acc = r.download_accum()
    assert np.all(acc[:, 3] == S), "every sub-pixel must receive exactly S samples"
    assert np.abs(img - ref).mean() < 2e-3
    # pixels holding a chaotic (deep specular) path differ; 32 samples per pixel here, ~0.4 % of samples chaotic
    assert (np.abs(img - ref) < 1e-4).mean() > (0.85 if name == "box_mirror" else 0.95)
    assert abs(st.rays / st.paths - {"simple": 2.09, "box": 12.33, "box_mirror": 12.33, "dof_glass": 2.1}[name]) < 0.25


@pytest.mark.parametrize("name", ["box", "box_mirror", "dof_glass"])
def test_fp32_image_rmse_within_monte_carlo_noise(gpu, oracle_port, name):
    """Independent seeds: RMSE(GPU, reference mean) within the Monte-Carlo bound estimated from K reference renders
    ('Image RMSE')."""
    W, H, S, K = 96, 72, 8, 6
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    refs = np.stack([oracle_port.render(sph, cam, W, H, S, 2, 1000 + k, 0) for k in range(K)])
    mean_ref, var_px = refs.mean(axis=0), refs.var(axis=0, ddof=1)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(4242, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        img = r.resolve()
    for c in range(3):
        rmse = np.sqrt(np.mean((img[..., c] - mean_ref[..., c]) ** 2))
        bound = 1.25 * np.sqrt(np.mean(var_px[..., c]) * (1.0 / K + 1.0))
        assert rmse <= bound, f"channel {c}: RMSE {rmse:.4g} > noise bound {bound:.4g}"


# ---- the wavefront / material-sorted variant traces the same paths --------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_wavefront_variant_matches_megakernel(gpu, oracle_port, name):
    W, H, S = 192, 108, 12
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(sph, cam, W, H, S, 2, 23, 0)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(23, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL)
        mega = r.download_accum()
        st_m = r.stats()
        r.clear()
        r.render(23, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_WAVEFRONT)
        wave = r.download_accum()
        st_w = r.stats()
        img = r.resolve()
    assert np.all(mega[:, 3] == S) and np.all(wave[:, 3] == S)
    assert st_w.paths == st_m.paths and abs(st_w.rays - st_m.rays) <= 1e-3 * st_m.rays
    close = np.isclose(mega[:, :3], wave[:, :3], rtol=1e-4, atol=1e-4).all(axis=1)
    assert close.mean() > 0.995  # same arithmetic, same uniforms: only summation order and a few chaotic paths differ
    assert np.abs(img - ref).mean() < 2e-3


@pytest.mark.parametrize("name", SCENES + ("spheres10k",))
def test_sorted_megakernel_matches_megakernel(gpu, name):
    """PTB_VARIANT_MEGAKERNEL_SORTED (shared-memory material sorting) traces the very same paths as the in-place
    megakernel: same arithmetic, same uniforms; only the order of the float additions into a slot differs."""
    W, H, S = (192, 108, 12) if name != "spheres10k" else (64, 36, 3)
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(23, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL)
        mega = r.download_accum()
        st_m = r.stats()
        r.clear()
        r.render(23, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        srt = r.download_accum()
        st_s = r.stats()
    assert np.all(mega[:, 3] == S) and np.all(srt[:, 3] == S)
    assert np.isfinite(mega).all() and np.isfinite(srt).all()
    assert st_s.paths == st_m.paths
    assert (st_s.rays, st_s.hits_diffuse, st_s.hits_specular, st_s.hits_dielectric) == \
        (st_m.rays, st_m.hits_diffuse, st_m.hits_specular, st_m.hits_dielectric)
    assert np.isclose(mega[:, :3], srt[:, :3], rtol=1e-5, atol=1e-5).all(axis=1).mean() > 0.9999


def test_sorted_megakernel_small_and_progressive(gpu):
    W, H = 9, 7  # a single warp's worth of slots: the drain phase of the rings is most of the run
    sph, cfg = gpu.builtin_scene("box", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(3, 0, 2, gpu.VARIANT_MEGAKERNEL_SORTED)
        r.render(3, 2, 3, gpu.VARIANT_MEGAKERNEL_SORTED)
        a = r.download_accum()
        r.clear()
        r.render(3, 0, 5, gpu.VARIANT_MEGAKERNEL)
        b = r.download_accum()
    assert np.all(a[:, 3] == 5) and np.all(b[:, 3] == 5)
    assert np.isclose(a[:, :3], b[:, :3], rtol=1e-5, atol=1e-5).all()


def test_wavefront_small_and_progressive(gpu):
    W, H = 9, 7  # fewer items than one block of the pool
    sph, cfg = gpu.builtin_scene("box", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(3, 0, 2, gpu.VARIANT_WAVEFRONT)
        r.render(3, 2, 3, gpu.VARIANT_WAVEFRONT)
        a = r.download_accum()
        r.clear()
        r.render(3, 0, 5, gpu.VARIANT_MEGAKERNEL)
        b = r.download_accum()
    assert np.all(a[:, 3] == 5) and np.all(b[:, 3] == 5)
    assert np.isclose(a[:, :3], b[:, :3], rtol=1e-4, atol=1e-4).all(axis=1).mean() > 0.97


# ---- BASELINE config 5: 10 001 spheres, closest hit through the bounding-volume hierarchy ----------------------------------
def test_spheres10k_fp64_against_oracle(gpu, oracle_port):
    """Deterministic mode on the 10 001-sphere scene (linear scan in FP64, like the reference)."""
    rng = np.random.default_rng(3)
    W, H, n = 160, 90, 3000
    sph, cfg = gpu.builtin_scene("spheres10k", W, H)
    cam = gpu.camera_with_config(cfg)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, n)
    ohit, orad, oray, _ = oracle_port.samples(sph, cam, W, H, 2, 41, xs, ys, sx, sy, ss)
    with make_renderer(gpu, sph, cam, W, H) as r:
        hit, rad, ray, _ = r.trace_samples(41, xs, ys, sx, sy, ss, gpu.PRECISION_FP64)
    assert np.array_equal(hit, ohit)
    assert np.array_equal(ray, oray)
    assert (rel_err(rad, orad) <= REL_TOL).mean() >= 0.999


def test_spheres10k_hierarchy_gives_the_scan_s_hits(gpu, oracle_port):
    """PTB_ACCEL_AUTO (hierarchy) against PTB_ACCEL_SCAN (every sphere, main.cpp:30-42) in FP32: the same sphere test
    decides every hit, so per-sample primary hits and radiance are IDENTICAL; and both agree with the FP64 oracle."""
    rng = np.random.default_rng(5)
    W, H, n = 160, 90, 20000
    sph, cfg = gpu.builtin_scene("spheres10k", W, H)
    cam = gpu.camera_with_config(cfg)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, n)
    ohit, orad, _, _ = oracle_port.samples(sph, cam, W, H, 2, 43, xs[:4000], ys[:4000], sx[:4000], sy[:4000], ss[:4000])
    with make_renderer(gpu, sph, cam, W, H) as r:
        hit_a, rad_a, ray_a, _ = r.trace_samples(43, xs, ys, sx, sy, ss, gpu.PRECISION_FP32 | gpu.ACCEL_AUTO)
        hit_s, rad_s, ray_s, _ = r.trace_samples(43, xs, ys, sx, sy, ss, gpu.PRECISION_FP32 | gpu.ACCEL_SCAN)
    assert np.array_equal(hit_a, hit_s)
    assert np.array_equal(rad_a, rad_s), "the hierarchy must not change a single bounce"
    assert np.isfinite(rad_a).all()
    assert (hit_a[:4000] == ohit).mean() >= 0.999
    rel = rel_err(rad_a[:4000], orad)
    assert (rel <= 1e-3).mean() >= 0.97
    se = np.sqrt((rad_a[:4000].var(axis=0) + orad.var(axis=0)) / 4000)
    assert np.abs((rad_a[:4000].mean(axis=0) - orad.mean(axis=0)) / se).max() < 4.0


@pytest.mark.parametrize("variant", ["VARIANT_MEGAKERNEL", "VARIANT_MEGAKERNEL_SORTED"])
def test_spheres10k_render_hierarchy_against_scan(gpu, variant):
    W, H, S = 96, 54, 4
    sph, cfg = gpu.builtin_scene("spheres10k", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(9, 0, S, gpu.PRECISION_FP32 | getattr(gpu, variant) | gpu.ACCEL_AUTO)
        a, st_a = r.download_accum(), r.stats()
        r.clear()
        r.render(9, 0, S, gpu.PRECISION_FP32 | getattr(gpu, variant) | gpu.ACCEL_SCAN)
        b, st_b = r.download_accum(), r.stats()
    assert np.all(a[:, 3] == S) and np.all(b[:, 3] == S) and np.isfinite(a).all()
    assert (st_a.rays, st_a.hits_diffuse, st_a.hits_specular, st_a.hits_dielectric) == \
        (st_b.rays, st_b.hits_diffuse, st_b.hits_specular, st_b.hits_dielectric)
    assert np.isclose(a[:, :3], b[:, :3], rtol=1e-5, atol=1e-5).all()  # same paths; only the order of the additions differs
    assert st_a.last_render_ms < st_b.last_render_ms  # and that is the point of it


# ---- BASELINE config 3 at its full image size: properties that do not depend on the size ------------------------------------
def test_full_size_box_mirror_properties(gpu, oracle_port):
    """1920x1080 box_mirror (the north-star workload) at 64 spp, through the product path.  Checked: every one of the
    8.3 M sub-pixel slots received exactly its samples; the path statistics are the reference's (SURVEY.md 8c:
    12.33 rays, 0.53 diffuse / 9.81 mirror / 0.99 glass hits per path); rendering the samples in two passes gives the
    image of one pass (what progressive rendering and the multi-GPU sample split rely on); and the 8x8 box-filtered
    image is the oracle's 240x135 image of the same scene within Monte-Carlo noise."""
    W, H, S = 1920, 1080, 16
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(77, 0, S, flags)
        one = r.download_accum()
        st = r.stats()
        img = r.resolve()
        r.clear()
        r.render(77, 0, 5, flags)
        r.render(77, 5, S - 5, flags)
        two = r.download_accum()
    assert one.shape == (W * H * 4, 4) and np.all(one[:, 3] == S) and np.all(two[:, 3] == S)
    assert np.isfinite(one).all()
    paths = W * H * 4 * S
    assert st.paths == paths
    assert abs(st.rays / paths - 12.33) < 0.15
    assert abs(st.hits_diffuse / paths - 0.53) < 0.03
    assert abs(st.hits_specular / paths - 9.81) < 0.15
    assert abs(st.hits_dielectric / paths - 0.99) < 0.05
    assert np.isclose(one[:, :3], two[:, :3], rtol=1e-5, atol=1e-5).all()
    # size-independent image check: an oracle pixel at 240x135 covers 8x8 pixels of this render
    sph_s, cfg_s = gpu.builtin_scene("box_mirror", W // 8, H // 8)
    ref = oracle_port.render(sph_s, gpu.camera_with_config(cfg_s), W // 8, H // 8, 16, 2, 5, 0)
    small = img.reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
    dark = ref.max(axis=2) < 0.8  # the per-sub-pixel clamp (main.cpp:195) is not linear: compare below it
    assert dark.mean() > 0.5
    assert np.abs(small - ref)[dark].mean() < 0.02
    assert abs(small[dark].mean() - ref[dark].mean()) < 0.005


# ---- run-time specialised code (PTB_CODEGEN_AUTO) ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_runtime_compiled_kernel_is_the_precompiled_kernel(gpu, name):
    """The sorted megakernel compiled at run time for THIS scene (sphere coefficients as immediates) traces the very
    paths of the precompiled constant-bank kernel; it is compiled once per scene and reused."""
    W, H, S = 192, 108, 10
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, sph, cam, W, H) as r:
        info0 = r.jit_info()
        if not info0["available"]:
            pytest.skip("libnvrtc / libcuda not available: " + info0["last_error"])
        r.render(31, 0, S, flags | gpu.CODEGEN_PRECOMPILED)
        assert r.jit_info()["last_launch_jit"] == 0
        pre, st_p = r.download_accum(), r.stats()
        r.clear()
        r.render(31, 0, S, flags | gpu.CODEGEN_AUTO)
        # a small job does not pay for a compilation the first time a scene shows up ...
        assert r.jit_info()["last_launch_jit"] == 0 and r.jit_info()["compiled"] == info0["compiled"]
        r.clear()
        r.render(31, 0, S, flags | gpu.CODEGEN_AUTO)  # ... the second time it does
        info = r.jit_info()
        assert info["failures"] == 0, info["last_error"]
        assert info["last_launch_jit"] == 1 and info["compiled"] == info0["compiled"] + 1
        jit, st_j = r.download_accum(), r.stats()
        r.upload_scene(sph)  # same numbers again: the cache must answer
        r.set_camera(cam)
        r.render(31, S, S, flags)
        assert r.jit_info()["compiled"] == info["compiled"]
    assert np.all(pre[:, 3] == S) and np.all(jit[:, 3] == S)
    assert (st_j.rays, st_j.hits_diffuse, st_j.hits_specular, st_j.hits_dielectric) == \
        (st_p.rays, st_p.hits_diffuse, st_p.hits_specular, st_p.hits_dielectric)
    assert np.isclose(pre[:, :3], jit[:, :3], rtol=1e-5, atol=1e-5).all()


def test_runtime_compilation_specialises_layouts_the_library_does_not_ship(gpu, oracle_port):
    """box_mirror plus two more balls has no precompiled unrolled kernel (it would take the run-time-count scan);
    compiled at run time it gets one.  Both agree with the FP64 oracle and with each other."""
    W, H, S = 160, 90, 8
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    extra = sph[6:8].copy()  # a second mirror ball and a second glass ball, higher up
    extra["position"][:, 1] += 0.35
    extra["position"][:, 2] -= 0.15
    scene = np.concatenate([sph, extra])
    ref = oracle_port.render(scene, cam, W, H, S, 2, 13, 0)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, scene, cam, W, H) as r:
        if not r.jit_info()["available"]:
            pytest.skip("run-time compilation not available")
        assert r.scene_layout()["specialised"] == 0
        r.render(13, 0, S, flags | gpu.CODEGEN_PRECOMPILED)
        pre, img_pre, st_p = r.download_accum(), r.resolve(), r.stats()
        r.clear()
        r.render(13, 0, S, flags)
        r.clear()
        r.render(13, 0, S, flags)  # second sight of the scene: compiled
        info = r.jit_info()
        assert info["failures"] == 0 and info["last_launch_jit"] == 1, info["last_error"]
        jit, img_jit, st_j = r.download_accum(), r.resolve(), r.stats()
    assert np.all(pre[:, 3] == S) and np.all(jit[:, 3] == S) and np.isfinite(jit).all()
    assert abs(st_j.rays - st_p.rays) <= 2e-3 * st_p.rays  # different FP32 root formulas: a few chaotic paths differ
    assert np.abs(img_pre - ref).mean() < 2e-3 and np.abs(img_jit - ref).mean() < 2e-3
    assert (np.abs(img_jit - ref) < 1e-4).mean() > 0.8


def _inside_scenes(gpu):
    """Scenes whose rays travel INSIDE a sphere that is large against epsilon (found by dev/fuzz_scenes.py): the plain
    root formulas of the index-in-key kernels are not good enough there, the layout must say so (embed_ok = 0)."""
    W, H = 64, 48
    _, cfg = gpu.builtin_scene("simple", W, H)
    ground = (1000.0, (0, -1000.5, -1.5), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)
    lamp = (0.2, (0.3, 0.0, -1.0), (4, 4, 4), (0.7, 0.7, 0.7), 0, 0)
    glass = (0.25, (-0.4, -0.2, -1.2), (0, 0, 0), (0.9, 0.9, 0.9), 2, 0)
    out = []
    # the camera 2 cm below the surface of the R = 1000 ground
    c = cfg.copy()
    c["position"][0], c["direction"][0], c["aperture"][0] = (0.0, -0.52, 0.5), (0.0, -0.6, -1.0), 0.0
    out.append(("camera inside the ground", [ground, lamp, glass], c))
    # the camera inside a r = 5 mirror ball that also holds a lamp: grazing rays stay grazing
    c = cfg.copy()
    c["position"][0], c["direction"][0], c["aperture"][0] = (0.5, 0.2, -1.0), (0.0, 0.0, -3.0), 0.0
    out.append(("camera inside a mirror ball", [(5.0, (0, 0, -1.5), (0, 0, 0), (0.9, 0.9, 0.9), 1, 0), lamp, glass], c))
    # a big glass ball seen from outside
    c = cfg.copy()
    out.append(("glass ball of radius 3", [(3.0, (0, 0, -5.0), (0, 0, 0), (0.95, 0.95, 0.95), 2, 0), lamp,
                                             (100.0, (0, -103.5, -5.0), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)], c))
    return W, H, out


def test_degenerate_radii(gpu, oracle_port):
    """Radius 0: never hit in the reference (discriminant -|perp|^2), so never hit here -- binary32 rounding noise must
    not turn a point into a target (it did: NaN pixels, found by dev/fuzz_scenes.py).  Negative radius: the reference
    only squares it and normalises P - centre, so it renders like |r|."""
    W, H, S = 64, 48, 4
    sph, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    point = np.zeros(1, dtype=gpu.SPHERE_DTYPE)
    point[0] = (0.0, (0.0, 0.0, -1.0), (1e6, 1e6, 1e6), (1.1, 1.1, 1.1), 2, 0)
    ref = oracle_port.render(point, cam, W, H, S, 2, 3, 0)
    for flags in (gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL,
                  gpu.PRECISION_FP32 | gpu.VARIANT_WAVEFRONT):
        with make_renderer(gpu, point, cam, W, H) as r:
            r.render(3, 0, S, flags)
            acc, img = r.download_accum(), r.resolve()
        assert np.isfinite(acc).all() and np.all(acc[:, 3] == S)
        assert np.abs(img - ref).max() < 1e-5  # sky everywhere
    neg = sph.copy()
    neg["radius"][1:] = -neg["radius"][1:]
    ref = oracle_port.render(neg, cam, W, H, S, 2, 3, 0)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(3, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        img_pos, lay_pos = r.resolve(), r.scene_layout()
    with make_renderer(gpu, neg, cam, W, H) as r:
        r.render(3, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        img_neg, lay_neg = r.resolve(), r.scene_layout()
    assert lay_pos == lay_neg
    assert np.array_equal(img_pos, img_neg)
    assert np.abs(img_neg - ref).mean() < 2e-3


@pytest.mark.parametrize("which", [0, 1, 2])
def test_rays_inside_large_spheres_take_the_exact_self_roots(gpu, oracle_port, which):
    W, H, scenes = _inside_scenes(gpu)
    label, rows, cfg = scenes[which]
    s = np.zeros(len(rows), dtype=gpu.SPHERE_DTYPE)
    for i, row in enumerate(rows):
        s[i] = row
    cam = gpu.camera_with_config(cfg)
    S = 6
    ref = oracle_port.render(s, cam, W, H, S, 2, 5, 0)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, s, cam, W, H) as r:
        assert r.scene_layout()["embed_ok"] == 0, label
        r.render(5, 0, S, flags | gpu.CODEGEN_PRECOMPILED)
        img_pre, st_pre = r.resolve(), r.stats()
        r.clear()
        r.render(5, 0, S, flags)
        r.clear()
        r.render(5, 0, S, flags)
        img_jit, st_jit = r.resolve(), r.stats()
        jit_ran = r.jit_info()["last_launch_jit"] == 1
    assert np.abs(img_pre - ref).mean() < 2e-3, label
    assert np.abs(img_jit - ref).mean() < 2e-3, label
    assert abs(st_jit.rays - st_pre.rays) <= 2e-3 * st_pre.rays, (label, jit_ran, st_jit.rays, st_pre.rays)


def test_runtime_compilation_of_a_20_sphere_scene_that_is_large_against_epsilon(gpu, oracle_port):
    """17 .. 32 spheres fit the unrolled scan only with full-precision keys (the index has 4 spare bits in the key): a scene
    that spreads over tens of units takes that path, and its run-time build must compile -- it once tripped over a
    static_assert meant for the index-in-key kernels only and fell back after a wasted compilation on every new scene."""
    W, H, S = 96, 54, 6
    rng = np.random.default_rng(20)
    _, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    sph = np.zeros(20, dtype=gpu.SPHERE_DTYPE)
    sph["radius"] = rng.uniform(0.3, 1.2, 20)
    sph["position"] = rng.uniform(-15.0, 15.0, (20, 3))
    sph["position"][:, 2] -= 20.0
    sph["color"] = rng.uniform(0.2, 0.9, (20, 3))
    sph["emission"][::5] = 3.0
    sph["reflection"] = rng.integers(0, 3, 20)
    ref = oracle_port.render(sph, cam, W, H, S, 2, 9, 0)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, sph, cam, W, H) as r:
        if not r.jit_info()["available"]:
            pytest.skip("run-time compilation not available")
        lay = r.scene_layout()
        assert lay["embed_ok"] == 0 and lay["fits_const"] == 1 and lay["specialised"] == 0
        r.render(9, 0, S, flags | gpu.CODEGEN_PRECOMPILED)
        pre, st_p = r.download_accum(), r.stats()
        r.clear()
        r.render(9, 0, S, flags)
        r.clear()
        r.render(9, 0, S, flags)
        info = r.jit_info()
        assert info["failures"] == 0 and info["last_launch_jit"] == 1, info["last_error"]
        jit, st_j, img = r.download_accum(), r.stats(), r.resolve()
    assert np.all(jit[:, 3] == S) and np.isfinite(jit).all()
    assert abs(st_j.rays - st_p.rays) <= 2e-3 * st_p.rays
    assert np.abs(img - ref).mean() < 2e-3


@pytest.mark.parametrize("refl", [0, 1, 2])
def test_runtime_compilation_of_a_one_sphere_scene(gpu, oracle_port, refl):
    """The smallest layouts (one small sphere, no big one; near-only or both-roots list alone)."""
    W, H, S = 64, 36, 6
    sph, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    one = sph[1:2].copy()
    one["reflection"][0] = refl
    ref = oracle_port.render(one, cam, W, H, S, 2, 6, 0)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, one, cam, W, H) as r:
        if not r.jit_info()["available"]:
            pytest.skip("run-time compilation not available")
        r.render(6, 0, S, flags)
        r.clear()
        r.render(6, 0, S, flags)
        info = r.jit_info()
        assert info["failures"] == 0 and info["last_launch_jit"] == 1, info["last_error"]
        acc = r.download_accum()
        assert np.all(acc[:, 3] == S) and np.isfinite(acc).all()
        assert np.abs(r.resolve() - ref).mean() < 2e-3


def test_runtime_compiled_in_place_megakernels(gpu):
    """The in-place megakernel is compiled for the scene too -- with the src/main.cpp integrator and with the sandbox
    one: same slots, same counters as the precompiled kernels."""
    W, H, S = 160, 90, 8
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    sb_sph, sb_cam = gpu.builtin_smallpt_scene()
    cases = [(sph, cam, None, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL),
             (sb_sph, None, sb_cam, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL | gpu.INTEGRATOR_SMALLPT)]
    for spheres, camera, sbcam, flags in cases:
        with gpu.Renderer(0) as r:
            r.upload_scene(spheres)
            if camera is not None:
                r.set_camera(camera)
            if sbcam is not None:
                r.set_smallpt_camera(sbcam)
            r.set_image(W, H, 2)
            if not r.jit_info()["available"]:
                pytest.skip("run-time compilation not available")
            r.render(3, 0, S, flags | gpu.CODEGEN_PRECOMPILED)
            pre, st_p = r.download_accum(), r.stats()
            r.clear()
            r.render(3, 0, S, flags)
            r.clear()
            r.render(3, 0, S, flags)
            info = r.jit_info()
            assert info["failures"] == 0 and info["last_launch_jit"] == 1, info["last_error"]
            jit, st_j = r.download_accum(), r.stats()
        assert np.all(pre[:, 3] == S) and np.all(jit[:, 3] == S)
        assert (st_j.rays, st_j.hits_diffuse, st_j.hits_specular, st_j.hits_dielectric) == \
            (st_p.rays, st_p.hits_diffuse, st_p.hits_specular, st_p.hits_dielectric)
        assert np.isclose(pre[:, :3], jit[:, :3], rtol=1e-5, atol=1e-5).all()


def test_runtime_compilation_follows_the_scene(gpu):
    """Different coefficients -> a different kernel; a scene without a specialised layout -> the precompiled path."""
    W, H = 64, 36
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, sph, cam, W, H) as r:
        if not r.jit_info()["available"]:
            pytest.skip("run-time compilation not available")
        r.render(1, 0, 2, flags)
        r.render(1, 2, 2, flags)
        n1 = r.jit_info()["compiled"]
        assert n1 >= 1 and r.jit_info()["last_launch_jit"] == 1
        moved = sph.copy()
        for step in range(3):  # geometry that changes on every upload is never compiled ...
            moved["position"][6][0] += 0.01  # nudge the mirror ball
            r.upload_scene(moved)
            r.render(1, 0, 2, flags)
            assert r.jit_info()["compiled"] == n1 and r.jit_info()["last_launch_jit"] == 0
        r.render(1, 2, 2, flags)  # ... until it stays
        assert r.jit_info()["compiled"] == n1 + 1 and r.jit_info()["last_launch_jit"] == 1
        big, cfg2 = gpu.builtin_scene("spheres10k", W, H)
        r.upload_scene(big)
        r.set_camera(gpu.camera_with_config(cfg2))
        r.render(1, 0, 1, flags)
        assert r.jit_info()["last_launch_jit"] == 0 and r.jit_info()["failures"] == 0


# ---- degenerate scenes ----------------------------------------------------------------------------------------------------
def test_empty_scene_is_all_sky(gpu, oracle_port):
    """No spheres: every ray misses (main.cpp:114-120).  FP64 equals the oracle to rounding, every FP32 variant agrees."""
    W, H, S = 48, 27, 3
    _, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    sph = gpu.builtin_scene("simple", W, H)[0][:0].copy()
    ref = oracle_port.render(sph, cam, W, H, S, 2, 4, 0)
    assert ref.min() > 0.4  # the sky gradient (0.5..1 per channel), nothing else
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(4, 0, S, gpu.PRECISION_FP64)
        assert np.abs(r.resolve() - ref).max() < 1e-12  # the order of the S additions into a slot is not fixed
        for v in (gpu.VARIANT_MEGAKERNEL, gpu.VARIANT_MEGAKERNEL_SORTED, gpu.VARIANT_WAVEFRONT):
            r.clear()
            r.render(4, 0, S, gpu.PRECISION_FP32 | v)
            st = r.stats()
            assert st.rays == st.paths == W * H * 4 * S and st.hits_diffuse == 0
            assert np.abs(r.resolve() - ref).max() < 1e-5


@pytest.mark.parametrize("refl", [0, 1, 2])
def test_single_sphere_of_each_material(gpu, oracle_port, refl):
    """One sphere in front of the camera, each material in turn (exercises the one-sphere kernel shapes; the glass
    ball is the both-roots-only list)."""
    W, H, S = 64, 36, 6
    sph, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    one = sph[1:2].copy()
    one["reflection"][0] = refl
    ref = oracle_port.render(one, cam, W, H, S, 2, 6, 0)
    with make_renderer(gpu, one, cam, W, H) as r:
        r.render(6, 0, S, gpu.PRECISION_FP64)
        assert (np.abs(r.resolve() - ref) <= 1e-9).mean() >= 0.998
        for v in (gpu.VARIANT_MEGAKERNEL, gpu.VARIANT_MEGAKERNEL_SORTED):
            r.clear()
            r.render(6, 0, S, gpu.PRECISION_FP32 | v)
            acc = r.download_accum()
            assert np.all(acc[:, 3] == S) and np.isfinite(acc).all()
            assert np.abs(r.resolve() - ref).mean() < 2e-3


# ---- semantics that must survive the boundary ------------------------------------------------------
def test_progressive_accumulation_and_sample_split(gpu):
    """render(0,8) == render(0,3) + render(3,5): what lets ranks split the samples of a sub-pixel."""
    W, H = 64, 40
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(5, 0, 8, gpu.PRECISION_FP64)
        a = r.resolve()
        r.clear()
        r.render(5, 0, 3, gpu.PRECISION_FP64)
        r.render(5, 3, 5, gpu.PRECISION_FP64)
        b = r.resolve()
        assert np.abs(a - b).max() < 1e-12
        r.clear()
        r.render(5, 0, 8, gpu.PRECISION_FP32)
        c = r.download_accum()
        r.clear()
        r.render(5, 3, 5, gpu.PRECISION_FP32)
        r.render(5, 0, 3, gpu.PRECISION_FP32)
        d = r.download_accum()
    assert np.all(c[:, 3] == 8) and np.all(d[:, 3] == 8)
    assert np.allclose(c[:, :3], d[:, :3], rtol=1e-5, atol=1e-5)  # same samples, different summation order


def test_resolve_clamps_per_subpixel_and_flips_rows(gpu):
    """main.cpp:181,195-196: mean -> clamp EACH stratum -> average; output row H-1-y."""
    W, H = 32, 20
    sph, cfg = gpu.builtin_scene("simple", W, H)  # directly visible E=30 light: strata means far above 1
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(2, 0, 16, gpu.PRECISION_FP32)
        acc = r.download_accum().astype(np.float64)
        img = r.resolve()
        rgb8 = r.resolve_rgb8()
    mean = acc[:, :3] / acc[:, 3:4]
    assert mean.max() > 1.5, "test needs over-bright strata"
    expect = np.clip(mean, 0.0, 1.0).reshape(H, W, 4, 3)
    expect = ((expect[:, :, 0] * 0.25 + expect[:, :, 1] * 0.25) + expect[:, :, 2] * 0.25) + expect[:, :, 3] * 0.25
    assert np.abs(img - expect[::-1]).max() < 1e-12
    wrong = np.clip(mean.reshape(H, W, 4, 3).mean(axis=2), 0, 1)[::-1]  # clamping the pixel instead
    assert np.abs(img - wrong).max() > 0.05
    assert np.array_equal(rgb8.astype(np.int32), np.round(np.clip(img, 0, 1) ** (1 / 2.2) * 255).astype(np.int32))


def test_spp_below_four_is_black_like_the_reference(gpu):
    img = gpu.render_scene("box", 16, 12, 3)  # main.cpp:206: 3 / 4 == 0 samples
    assert np.all(img == 0.0)


def test_edge_geometries(gpu, oracle_port):
    for (W, H) in ((1, 1), (3, 1), (7, 5), (33, 2)):  # far fewer / not a multiple of 32 sub-pixels
        sph, cfg = gpu.builtin_scene("box", W, H)
        cam = gpu.camera_with_config(cfg)
        ref = oracle_port.render(sph, cam, W, H, 5, 2, 31, 0)
        with make_renderer(gpu, sph, cam, W, H) as r:
            r.render(31, 0, 5, gpu.PRECISION_FP32)
            acc = r.download_accum()
            img = r.resolve()
        assert np.all(acc[:, 3] == 5)
        assert np.abs(img - ref).mean() < 2e-2


def test_single_sphere_and_strided_records(gpu, oracle_port):
    W, H = 48, 32
    one = np.zeros(1, dtype=gpu.SPHERE_DTYPE)
    one["radius"], one["position"], one["color"], one["emission"] = 0.5, (0, 0, -1), (0.5, 0.6, 0.7), (0.2, 0.1, 0.0)
    _, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(one, cam, W, H, 4, 2, 1, 0)
    padded = np.zeros((1, 128), dtype=np.uint8)  # a caller whose sphere struct is bigger than 88 bytes
    padded[0, :88] = one.view(np.uint8)
    with gpu.Renderer(0) as r:
        r.upload_scene(padded.reshape(-1), stride=128)
        r.set_camera(cam)
        r.set_image(W, H, 2)
        r.render(1, 0, 4, gpu.PRECISION_FP64)
        img = r.resolve()
    assert np.abs(img - ref).max() < 1e-9


def test_generic_kernel_for_unspecialised_scene(gpu, oracle_port):
    """A sphere count with no unrolled specialisation goes through the run-time-count kernel."""
    rng = np.random.default_rng(4)
    n = 23
    s = np.zeros(n, dtype=gpu.SPHERE_DTYPE)
    s["radius"] = rng.uniform(0.05, 0.3, n)
    s["position"] = rng.uniform(-1, 1, (n, 3)) + (0, 0, -1)
    s["color"] = rng.uniform(0.2, 0.9, (n, 3))
    s["emission"][::5] = (2, 2, 2)
    s["reflection"] = rng.integers(0, 3, n)
    s[0] = (1000.0, (0, -1001.2, -1), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)
    W, H, S = 80, 60, 6
    _, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(s, cam, W, H, S, 2, 9, 0)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, 60000)
    ohit, orad, _, _ = oracle_port.samples(s, cam, W, H, 2, 9, xs, ys, sx, sy, ss)
    with make_renderer(gpu, s, cam, W, H) as r:
        r.render(9, 0, S, gpu.PRECISION_FP32)
        img = r.resolve()
        r.clear()
        r.render(9, 0, S, gpu.PRECISION_FP64)
        img64 = r.resolve()
        hit, rad, _, _ = r.trace_samples(9, xs, ys, sx, sy, ss, gpu.PRECISION_FP32)
    assert (np.abs(img64 - ref) <= 1e-9).mean() >= 0.995
    assert (hit == ohit).mean() >= 0.9995
    assert (rel_err(rad, orad) <= 1e-3).mean() >= 0.97  # many glass/mirror spheres: more chaotic paths than the box
    se = np.sqrt((rad.var(axis=0) + orad.var(axis=0)) / len(rad))
    assert np.abs((rad.mean(axis=0) - orad.mean(axis=0)) / se).max() < 4.0
    assert np.abs(img - ref).mean() < 1e-2


@pytest.mark.parametrize("n", [16, 17, 64, 65, 400])
def test_random_scenes_on_both_sides_of_every_size_threshold(gpu, oracle_port, n):
    """Random overlapping spheres of every material around the library's size thresholds: <= 16 small spheres can be
    unrolled (run-time build), <= 64 take the run-time-count scan over constant memory, more go through the hierarchy.
    Each path must give the FP64 oracle's primary hits, an unbiased image, and -- scan against hierarchy -- the very
    same bounces."""
    rng = np.random.default_rng(100 + n)
    s = np.zeros(n, dtype=gpu.SPHERE_DTYPE)
    s["radius"] = rng.uniform(0.03, 0.25, n) * (1.0 if n <= 65 else 0.4)
    s["position"] = rng.uniform(-1.2, 1.2, (n, 3)) * (1, 0.6, 1) + (0, 0.1, -1.2)
    s["color"] = rng.uniform(0.2, 0.95, (n, 3))
    s["emission"][::7] = (3, 2.5, 2)
    s["reflection"] = rng.integers(0, 3, n)
    s[0] = (1000.0, (0, -1000.6, -1), (0, 0, 0), (0.5, 0.5, 0.5), 0, 0)  # the ground: a big sphere
    W, H, S = 96, 64, 4
    _, cfg = gpu.builtin_scene("simple", W, H)
    cam = gpu.camera_with_config(cfg)
    ref = oracle_port.render(s, cam, W, H, S, 2, 21, 0)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, 20000)
    ohit, orad, _, _ = oracle_port.samples(s, cam, W, H, 2, 21, xs, ys, sx, sy, ss)
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    with make_renderer(gpu, s, cam, W, H) as r:
        hit_a, rad_a, ray_a, _ = r.trace_samples(21, xs, ys, sx, sy, ss, gpu.PRECISION_FP32 | gpu.ACCEL_AUTO)
        hit_s, rad_s, ray_s, _ = r.trace_samples(21, xs, ys, sx, sy, ss, gpu.PRECISION_FP32 | gpu.ACCEL_SCAN)
        r.render(21, 0, S, flags)
        r.clear()
        r.render(21, 0, S, flags)  # second sight: the run-time build where the layout allows one
        acc, img, st_a = r.download_accum(), r.resolve(), r.stats()
        r.clear()
        r.render(21, 0, S, flags | gpu.ACCEL_SCAN | gpu.CODEGEN_PRECOMPILED)
        img_scan, st_s = r.resolve(), r.stats()
        r.clear()
        r.render(21, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL)
        img_inplace = r.resolve()
    assert np.array_equal(hit_a, hit_s) and np.array_equal(rad_a, rad_s) and np.array_equal(ray_a, ray_s)
    assert np.isfinite(rad_a).all() and np.isfinite(acc).all() and np.all(acc[:, 3] == S)
    assert (hit_a == ohit).mean() >= 0.999
    assert (rel_err(rad_a, orad) <= 1e-3).mean() >= 0.96
    se = np.sqrt((rad_a.var(axis=0) + orad.var(axis=0)) / len(rad_a))
    assert np.abs((rad_a.mean(axis=0) - orad.mean(axis=0)) / se).max() < 4.0
    assert abs(st_a.rays - st_s.rays) <= 3e-3 * st_s.rays
    for im in (img, img_scan, img_inplace):
        assert np.abs(im - ref).mean() < 1e-2
        assert abs(im.mean() - ref.mean()) < 5e-3


def test_scene_layouts_take_the_specialised_kernels(gpu):
    expect = {
        "box": dict(small_near=2, small_both=1, big_near=5, big_both=0, big_x=2, big_y=2, big_z=1, uniform_k=1, specialised=1),
        "box_mirror": dict(small_near=2, small_both=1, big_near=5, big_both=0, big_x=2, big_y=2, big_z=1, uniform_k=1, specialised=1),
        "simple": dict(small_near=3, small_both=1, big_near=1, big_both=0, big_x=0, big_y=1, big_z=0, uniform_k=1, specialised=1),
        "dof_glass": dict(small_near=2, small_both=2, big_near=1, big_both=0, big_x=0, big_y=1, big_z=0, uniform_k=1, specialised=1),
        "spheres10k": dict(fits_const=0, specialised=0, big_near=1),
    }
    for name, want in expect.items():
        sph, cfg = gpu.builtin_scene(name, 64, 48)
        with make_renderer(gpu, sph, gpu.camera_with_config(cfg), 64, 48) as r:
            got = r.scene_layout()
        for k, v in want.items():
            assert got[k] == v, (name, k, got)
    # a camera INSIDE an opaque sphere keeps both roots for it
    sph, cfg = gpu.builtin_scene("box", 64, 48)
    sph = sph.copy()
    sph["radius"][6] = 50.0  # the mirror ball now swallows the camera (and is "big")
    with make_renderer(gpu, sph, gpu.camera_with_config(cfg), 64, 48) as r:
        assert r.scene_layout()["big_both"] == 1


def test_two_contexts_on_one_gpu_from_two_threads(gpu):
    """Contexts are single-threaded, but two of them may live in two host threads on the same GPU: the precompiled
    kernels share one constant-memory scene per device, so concurrent renders must take turns and not mix scenes."""
    import threading
    W, H, S = 160, 90, 16
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED | gpu.CODEGEN_PRECOMPILED
    jobs = []
    for name in ("box", "box_mirror"):
        sph, cfg = gpu.builtin_scene(name, W, H)
        jobs.append((sph, gpu.camera_with_config(cfg)))
    alone = []
    for sph, cam in jobs:
        with make_renderer(gpu, sph, cam, W, H) as r:
            r.render(3, 0, S, flags)
            alone.append((r.download_accum(), r.stats().rays))
    together = [None, None]

    def work(i):
        sph, cam = jobs[i]
        with make_renderer(gpu, sph, cam, W, H) as r:
            rays = 0
            for k in range(8):  # many short renders so that the two threads really interleave
                r.render(3, 2 * k, 2, flags)
            together[i] = (r.download_accum(), r.stats().rays)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for (acc_a, rays_a), got in zip(alone, together):
        assert got is not None
        acc_t, rays_t = got
        assert rays_t == rays_a
        assert np.all(acc_t[:, 3] == S)
        assert np.isclose(acc_t[:, :3], acc_a[:, :3], rtol=1e-5, atol=1e-5).all()


def test_camera_inside_a_mirror_ball_keeps_full_precision_keys(gpu):
    """dev/fuzz_scenes.py seed 51 scene 1: a r = 0.0016 lamp and a camera inside a r = 0.25 mirror ball whose colour exceeds 1
    (no roulette death: 99 reflections per path off one concave surface).  With the list position riding in the key's low
    mantissa bits every hit distance is truncated the same way and the chord pattern drifts: the run-time compiled kernel
    lit 23 % of the pixels, the FP64 oracle and the full-precision keys 1 %.  Such scenes must not take the index-in-key
    layout, and the run-time build must trace what the precompiled kernels trace."""
    sph = np.zeros(3, dtype=gpu.SPHERE_DTYPE)
    sph["radius"] = [0.25, 0.014588611048801227, 0.001619232652198674]
    sph["position"] = [[-0.03296013, -0.07278335, -0.03793183], [-0.06246247, 0.00336563, -0.02177486],
                       [-0.06036928, 0.02285544, -0.0830866]]
    sph["color"] = [[1.04220626, 1.04026928, 1.18389618], [0.93259856, 0.83243481, 0.71187968], [0.23217308, 0.28617143, 0.50365051]]
    sph["emission"][2] = 1e6
    sph["reflection"] = 1
    W, H, S = 11, 67, 9
    _, cfg = gpu.builtin_scene("simple", W, H)
    cfg["position"][0] = [-0.05835709, 0.0216139, -0.08301402]
    cfg["direction"][0] = [-0.05715281, 0.02583565, -0.08152568]
    cfg["aperture"][0] = 0.0
    cfg["vertical_fov_radians"][0] = 0.97654284
    cfg["focus_distance"][0] = 0.19112383
    cam = gpu.camera_with_config(cfg)
    lit = {}
    with make_renderer(gpu, sph, cam, W, H) as r:
        assert r.scene_layout()["embed_ok"] == 0
        for label, flags, reps in (("precompiled", gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED | gpu.CODEGEN_PRECOMPILED, 1),
                                   ("run-time", gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED, 2),
                                   ("fp64", gpu.PRECISION_FP64, 1)):
            for _ in range(reps):
                r.clear()
                r.render(8, 0, S, flags)
            lit[label] = (r.stats().rays, float((r.resolve() > 0.5).mean()))
        assert r.jit_info()["compiled"] >= 1
    assert lit["run-time"][0] == lit["precompiled"][0]          # the same rays, to the last one
    assert lit["run-time"][1] == lit["precompiled"][1]
    assert abs(lit["precompiled"][1] - lit["fp64"][1]) < 0.02   # a per cent of the pixels see the lamp, not a quarter


def test_error_behaviour(gpu):
    with gpu.Renderer(0) as r:
        with pytest.raises(gpu.PtbError) as e:
            r.render(1, 0, 1)
        assert e.value.code == -4  # PTB_ERR_STATE: no scene yet
        sph, cfg = gpu.builtin_scene("box", 8, 8)
        bad = sph.copy()
        bad["reflection"][3] = 7
        with pytest.raises(gpu.PtbError) as e:
            r.upload_scene(bad)
        assert e.value.code == -1
        for field, value in (("radius", np.nan), ("position", np.inf), ("color", -np.inf), ("emission", np.nan)):
            bad = sph.copy()
            bad[field][2] = value
            with pytest.raises(gpu.PtbError) as e:
                r.upload_scene(bad)
            assert e.value.code == -1 and "non-finite" in str(e.value)
        # 2^24 spheres: refused before a byte is read (list positions travel in 24 bits, all ones = "no sphere")
        assert gpu._ptb_upload_scene(r._ctx, sph.ctypes.data, 1 << 24, sph.dtype.itemsize) == -1
        assert "2^24" in gpu._ptb_last_error(r._ctx).decode()
        r.upload_scene(sph)
        with pytest.raises(gpu.PtbError) as e:
            r.set_camera(np.zeros(10))
        assert e.value.code == -1
        nan_cam = gpu.camera_with_config(cfg).copy()
        nan_cam["u"][0][1] = np.nan
        with pytest.raises(gpu.PtbError) as e:
            r.set_camera(nan_cam)
        assert e.value.code == -1
        r.set_camera(gpu.camera_with_config(cfg))
        with pytest.raises(gpu.PtbError) as e:
            r.render(1, 0, 1)  # no image yet
        assert e.value.code == -4
        with pytest.raises(gpu.PtbError):
            r.set_image(0, 8, 2)
        r.set_image(8, 8, 2)
        with pytest.raises(gpu.PtbError) as e:
            r.render(1, 0, 1, 0x100000)  # unknown flag
        assert e.value.code == -1
        with pytest.raises(gpu.PtbError):
            r.render(1, 0xFFFFFFFF, 2)  # sample range overflow
        with pytest.raises(gpu.PtbError):
            r.trace_samples(1, [8], [0], [0], [0], [0])  # x outside the image
        r.render(1, 0, 1)
        assert r.stats().paths == 8 * 8 * 4
    with pytest.raises(gpu.PtbError) as e:
        gpu.Renderer(10_000)
    assert e.value.code == -2


# ---- BASELINE.json full sizes through size-independent properties -------------------------------------------------------------------
@pytest.mark.parametrize("cfg_name,scene,W,H", [("C2", "box", 1024, 768), ("C3", "box_mirror", 1920, 1080),
                                                ("C4", "dof_glass", 3840, 2160)])
def test_full_size_frames(gpu, cfg_name, scene, W, H):
    S = 2
    sph, cfg = gpu.builtin_scene(scene, W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(1, 0, S)
        r.render(1, S, S)  # progressive second pass
        acc = r.download_accum()
        st = r.stats()
        img = r.resolve()
    assert acc.shape[0] == W * H * 4
    assert np.all(acc[:, 3] == 2 * S), "sample count per sub-pixel"
    assert st.paths == W * H * 4 * 2 * S
    assert np.isfinite(acc).all() and acc[:, :3].min() >= 0.0
    assert img.shape == (H, W, 3) and img.min() >= 0.0 and img.max() <= 1.0
    expect_rpp = {"box": 12.33, "box_mirror": 12.33, "dof_glass": 2.05}[scene]
    assert abs(st.rays / st.paths - expect_rpp) < 0.05 * expect_rpp + 0.05
    # checksum of checksums: total radiance == sum over per-row sums, and the two halves of the frame are
    # statistically alike for the left/right-symmetric box scenes
    total = acc[:, :3].astype(np.float64).sum()
    rows = acc[:, :3].astype(np.float64).reshape(H, -1).sum(axis=1)
    assert abs(total - rows.sum()) <= 1e-9 * abs(total)


@pytest.mark.parametrize("scene,W,H", [("box", 1024, 768), ("box_mirror", 1920, 1080), ("dof_glass", 3840, 2160)])
def test_full_size_frames_fp32_against_the_fp64_kernel(gpu, scene, W, H):
    """BASELINE resolutions, same seed: the FP32 product path against the FP64 deterministic kernel (itself pinned to the
    oracle sample by sample).  Same uniforms, so the two images differ only by rounding and by the few chaotic paths:
    per-pixel agreement at full frame size, not a statistic of a small frame."""
    S = 2
    sph, cfg = gpu.builtin_scene(scene, W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(11, 0, S, gpu.PRECISION_FP64)
        img64, st64 = r.resolve(), r.stats()
        r.clear()
        r.render(11, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        img32, st32 = r.resolve(), r.stats()
    assert st32.paths == st64.paths == W * H * 4 * S
    assert abs(st32.rays - st64.rays) <= 2e-3 * st64.rays
    diff = np.abs(img32 - img64)
    assert diff.mean() < 1e-3, diff.mean()
    # 8 samples per pixel; a chaotic path moves its pixel, everything else agrees to FP32 rounding
    assert (diff.max(axis=2) < 1e-4).mean() > (0.93 if scene == "box_mirror" else 0.97)
    for c in range(3):
        assert abs(img32[..., c].mean() - img64[..., c].mean()) < 2e-4


def test_spheres10k_small_frame(gpu, oracle_port):
    """BASELINE config 5 geometry (10 001 spheres) on a small frame: the run-time-count kernel with geometry
    streamed from global memory."""
    W, H, S = 64, 36, 2
    sph, cfg = gpu.builtin_scene("spheres10k", W, H)
    cam = gpu.camera_with_config(cfg)
    rng = np.random.default_rng(2)
    xs, ys, sx, sy, ss = probe_inputs(rng, W, H, 3000)
    ohit, orad, _, _ = oracle_port.samples(sph, cam, W, H, 2, 6, xs, ys, sx, sy, ss)
    with make_renderer(gpu, sph, cam, W, H) as r:
        hit, rad, _, _ = r.trace_samples(6, xs, ys, sx, sy, ss, gpu.PRECISION_FP64)
        assert np.array_equal(hit, ohit)
        assert (rel_err(rad, orad) <= REL_TOL).mean() >= 0.995
        hit32, rad32, _, _ = r.trace_samples(6, xs, ys, sx, sy, ss, gpu.PRECISION_FP32)
        assert (hit32 == ohit).mean() >= 0.999
        r.render(6, 0, S)
        acc = r.download_accum()
    assert np.all(acc[:, 3] == S)
