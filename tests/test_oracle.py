"""The oracle (oracle/pt_oracle.c) against the reference ITSELF: committed golden vectors always,
the live compiled reference (oracle/_ref) where this checkout has it; plus structural properties."""
import numpy as np
import pytest

SCENES = ("simple", "box", "box_mirror")


# ---- golden vectors generated from the compiled reference (tests/golden/make_golden.py) -----------------
@pytest.mark.parametrize("name", SCENES)
def test_samples_match_reference_golden(oracle_port, golden, golden_scene, name):
    z = golden(f"samples_{name}.npz")
    W, H = int(z["width"]), int(z["height"])
    sph, _, cam = golden_scene(name, W, H)
    hit, rad, ray, draws = oracle_port.samples(sph, cam, W, H, int(z["nsub"]), int(z["seed"]), z["x"], z["y"], z["sx"],
                                               z["sy"], z["sample"])
    assert np.array_equal(hit, z["hit"])
    assert np.array_equal(ray, z["ray"])
    assert np.array_equal(draws.astype(np.uint32), z["draws"])
    assert np.array_equal(rad, z["radiance"]), "radiance must be bit-identical to the reference's radiance()"


@pytest.mark.parametrize("name", SCENES)
def test_stock_mt19937_rows_match_reference_golden(oracle_port, golden, name):
    """render_subpixel driven by the reference's STOCK stream (mt19937{0}), bit for bit."""
    z = golden(f"mtrows_{name}.npz")
    w, h = int(z["width"]), int(z["height"])
    from __graft_entry__ import load_package

    pkg = load_package()
    sph, cfg = pkg.builtin_scene(name, w, h)
    cam = pkg.camera_with_config(cfg)
    for y, row in zip(z["rows_y"], z["rows"]):
        img = oracle_port.mt_render(sph, cam, w, h, int(z["samps"]), int(z["nsub"]), seed_mode=1, y0=int(y), y1=int(y) + 1)
        assert np.array_equal(img[h - 1 - int(y)], row)


@pytest.mark.parametrize("name", SCENES)
def test_image_and_sums_match_reference_golden(oracle_port, golden, name):
    z = golden(f"image_{name}.npz")
    w, h = int(z["width"]), int(z["height"])
    from __graft_entry__ import load_package

    pkg = load_package()
    sph, cfg = pkg.builtin_scene(name, w, h)
    cam = pkg.camera_with_config(cfg)
    img, sums = oracle_port.render(sph, cam, w, h, int(z["samps"]), int(z["nsub"]), int(z["seed"]), int(z["first_sample"]),
                                   want_sums=True)
    assert np.array_equal(img, z["image"])
    assert np.array_equal(sums, z["sums"])


# ---- live: the compiled reference next to the restatement -----------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_live_counter_stream_bit_exact(oracle_port, ref_ctr, ref_stock, name):
    rng = np.random.default_rng(99)
    W, H, N = 200, 150, 6000
    sph, cfg, cam = ref_stock.scene(name, W, H)
    xs, ys = rng.integers(0, W, N), rng.integers(0, H, N)
    sx, sy = rng.integers(0, 3, N), rng.integers(0, 3, N)
    ss = rng.integers(0, 1 << 31, N)
    a = oracle_port.samples(sph, cam, W, H, 3, 12345678901234, xs, ys, sx, sy, ss)
    b = ref_ctr.samples(sph, cam, W, H, 3, 12345678901234, xs, ys, sx, sy, ss)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("name", SCENES)
def test_live_stock_stream_bit_exact(oracle_port, ref_stock, name):
    W, H = 96, 64
    sph, cfg, cam = ref_stock.scene(name, W, H)
    assert np.array_equal(oracle_port.mt_render(sph, cam, W, H, 3, 2), ref_stock.mt_render(sph, cam, W, H, 3, 2, seed_mode=1))


def test_live_intersect_and_color_to_int(oracle_port, ref_stock):
    rng = np.random.default_rng(3)
    sph, _, _ = ref_stock.scene("box_mirror", 64, 48)
    for _ in range(500):
        o = rng.normal(size=3) * 0.3 + (0, 0, 0.5)
        d = rng.normal(size=3)
        assert oracle_port.intersect(sph, o, d) == ref_stock.intersect(sph, o, d)
    v = rng.uniform(-0.5, 1.5, 4000)
    assert np.array_equal(oracle_port.color_to_int(v), ref_stock.color_to_int(v))


def test_reference_program_rows_reproducible(ref_stock, golden, tmp_path):
    """The shipped program, run twice: rows y%64==0 are identical (seed (unsigned short)(y^3) == 0)."""
    z = golden("refmain_rows.npz")
    assert ref_stock.reference_main(4, str(tmp_path)) == 0
    tok = open(tmp_path / "image.ppm").read().split()
    px = np.array(tok[4:], dtype=np.int32).reshape(768, 1024, 3)
    for y, row in zip(z["rows_y"], z["rows"]):
        assert np.array_equal(px[768 - 1 - int(y)], row.astype(np.int32))


# ---- structural properties ----------------------------------------
def test_path_statistics_match_survey(oracle_port, golden_scene):
    expect = {"box_mirror": (12.33, 0.530, 9.81, 0.988), "box": (12.33, 9.63, 0.577, 1.117), "simple": (2.085, 0.890, 0.032, 0.163)}
    for name, (rpp, dif, spe, die) in expect.items():
        sph, _, cam = golden_scene(name, 256, 192)
        oracle_port.stats_reset()
        oracle_port.render(sph, cam, 256, 192, 4, 2, 21)
        st = oracle_port.stats()
        p = st["paths"]
        assert p == 256 * 192 * 4 * 4
        assert abs(st["rays"] / p - rpp) < 0.03 * rpp
        assert abs(st["hit_diffuse"] / p - dif) < 0.05 * dif + 0.01
        assert abs(st["hit_specular"] / p - spe) < 0.05 * spe + 0.01
        assert abs(st["hit_dielectric"] / p - die) < 0.05 * die + 0.01
        # Russian roulette never fires before depth 5 (main.cpp:130): every kill needed >= 6 rays
        assert st["rr_draws"] <= st["rays"] - 5 * (p - st["misses"] - st["depth_limit"]) + 5 * st["misses"] + 10 * p
        if name != "simple":
            assert st["misses"] / p < 1e-3  # closed box


def test_closest_hit_tie_rule_and_no_hit(oracle_port, pkg):
    """Strict '<' in main.cpp:35: of two identical spheres the lower index wins; a miss returns -1 / t = inf."""
    s = np.zeros(2, dtype=pkg.SPHERE_DTYPE)
    s["radius"] = 1.0
    s["position"] = (0, 0, -5)
    idx, t = oracle_port.intersect(s, (0, 0, 0), (0, 0, -1))
    assert idx == 0 and abs(t - 4.0) < 1e-12
    idx, t = oracle_port.intersect(s, (0, 0, 0), (0, 1, 0))
    assert idx == -1 and t == 1e20
    # origin on the surface: the near root (t ~ 0 < epsilon) is skipped, the far root is taken (sphere.cpp:21-27)
    idx, t = oracle_port.intersect(s[:1], (0, 0, -4), (0, 0, -1))
    assert idx == 0 and abs(t - 2.0) < 1e-9
    # un-normalised direction: t scales with 1/|d| (camera.cpp:36 keeps |d| ~ focus distance)
    idx, t = oracle_port.intersect(s[:1], (0, 0, 0), (0, 0, -2))
    assert idx == 0 and abs(t - 2.0) < 1e-12
