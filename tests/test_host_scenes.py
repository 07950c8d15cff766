"""Host-side mirror of the reference's scene layer (host/pt.hpp behind the C ABI) against the
reference's own builders: golden fixtures always, the live oracle/_ref build where present."""
import os

import numpy as np
import pytest

SIZES = ((1024, 768), (1920, 1080), (256, 192))


@pytest.mark.parametrize("name", ["simple", "box", "box_mirror"])
@pytest.mark.parametrize("size", SIZES)
def test_builtin_scene_bytes_equal_reference(pkg, golden_scene, name, size):
    w, h = size
    g_sph, g_cfg, g_cam = golden_scene(name, w, h)
    sph, cfg = pkg.builtin_scene(name, w, h)
    assert sph.view(np.uint8).reshape(-1, 88).shape == g_sph.shape
    # padding bytes after `reflection` are unspecified in the reference struct: compare the 84 defined bytes
    assert np.array_equal(sph.view(np.uint8).reshape(-1, 88)[:, :84], g_sph[:, :84])
    assert np.array_equal(cfg.view(np.uint8).ravel(), g_cfg)
    cam = pkg.camera_with_config(cfg)
    assert np.array_equal(cam.view(np.uint8).ravel(), g_cam), "camera::with_config differs from the reference bit pattern"


@pytest.mark.parametrize("name", ["simple", "box", "box_mirror"])
def test_builtin_scene_equals_live_reference(pkg, ref_stock, name):
    for (w, h) in ((640, 480), (3840, 2160), (17, 5)):
        r_sph, r_cfg, r_cam = ref_stock.scene(name, w, h)
        sph, cfg = pkg.builtin_scene(name, w, h)
        assert np.array_equal(sph.view(np.uint8).reshape(-1, 88)[:, :84], r_sph[:, :84])
        assert np.array_equal(cfg.view(np.uint8).ravel(), r_cfg)
        assert np.array_equal(pkg.camera_with_config(cfg).view(np.uint8).ravel(), r_cam)


def test_camera_with_config_random_configs(pkg, oracle_port):
    rng = np.random.default_rng(5)
    for _ in range(200):
        cfg = np.zeros(1, dtype=pkg.CAMERA_CONFIG_DTYPE)
        cfg["position"] = rng.normal(size=3) * 3
        cfg["direction"] = rng.normal(size=3)
        cfg["up"] = (0, 1, 0)
        cfg["aspect_ratio"] = rng.uniform(0.5, 2.5)
        cfg["vertical_fov_radians"] = rng.uniform(0.2, 2.0)
        cfg["aperture"] = rng.uniform(0, 0.5)
        cfg["focus_distance"] = rng.uniform(0.5, 10)
        assert np.array_equal(pkg.camera_with_config(cfg).view(np.uint8).ravel(), oracle_port.camera_with_config(cfg))


def test_dof_glass_scene(pkg):
    base, bcfg = pkg.builtin_scene("simple", 3840, 2160)
    sph, cfg = pkg.builtin_scene("dof_glass", 3840, 2160)
    assert len(sph) == 5
    assert sph["reflection"].tolist() == [0, 2, 2, 0, 0]  # both r=0.5 side spheres are glass
    other = [f for f in sph.dtype.names if f != "reflection"]
    for f in other:
        assert np.array_equal(sph[f], base[f])
    assert cfg["aperture"][0] == 0.4 and cfg["focus_distance"][0] == bcfg["focus_distance"][0]


def test_spheres10k_scene(pkg):
    sph, cfg = pkg.builtin_scene("spheres10k", 1920, 1080)
    assert len(sph) == 10001
    assert sph["radius"][0] == 1000.0 and np.all(sph["radius"][1:] == 0.2)
    assert np.all(sph["position"][1:, 1] == 0.2)
    frac = np.bincount(sph["reflection"][1:], minlength=3) / 10000.0
    assert abs(frac[0] - 0.70) < 0.02 and abs(frac[1] - 0.20) < 0.02 and abs(frac[2] - 0.10) < 0.02
    assert 100 < int((sph["emission"][:, 0] > 0).sum()) < 300
    sph2, _ = pkg.builtin_scene("spheres10k", 640, 480)  # geometry does not depend on the image size
    assert np.array_equal(sph.view(np.uint8), sph2.view(np.uint8))
    assert cfg["focus_distance"][0] == 10.0 and np.allclose(cfg["position"][0], (13, 2, 3))


def test_write_ppm_matches_reference_format(pkg, oracle_port, tmp_path):
    rng = np.random.default_rng(2)
    img = rng.uniform(-0.2, 1.3, size=(7, 9, 3))
    path = os.path.join(tmp_path, "image.ppm")
    pkg.write_ppm(path, img)
    text = open(path).read()
    assert text.startswith("P3\n9 7\n255\n")
    vals = np.array(text.split()[4:], dtype=np.int32)
    assert np.array_equal(vals, oracle_port.color_to_int(img).ravel())
    assert text.endswith(" ") and text.count("\n") == 3  # main.cpp:241-246: one token + space per channel, no line breaks


def test_write_ppm_rgb8_p3_is_the_reference_layout_and_p6_is_raw(pkg, oracle_port, tmp_path):
    """main.cpp:240-247 from 8-bit values (ptb_resolve_rgb8 does color_to_int on the GPU): same bytes as ptb_write_ppm."""
    rng = np.random.default_rng(5)
    img = rng.uniform(-0.2, 1.3, size=(11, 13, 3))
    img[0, 0] = (0.0, 1.0, 0.5)
    rgb8 = oracle_port.color_to_int(img).astype(np.uint8)
    a, b, c = (os.path.join(tmp_path, n) for n in ("a.ppm", "b.ppm", "c.ppm"))
    pkg.write_ppm(a, img)
    pkg.write_ppm_rgb8(b, rgb8, binary=False)
    assert open(a, "rb").read() == open(b, "rb").read()
    pkg.write_ppm_rgb8(c, rgb8, binary=True)
    raw = open(c, "rb").read()
    head = b"P6\n13 11\n255\n"
    assert raw.startswith(head) and raw[len(head):] == rgb8.tobytes()
    lib = __import__("ctypes").CDLL(pkg.LIB_PATH)
    assert lib.ptb_write_ppm_rgb8(None, None, 1, 1, 1) == -1


def test_refmain_golden_rows_are_the_shipped_scene(golden, golden_scene, oracle_port):
    """The reference PROGRAM's reproducible rows (y%64==0) re-derived by the oracle from mt19937{0}."""
    z = golden("refmain_rows.npz")
    sph, _, cam = golden_scene("box_mirror", 1024, 768)
    for y, row in zip(z["rows_y"], z["rows"]):
        img = oracle_port.mt_render(sph, cam, 1024, 768, int(z["spp"]) // 4, 2, seed_mode=1, y0=int(y), y1=int(y) + 1)
        got = oracle_port.color_to_int(img[768 - 1 - int(y)])
        assert np.array_equal(got, row.astype(np.int32)), f"row y={y}"
