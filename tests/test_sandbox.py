"""sandbox/main.cpp -- the reference's stand-alone smallpt fork.

CPU: the oracle restatement (oracle/pt_oracle_sandbox.c) against golden vectors produced by the sandbox's own
object code and by the sandbox PROGRAM, and live against oracle/_ref where it exists.
GPU: the smallpt integrator mode of the C ABI against the oracle."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def sb(golden):
    return golden("sandbox.npz")


def test_oracle_samples_bit_exact_to_sandbox_golden(oracle_port, sb):
    hit, rad, ray, draws = oracle_port.sb_samples(sb["spheres"], sb["cam8"], int(sb["width"]), int(sb["height"]),
                                                  int(sb["seed"]), sb["x"], sb["y"], sb["sx"], sb["sy"], sb["sample"])
    assert np.array_equal(hit, sb["hit"])
    assert np.array_equal(ray, sb["ray"])
    assert np.array_equal(draws.astype(np.uint32), sb["draws"])
    assert np.array_equal(rad, sb["radiance"])


def test_oracle_images_bit_exact_to_sandbox_golden(oracle_port, sb):
    a = oracle_port.sb_render(sb["spheres"], sb["cam8"], 96, 72, 2, mode=0)
    assert np.array_equal(a, sb["img_stock"]), "erand48 stream, Xi = {0,0,y^3} per row"
    b = oracle_port.sb_render(sb["spheres"], sb["cam8"], 96, 72, 3, mode=1, seed=21, first_sample=5)
    assert np.array_equal(b, sb["img_ctr"])


def test_oracle_reproduces_the_sandbox_program_output(oracle_port, sb):
    """16 rows of the image.ppm written by the sandbox program at 4 spp (1024x768), re-derived from scratch."""
    for y, row in zip(sb["program_rows_y"], sb["program_rows"]):
        img = oracle_port.sb_render(sb["spheres"], sb["cam8"], 1024, 768, 1, mode=0, y0=int(y), y1=int(y) + 1)
        assert np.array_equal(oracle_port.sb_to_int(img[768 - 1 - int(y)]), row.astype(np.int32)), f"row y={y}"


def test_live_sandbox_reference(oracle_port, tmp_path):
    from oracle import Oracle, available

    if not available("ref_sandbox"):
        pytest.skip("oracle/_ref not built here")
    ref = Oracle("ref_sandbox")
    sph, cam8 = ref.sb_scene()
    for mode in (0, 1):
        a = ref.sb_render(sph, cam8, 120, 90, 2, mode=mode, seed=3, first_sample=1)
        b = oracle_port.sb_render(sph, cam8, 120, 90, 2, mode=mode, seed=3, first_sample=1)
        assert np.array_equal(a, b)
    assert ref.sandbox_program(4, str(tmp_path)) == 0
    tok = open(tmp_path / "image.ppm").read().split()
    px = np.array(tok[4:], dtype=np.int32).reshape(768, 1024, 3)
    img = oracle_port.sb_render(sph, cam8, 1024, 768, 1, mode=0, y0=100, y1=104)
    assert np.array_equal(oracle_port.sb_to_int(img[768 - 104:768 - 100]), px[768 - 104:768 - 100])


def test_builtin_smallpt_scene_equals_sandbox(pkg, sb):
    sph, cam8 = pkg.builtin_smallpt_scene()
    assert np.array_equal(sph.view(np.uint8).reshape(-1, 88)[:, :84], sb["spheres"][:, :84])
    assert np.array_equal(cam8, sb["cam8"])


# ---- GPU ------------------------------------------------------------------------------------------------------
def _renderer(gpu, sb, W, H):
    r = gpu.Renderer(0)
    r.upload_scene(sb["spheres"])
    r.set_smallpt_camera(sb["cam8"])
    r.set_image(W, H, 2)
    return r


@pytest.mark.gpu
def test_gpu_smallpt_fp64_samples(gpu, oracle_port, sb):
    W, H = int(sb["width"]), int(sb["height"])
    flags = gpu.PRECISION_FP64 | gpu.INTEGRATOR_SMALLPT
    with _renderer(gpu, sb, W, H) as r:
        hit, rad, ray, draws = r.trace_samples(int(sb["seed"]), sb["x"], sb["y"], sb["sx"], sb["sy"], sb["sample"], flags)
    assert np.array_equal(hit, sb["hit"]), "primary-hit indices must be bit-exact"
    assert np.array_equal(ray, sb["ray"])
    rel = np.abs(rad - sb["radiance"]).max(axis=1) / np.maximum(np.abs(sb["radiance"]).max(axis=1), 1e-12)
    assert (rel <= 1e-4).mean() >= 0.999
    assert (draws == sb["draws"]).mean() >= 0.999


@pytest.mark.gpu
def test_gpu_smallpt_images(gpu, oracle_port, sb):
    W, H, S = 96, 72, 6
    ref = oracle_port.sb_render(sb["spheres"], sb["cam8"], W, H, S, mode=1, seed=8)
    with _renderer(gpu, sb, W, H) as r:
        r.render(8, 0, S, gpu.PRECISION_FP64 | gpu.INTEGRATOR_SMALLPT)
        img64 = r.resolve()
        r.clear()
        r.render(8, 0, S, gpu.PRECISION_FP32 | gpu.INTEGRATOR_SMALLPT)
        img32 = r.resolve()
        acc = r.download_accum()
        st = r.stats()
    assert (np.abs(img64 - ref) <= 1e-9).mean() >= 0.995
    assert np.all(acc[:, 3] == S)
    assert np.abs(img32 - ref).mean() < 3e-3
    assert st.paths == W * H * 4 * S and st.rays > 3 * st.paths


@pytest.mark.gpu
def test_gpu_smallpt_fp32_unbiased(gpu, oracle_port, sb):
    rng = np.random.default_rng(5)
    W, H, n = 256, 192, 200000
    xs, ys = rng.integers(0, W, n), rng.integers(0, H, n)
    sx, sy, ss = rng.integers(0, 2, n), rng.integers(0, 2, n), rng.integers(0, 1 << 20, n)
    ohit, orad, _, _ = oracle_port.sb_samples(sb["spheres"], sb["cam8"], W, H, 4, xs, ys, sx, sy, ss)
    with _renderer(gpu, sb, W, H) as r:
        hit, rad, _, _ = r.trace_samples(4, xs, ys, sx, sy, ss, gpu.PRECISION_FP32 | gpu.INTEGRATOR_SMALLPT)
    assert (hit == ohit).mean() >= 0.9999
    rel = np.abs(rad - orad).max(axis=1) / np.maximum(np.abs(orad).max(axis=1), 1e-12)
    assert (rel <= 1e-3).mean() >= 0.98
    se = np.sqrt((rad.var(axis=0) + orad.var(axis=0)) / n)
    assert np.abs((rad.mean(axis=0) - orad.mean(axis=0)) / se).max() < 4.0
