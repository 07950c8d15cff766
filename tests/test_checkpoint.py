"""Checkpoint / resume (the reference's own TODO, README.md:9 "save progress to resume").

ptb_download_accum is the checkpoint of a progressive render, ptb_upload_accum restores it into a NEW context; the
stream is keyed by the absolute sample index, so save -> destroy -> create -> restore -> continue must give the slots of
an uninterrupted run: bit for bit in the FP64 parity mode, to float summation order in FP32.
"""
import numpy as np
import pytest

from conftest import make_renderer

pytestmark = pytest.mark.gpu

W, H, SEED = 96, 54, 21


@pytest.mark.parametrize("name", ["box_mirror", "dof_glass"])
def test_fp32_save_destroy_restore_continue(gpu, tmp_path, name):
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    sph, cfg = gpu.builtin_scene(name, W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(SEED, 0, 12, flags)
        whole, whole_img = r.download_accum(), r.resolve()
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(SEED, 0, 5, flags)
        np.save(tmp_path / "ckpt.npy", r.download_accum())  # the checkpoint goes through a file, the context goes away
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.upload_accum(np.load(tmp_path / "ckpt.npy"))
        assert np.all(r.download_accum()[:, 3] == 5)
        r.render(SEED, 5, 7, flags)
        resumed, resumed_img = r.download_accum(), r.resolve()
    assert np.all(resumed[:, 3] == 12)
    assert np.allclose(resumed[:, :3], whole[:, :3], rtol=1e-5, atol=1e-5)
    assert np.abs(resumed_img - whole_img).max() < 1e-5


def test_fp64_resume_is_bit_exact(gpu):
    sph, cfg = gpu.builtin_scene("box", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        assert np.all(r.download_accum64() == 0)  # nothing rendered yet
        r.render(SEED, 0, 4, gpu.PRECISION_FP64)
        whole, whole_img = r.download_accum64(), r.resolve()
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.render(SEED, 0, 1, gpu.PRECISION_FP64)
        ckpt = r.download_accum64()
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.upload_accum64(ckpt)
        r.render(SEED, 1, 3, gpu.PRECISION_FP64)
        resumed, resumed_img = r.download_accum64(), r.resolve()
    # one thread owns a slot and adds its samples in index order: the same additions in the same order
    assert np.array_equal(resumed, whole) and np.array_equal(resumed_img, whole_img)


def test_restore_checks_the_geometry(gpu):
    sph, cfg = gpu.builtin_scene("box", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H) as r:
        with pytest.raises(gpu.PtbError) as e:
            r.upload_accum(np.zeros((W * H * 4 - 1, 4), dtype=np.float32))
        assert e.value.code == -1
        with pytest.raises(gpu.PtbError):
            r.upload_accum64(np.zeros((7, 4), dtype=np.float64))
    with gpu.Renderer(0) as r:
        with pytest.raises(gpu.PtbError) as e:
            r.upload_accum(np.zeros((4, 4), dtype=np.float32))
        assert e.value.code == -4  # no image yet
