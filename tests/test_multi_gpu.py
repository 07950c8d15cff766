"""Several GPUs behind the C ABI (include/ptb200.h: ptb_create_multi, ptb_comm_init_rank; csrc/ptb_multi.cpp).

The stream is keyed by the absolute sample index, so however the samples are split over GPUs the image must be the
one a single GPU gives.  The one-GPU cases run everywhere (they drive the same code with a group
/ a job of size one); the rest needs two devices (`gpurun --gpus 2`).
"""
import os
import sys

import numpy as np
import pytest

from conftest import make_renderer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

W, H, S, SEED = 160, 90, 11, 5


def _single(gpu, flags, name="box_mirror", w=W, h=H, s=S):
    sph, cfg = gpu.builtin_scene(name, w, h)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, w, h) as r:
        r.render(SEED, 0, s, flags)
        return r.download_accum(), r.resolve(), r.resolve_rgb8(), r.stats()


def test_group_of_one_is_the_plain_context(gpu):
    """ptb_create_multi with one device: every entry point goes through the group dispatch, same slots, same image."""
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    acc1, img1, rgb1, st1 = _single(gpu, flags)
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H, device=[0]) as r:
        r.render(SEED, 0, S, flags)
        acc, img, rgb, st = r.download_accum(), r.resolve(), r.resolve_rgb8(), r.stats()
        info = r.comm_info()
        assert r.scene_layout()["specialised"] == 1
        with pytest.raises(gpu.PtbError):
            r.accum_buffer()
        r.clear()
        assert r.stats().paths == 0 and np.all(r.download_accum() == 0)
    assert info["n_gpus"] == 1 and info["mode"] == "one process" and info["last_transport"] == "peer"
    # same paths; the float additions into a slot (red.global.add) land in a different order from run to run
    assert np.all(acc[:, 3] == S) and np.allclose(acc[:, :3], acc1[:, :3], rtol=1e-5, atol=1e-5)
    assert np.abs(img - img1).max() < 1e-5 and np.abs(rgb.astype(int) - rgb1.astype(int)).max() <= 1
    assert st.paths == st1.paths and st.rays == st1.rays and st.last_resolve_ms > 0


def _rank_worker(rank, world, port, out_path, transport, flags, uid_path):
    sys.path.insert(0, ROOT)
    import time

    from __graft_entry__ import load_package

    pkg = load_package()
    # the side channel for the 128-byte id is a file here: the library needs nothing else from the launcher
    if rank == 0:
        uid = pkg.comm_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(uid_path + ".tmp", uid_path)
    else:
        while not os.path.exists(uid_path):
            time.sleep(0.01)
        uid = open(uid_path, "rb").read()
    sph, cfg = pkg.builtin_scene("box_mirror", W, H)
    cam = pkg.camera_with_config(cfg)
    with pkg.Renderer(rank % pkg.device_count()) as r:
        r.comm_init_rank(uid, world, rank)
        r.comm_set_transport(transport)
        r.upload_scene(sph)
        r.set_camera(cam)
        r.set_image(W, H, 2)
        r.render(SEED, 0, S, flags)
        first = r.resolve() if rank == 0 else r.resolve_collective()
        # progressive: the sum must not have consumed the buffers -- resolve again, as 8 bit, then add samples
        rgb8 = r.resolve_rgb8() if rank == 0 else r.resolve_collective()
        r.render(SEED, S, 3, flags)
        second = r.resolve() if rank == 0 else r.resolve_collective()
        r.set_image(W // 2, H // 2, 2)  # buffers move: peers must re-map (collective)
        r.render(SEED, 0, 4, flags)
        small = r.resolve() if rank == 0 else r.resolve_collective()
        if rank == 0:
            np.savez(out_path, first=first, rgb8=rgb8, second=second, small=small,
                     transport=r.comm_info()["last_transport"], paths=r.stats().paths)


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_job_of_one_rank(gpu, tmp_path, transport):
    """ptb_comm_init_rank with n_ranks = 1 on the driver's one-GPU box: communicator, record exchange, both transports."""
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    out = str(tmp_path / "r.npz")
    _rank_worker(0, 1, 0, out, {"nccl": gpu.TRANSPORT_NCCL, "peer": gpu.TRANSPORT_PEER}[transport], flags, str(tmp_path / "uid"))
    z = np.load(out)
    _, img1, rgb1, _ = _single(gpu, flags)
    assert str(z["transport"]) == transport
    assert np.abs(z["first"] - img1).max() < 1e-5 and np.abs(z["rgb8"].astype(int) - rgb1.astype(int)).max() <= 1


@pytest.mark.parametrize("precision", ["PRECISION_FP32", "PRECISION_FP64"])
@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_two_processes_match_one_gpu(gpu, tmp_path, transport, precision):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    flags = getattr(gpu, precision) | (gpu.VARIANT_MEGAKERNEL_SORTED if precision == "PRECISION_FP32" else 0)
    out = str(tmp_path / "r.npz")
    tr = {"nccl": gpu.TRANSPORT_NCCL, "peer": gpu.TRANSPORT_PEER}[transport]
    mp.spawn(_rank_worker, args=(2, 0, out, tr, flags, str(tmp_path / "uid")), nprocs=2, join=True)
    z = np.load(out)
    assert str(z["transport"]) == transport, "the requested transport must be the one that ran on an NVLink box"
    _, img1, rgb1, _ = _single(gpu, flags)
    _, img2, _, _ = _single(gpu, flags, s=S + 3)
    _, img3, _, _ = _single(gpu, flags, w=W // 2, h=H // 2, s=4)
    tol = 1e-5 if precision == "PRECISION_FP32" else 1e-12
    assert np.abs(z["first"] - img1).max() < tol
    assert np.abs(z["rgb8"].astype(int) - rgb1.astype(int)).max() <= (1 if precision == "PRECISION_FP32" else 0)
    assert np.abs(z["second"] - img2).max() < tol
    assert np.abs(z["small"] - img3).max() < tol
    assert int(z["paths"]) < W * H * 4 * (S + 3 + 4)  # rank 0 traced its share only


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_one_process_two_gpus_match_one(gpu, transport):
    """ptb_create_multi: what the patched reference main() uses."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    acc1, img1, rgb1, st1 = _single(gpu, flags)
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H, device=[0, 1]) as r:
        r.comm_set_transport({"nccl": gpu.TRANSPORT_NCCL, "peer": gpu.TRANSPORT_PEER}[transport])
        r.render(SEED, 0, S, flags)
        img, rgb, acc, st = r.resolve(), r.resolve_rgb8(), r.download_accum(), r.stats()
        info = r.comm_info()
        # FP64 parity mode across the GPUs: the per-slot sums are exact in double up to the order of two additions
        r.clear()
        r.render(SEED, 0, 3, gpu.PRECISION_FP64)
        img64 = r.resolve()
    assert info["n_gpus"] == 2 and info["last_transport"] == transport
    assert np.all(acc[:, 3] == S) and np.allclose(acc[:, :3], acc1[:, :3], rtol=1e-5, atol=1e-5)
    assert np.abs(img - img1).max() < 1e-5 and np.abs(rgb.astype(int) - rgb1.astype(int)).max() <= 1
    assert st.paths == st1.paths and st.rays == st1.rays
    _, ref64, _, _ = _single(gpu, gpu.PRECISION_FP64, s=3)
    assert np.abs(img64 - ref64).max() < 1e-12


def test_checkpoint_moves_between_gpu_counts(gpu):
    """download_accum of a two-GPU group = what one GPU would hold; restored into one GPU it continues the render."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    flags = gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED
    sph, cfg = gpu.builtin_scene("box", W, H)
    cam = gpu.camera_with_config(cfg)
    with make_renderer(gpu, sph, cam, W, H, device=[0, 1]) as r:
        r.render(SEED, 0, 6, flags)
        ckpt = r.download_accum()
    with make_renderer(gpu, sph, cam, W, H) as r:
        r.upload_accum(ckpt)
        r.render(SEED, 6, 5, flags)
        resumed = r.resolve()
    _, whole, _, _ = _single(gpu, flags, name="box", s=11)
    assert np.abs(resumed - whole).max() < 1e-5
