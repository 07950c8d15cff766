"""N > 1 on real GPUs (skipped unless >= 2 devices): torchrun-style 2-rank job over NCCL must give the image a
single GPU gives for the same seed (the stream is keyed by absolute sample index)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    pkg = load_package()
    from cpu_path_tracing_b200.distributed import DistributedRenderer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    W, H, S = 160, 90, 11
    sph, cfg = pkg.builtin_scene("box_mirror", W, H)
    cam = pkg.camera_with_config(cfg)
    dr = DistributedRenderer(pkg, rank, rank, world)
    dr.setup(sph, cam, W, H, 2)
    dr.step(5, S, pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL_SORTED)  # the product path
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out_path, dr.accum.cpu().numpy())
    dist.barrier()
    dr.close()
    dist.destroy_process_group()


def test_two_gpus_match_one(gpu, tmp_path):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "acc2.npy")
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 1000, out), nprocs=2, join=True)
    acc2 = np.load(out)
    W, H, S = 160, 90, 11
    sph, cfg = gpu.builtin_scene("box_mirror", W, H)
    cam = gpu.camera_with_config(cfg)
    with gpu.Renderer(0) as r:
        r.upload_scene(sph)
        r.set_camera(cam)
        r.set_image(W, H, 2)
        r.render(5, 0, S, gpu.PRECISION_FP32 | gpu.VARIANT_MEGAKERNEL_SORTED)
        acc1 = r.download_accum()
    assert np.all(acc2[:, 3] == S) and np.all(acc1[:, 3] == S)
    assert np.allclose(acc1[:, :3], acc2[:, :3], rtol=1e-5, atol=1e-5)
