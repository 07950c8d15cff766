"""The N > 1 host logic on CPU: world_size-2 gloo job that splits the samples of every sub-pixel,
sum-reduces the per-rank accumulation buffers and resolves on rank 0 -- the same calls bench.py makes with
NCCL.  The per-rank buffers are produced by the oracle (the GPU is not available here); what is under test
is the partition + reduce + resolve logic of cpu-path-tracing_b200/distributed.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_range_partitions(pkg):
    from cpu_path_tracing_b200.distributed import sample_range

    for total in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [sample_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for first, count in spans:
                assert first == pos and count >= 0
                pos += count
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sample_range(8, 2, 2)
    # the partition is the LIBRARY's (ptb_sample_share, what ptb_render applies on a rank of a job): host code, no GPU
    assert pkg.sample_share(1024, 8, 3) == (384, 128) and pkg.sample_share(10, 4, 1) == (3, 3)
    with pytest.raises(pkg.PtbError):
        pkg.sample_share(8, 0, 0)


def test_multi_gpu_entry_points_without_a_gpu(pkg):
    """No GPU here: the multi-GPU entry points must refuse, not fall back to anything."""
    import ctypes

    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PtbError) as e:
        pkg.Renderer([0, 1])
    assert e.value.code == -2  # PTB_ERR_NO_DEVICE from the first member
    lib = ctypes.CDLL(pkg.LIB_PATH)
    out = ctypes.c_void_p()
    devs = (ctypes.c_int * 2)(0, 0)
    assert lib.ptb_create_multi(devs, 2, ctypes.byref(out)) == -1  # a GPU listed twice
    assert lib.ptb_create_multi(None, 2, ctypes.byref(out)) == -1
    assert lib.ptb_comm_init_rank(None, None, 2, 0) == -1


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import Oracle

    pkg = load_package()
    from cpu_path_tracing_b200.distributed import reduce_accum_, sample_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, S, SEED = 40, 24, 9, 77
    sph, cfg = pkg.builtin_scene("box_mirror", W, H)
    cam = pkg.camera_with_config(cfg)
    orc = Oracle("port")
    first, count = sample_range(S, world, rank)
    _, sums = orc.render(sph, cam, W, H, max(count, 1), 2, SEED, first, nthreads=2, want_sums=True)
    acc = np.zeros((W * H * 4, 4), dtype=np.float32)  # the float4 {r,g,b,n} layout of the device buffer
    if count > 0:
        acc[:, :3] = sums
        acc[:, 3] = count
    t = torch.from_numpy(acc)
    reduce_accum_(t, 0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_matches_single_rank(pkg, oracle_port, tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "acc.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    acc = np.load(out).astype(np.float64)
    W, H, S, SEED = 40, 24, 9, 77
    sph, cfg = pkg.builtin_scene("box_mirror", W, H)
    cam = pkg.camera_with_config(cfg)
    ref_img, ref_sums = oracle_port.render(sph, cam, W, H, S, 2, SEED, 0, want_sums=True)
    assert np.all(acc[:, 3] == S)
    assert np.allclose(acc[:, :3], ref_sums, rtol=2e-6, atol=1e-6)  # float32 buffer
    mean = np.clip(acc[:, :3] / acc[:, 3:4], 0, 1).reshape(H, W, 4, 3)
    img = (((mean[:, :, 0] * 0.25 + mean[:, :, 1] * 0.25) + mean[:, :, 2] * 0.25) + mean[:, :, 3] * 0.25)[::-1]
    assert np.abs(img - ref_img).max() < 1e-5
