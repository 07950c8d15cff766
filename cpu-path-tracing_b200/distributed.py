"""One process per GPU (torchrun): the launcher-side glue around the library's own multi-GPU path.

The reference parallelises over image ROWS inside one process (one taskflow task per row,
/root/reference/src/main.cpp:217-236).  Here the samples of every sub-pixel are split across the GPUs and the
un-clamped per-sub-pixel sums are added before the non-linear resolve (main.cpp:192-196).  All of that lives behind
the C ABI (include/ptb200.h, csrc/ptb_multi.cpp): ptb_comm_init_rank makes this process a rank of the job, ptb_render
then traces the rank's share (ptb_sample_share), and ptb_resolve* are collective -- one fused kernel per GPU over peer
mappings of all accumulation buffers (reduce-scatter + resolve + gather), or ncclReduce + resolve on the root.

What is left for this file: carrying the 128-byte communicator id from rank 0 to the others over the launcher's own
side channel (torch.distributed here -- plumbing, like MPI_Bcast would be under mpirun).  A single process that drives
all GPUs needs none of it: ``Renderer([0, 1, ..., 7])`` (ptb_create_multi).
"""
from __future__ import annotations


def sample_range(total_samples: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first, count) of rank's contiguous share: the library's partition (ptb_sample_share)."""
    import sys

    pkg = sys.modules[__name__.rsplit(".", 1)[0]]
    if world_size < 1 or not (0 <= rank < world_size) or total_samples < 0:
        raise ValueError("bad (total_samples, world_size, rank)")
    return pkg.sample_share(total_samples, world_size, rank)


def reduce_accum_(accum, dst: int = 0, group=None):
    """In-place SUM-reduce of an accumulation tensor onto rank `dst` through torch.distributed -- the caller-side
    alternative (ptb_accum_buffer / ptb_set_accum_buffer) for hosts that run their own collective layer, and what the
    world-size-2 gloo test on CPU exercises."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def broadcast_comm_id(pkg, rank: int, world_size: int) -> bytes:
    """ptb_comm_unique_id on rank 0, handed to the other ranks through torch.distributed."""
    import torch.distributed as dist

    box = [pkg.comm_unique_id() if rank == 0 else None]
    if world_size > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


class DistributedRenderer:
    """One rank of a one-process-per-GPU job.

    Usage (every rank): dr = DistributedRenderer(pkg, device, rank, world); dr.setup(...);
    dr.step(seed, total_samples, flags) -> on rank 0 the resolved image lives on the device.
    """

    def __init__(self, pkg, device: int, rank: int, world_size: int, transport: int = 0):
        self.pkg, self.rank, self.world = pkg, rank, world_size
        self.renderer = pkg.Renderer(device)
        if world_size > 1:
            self.renderer.comm_init_rank(broadcast_comm_id(pkg, rank, world_size), world_size, rank)
            if transport:
                self.renderer.comm_set_transport(transport)

    def setup(self, spheres, camera, width: int, height: int, nsub: int = 2, smallpt_camera=None):
        """camera: a 176-byte pt::camera (src/ integrator) and/or smallpt_camera: cam8 (sandbox integrator)."""
        r = self.renderer
        r.upload_scene(spheres)
        if camera is not None:
            r.set_camera(camera)
        if smallpt_camera is not None:
            r.set_smallpt_camera(smallpt_camera)
        r.set_image(width, height, nsub)  # collective when world_size > 1

    def step(self, seed: int, total_samples: int, flags: int = 0, resolve: bool = True):
        """One frame: clear, trace this rank's share of the samples, sum + resolve across the GPUs (device-side)."""
        r = self.renderer
        r.clear()
        r.render(seed, 0, total_samples, flags)  # a rank of a job traces ptb_sample_share(total, world, rank)
        if resolve:
            return r.resolve_device()  # collective; the address is the root's, None elsewhere
        return None

    def close(self):
        self.renderer.close()
