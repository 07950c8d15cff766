"""Multi-GPU host logic: split the samples of every sub-pixel across ranks, sum the
accumulation buffers onto the root, resolve there.

The reference parallelises over image ROWS inside one process (one taskflow task per
row, /root/reference/src/main.cpp:217-236).  Here one process drives one GPU (torchrun);
rank g of G traces samples [g*S/G, (g+1)*S/G) of EVERY sub-pixel -- the stream is keyed by
the absolute sample index, so the image does not depend on G -- into its own float4
accumulation buffer; one NCCL sum-reduce over NVLink brings the un-clamped per-sub-pixel
sums (and their sample counts, carried in .w) to rank 0, where the non-linear resolve
(mean -> clamp -> average, main.cpp:192-196) runs once.  No other exchange exists on
this path, so there is nothing to fuse a collective into.

torch is plumbing here (device buffers as tensors, torch.distributed for the collective).
"""
from __future__ import annotations


def sample_range(total_samples: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first, count) of rank's contiguous share; shares differ by at most one sample."""
    if world_size < 1 or not (0 <= rank < world_size) or total_samples < 0:
        raise ValueError("bad (total_samples, world_size, rank)")
    base, extra = divmod(total_samples, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def reduce_accum_(accum, dst: int = 0, group=None):
    """In-place SUM-reduce of an accumulation tensor onto rank `dst` (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


class DistributedRenderer:
    """A Renderer whose accumulation buffer is a torch tensor, plus the reduce.

    Usage (every rank): dr = DistributedRenderer(pkg, device, rank, world); dr.setup(...);
    dr.step(seed, total_samples) -> on rank 0 the resolved image lives on the device.
    """

    def __init__(self, pkg, device: int, rank: int, world_size: int):
        import torch

        self.pkg, self.rank, self.world = pkg, rank, world_size
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.renderer = pkg.Renderer(device)
        self.accum = None

    def setup(self, spheres, camera, width: int, height: int, nsub: int = 2, smallpt_camera=None):
        """camera: a 176-byte pt::camera (src/ integrator) and/or smallpt_camera: cam8 (sandbox integrator)."""
        torch = self.torch
        r = self.renderer
        r.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        r.upload_scene(spheres)
        if camera is not None:
            r.set_camera(camera)
        if smallpt_camera is not None:
            r.set_smallpt_camera(smallpt_camera)
        r.set_image(width, height, nsub)
        self.accum = torch.zeros((width * height * nsub * nsub, 4), dtype=torch.float32, device=self.device)
        r.set_accum_buffer(self.accum.data_ptr(), self.accum.numel() * 4)

    def step(self, seed: int, total_samples: int, flags: int = 0, resolve: bool = True):
        """One frame: clear, trace this rank's sample share, reduce, resolve on rank 0 (device-side)."""
        first, count = sample_range(total_samples, self.world, self.rank)
        self.accum.zero_()
        self.renderer.render(seed, first, count, flags)
        reduce_accum_(self.accum, 0)
        if resolve and self.rank == 0:
            return self.renderer.resolve_device()
        return None

    def close(self):
        self.renderer.close()
