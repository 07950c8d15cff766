#!/usr/bin/env python
"""bench.py -- the reference's headline metric (Mpaths/s, Mrays/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2..C5]

A STEP is one whole frame of the workload: zero the accumulation buffer, trace every
sample of every sub-pixel (this rank's share of the samples when N > 1), sum the
buffers onto rank 0 (NCCL, N > 1 only), resolve on rank 0.  Default workload =
BASELINE.json configs[2], the north-star scene: box_mirror_scene.hpp, 1920x1080,
4096 spp (1024 samples per 2x2 sub-pixel), strong scaling over N GPUs by splitting the
samples of every sub-pixel.

  value      whole-job Mpaths/s, scene/camera/accumulation buffer resident in HBM,
             device time from CUDA events on the launching stream, max over ranks
  e2e        same metric through the C ABI with HOST buffers every step: scene and
             camera uploaded from host memory, image read back into host memory
  roofline   FP32-issue roofline (this path is branchy FP32 ALU work, not HBM- or
             tensor-bound): achieved = paths/s x algorithmic flop/path (SURVEY.md 8d
             constants x this run's counters); peak = FFMA rate measured live on this
             GPU by ptb_measure_fp32_peak (MEASURED_PEAKS.json has no FP32 entry)
  cpu_baseline  the reference's own code (oracle/_ref, kind "reference") or the C
             restatement (kind "port") timed on this box's host cores, bounded sample

--impl reference times the reference's CPU implementation of the same workload on the
host cores and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (scene, width, height, total spp)            BASELINE.json configs[i]
    "C1": ("simple", 1024, 768, 16),
    "C2": ("box", 1024, 768, 1024),
    "C3": ("box_mirror", 1920, 1080, 4096),
    "C4": ("dof_glass", 3840, 2160, 4096),
    "C5": ("spheres10k", 1920, 1080, 1024),
    # the reference's stand-alone smallpt fork, sandbox/main.cpp: its own scene, 1024x768
    "SB": ("smallpt", 1024, 768, 4096),
}
SEED = 1

# algorithmic flop constants, SURVEY.md section 8(d)
FLOP_RAY_FIXED, FLOP_TEST, FLOP_ROOT, FLOP_ROOT2 = 5, 17, 3, 2
FLOP_SHADE, FLOP_RR, FLOP_DIFFUSE, FLOP_SPECULAR = 33, 4, 65, 18
FLOP_DIELECTRIC, FLOP_REFRACT, FLOP_SKY, FLOP_PRIMARY = 31, 22, 28, 68


def flop_per_path(st: dict, n_spheres: int) -> float:
    """Algorithmic flop per path from oracle statistics of the same scene (SURVEY 8d formula)."""
    p = st["paths"]
    inter = (FLOP_RAY_FIXED * st["rays"] + FLOP_TEST * st["sphere_tests"] + FLOP_ROOT * st["disc_nonneg"]
             + FLOP_ROOT2 * st["second_root"])
    hits = st["rays"] - st["misses"]
    refl = st["dielectric_reflect"]
    scatter = (FLOP_DIFFUSE * st["hit_diffuse"] + FLOP_SPECULAR * st["hit_specular"]
               + FLOP_DIELECTRIC * st["hit_dielectric"] + FLOP_SPECULAR * refl
               + FLOP_REFRACT * (st["hit_dielectric"] - refl))
    rr = FLOP_RR * (st["rr_draws"] - st["rr_kills"])
    return (inter + FLOP_SHADE * hits + rr + scatter + FLOP_SKY * st["misses"] + FLOP_PRIMARY * p) / p


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l.split(",") for l in open(self.path).read().strip().splitlines() if l.count(",") >= 8]
        os.unlink(self.path)
        if not rows:
            return out
        mhz = [float(r[1]) for r in rows]
        power = [float(r[3]) for r in rows]
        # "under load" = samples at or above half the peak power seen
        load = [m for m, p in zip(mhz, power) if p >= 0.5 * max(power)] or mhz
        out["sm_mhz"] = statistics.median(load)
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(power)
        out["samples"] = len(rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for k, nm in enumerate(names):
            if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows):
                out["reasons"].append(nm)
        return out


def cpu_reference_rate(scene, width, height, samps, kind_pref=("ref_stock", "port")):
    """Mpaths/s of the reference's CPU implementation on this box's host cores.
    Returns (value, kind, cores, seconds, oracle-stats or None)."""
    from oracle import Oracle, available
    from __graft_entry__ import load_package

    pkg = load_package()
    cores = os.cpu_count() or 1
    paths = width * height * 4 * samps
    if scene == "smallpt":
        spheres, cam8 = pkg.builtin_smallpt_scene()
        kind = "reference" if available("ref_sandbox") else "port"
        orc = Oracle("ref_sandbox" if kind == "reference" else "port")
        t0 = time.perf_counter()
        orc.sb_render(spheres, cam8, width, height, samps, mode=0, nthreads=cores)  # the program's own erand48 stream
        dt = time.perf_counter() - t0
        return paths / dt / 1e6, kind, cores, dt
    spheres, cfg = pkg.builtin_scene(scene, width, height)  # host-side scene layer only (no GPU)
    cam = pkg.camera_with_config(cfg)
    if "ref_stock" in kind_pref and available("ref_stock"):
        orc = Oracle("ref_stock")
        t0 = time.perf_counter()
        orc.mt_render(spheres, cam, width, height, samps, 2, seed_mode=0, nthreads=cores)
        dt = time.perf_counter() - t0
        return paths / dt / 1e6, "reference", cores, dt
    orc = Oracle("port")
    t0 = time.perf_counter()
    orc.render(spheres, cam, width, height, samps, 2, SEED, 0, nthreads=cores)
    dt = time.perf_counter() - t0
    return paths / dt / 1e6, "port", cores, dt


def oracle_statistics(scene, width, height):
    """Path statistics of the scene from the C restatement on a small sample (for the flop model)."""
    from oracle import Oracle
    from __graft_entry__ import load_package

    pkg = load_package()
    w, h = max(64, width // 8), max(48, height // 8)
    if scene == "smallpt":
        return None, 10  # the sandbox restatement keeps no statistics: flop model from the GPU counters
    spheres, cfg = pkg.builtin_scene(scene, w, h)
    cam = pkg.camera_with_config(cfg)
    orc = Oracle("port")
    orc.stats_reset()
    orc.render(spheres, cam, w, h, 2, 2, SEED, 0)
    return orc.stats(), len(spheres)


def run_reference(args, scene, width, height, spp):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    samps = 1 if scene != "spheres10k" else 1  # 4 spp per step: a bounded sample of the same frame
    w, h = (width, height) if scene != "spheres10k" else (width // 8, height // 8)
    for _ in range(args.warmup):
        cpu_reference_rate(scene, w, h, samps)
    secs, kind, cores = [], None, None
    for _ in range(args.steps):
        v, kind, cores, dt = cpu_reference_rate(scene, w, h, samps)
        secs.append(dt)
    paths = w * h * 4 * samps
    value = paths * len(secs) / sum(secs) / 1e6
    sample = f"{scene} {w}x{h}, {4 * samps} spp per step ({paths / 1e6:.1f} Mpaths), all host threads"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {scene} {width}x{height} @ {spp} spp", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout
    when NCCL_DEBUG asks for it): keep the real stdout for the result line and point fd 1 at stderr for everybody else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--spp", type=int, default=0, help="override the config's total samples per pixel")
    ap.add_argument("--variant", default="sorted", choices=["mega", "wavefront", "sorted"],
                    help="sorted = material-sorted megakernel (the product path); mega = in-place megakernel; wavefront")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least three warm-up steps

    scene, width, height, spp = CONFIGS[args.config]
    if args.spp:
        spp = args.spp
    samps = spp // 4  # main.cpp:206

    if args.impl == "reference":
        return run_reference(args, scene, width, height, spp)

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    pkg = load_package()  # raises if libptb200.so is missing: there is no fallback
    from cpu_path_tracing_b200.distributed import DistributedRenderer, sample_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    flags = pkg.PRECISION_FP32 | {"mega": pkg.VARIANT_MEGAKERNEL, "wavefront": pkg.VARIANT_WAVEFRONT,
                                  "sorted": pkg.VARIANT_MEGAKERNEL_SORTED}[args.variant]

    dr = DistributedRenderer(pkg, local, rank, world)
    if scene == "smallpt":
        spheres, cam = pkg.builtin_smallpt_scene()  # cam = cam8
        args.variant = "mega"  # the sandbox integrator (splitting glass) runs on the in-place megakernel only
        flags = pkg.PRECISION_FP32 | pkg.VARIANT_MEGAKERNEL | pkg.INTEGRATOR_SMALLPT
        dr.setup(spheres, None, width, height, 2, smallpt_camera=cam)
        set_camera = dr.renderer.set_smallpt_camera
    else:
        spheres, cfg = pkg.builtin_scene(scene, width, height)
        cam = pkg.camera_with_config(cfg)
        dr.setup(spheres, cam, width, height, 2)
        set_camera = dr.renderer.set_camera
    r = dr.renderer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement ------------------------------------------------------------
    for _ in range(args.warmup):
        dr.step(SEED, samps, flags)
    launches0 = r.stats().kernel_launches
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    for a, b in ev:
        a.record()
        dr.step(SEED, samps, flags)
        b.record()
        kernel_ms.append(r.stats().last_render_ms)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    launches = r.stats().kernel_launches - launches0
    st = r.stats()
    my_first, my_count = sample_range(samps, world, rank)
    # statistics are cleared only by ptb_clear; DistributedRenderer zeroes its tensor itself, so rays and
    # paths have accumulated over every step since setup on this rank
    counters = torch.tensor([st.rays, st.paths], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    rays_per_path = float(counters[0].item() / max(counters[1].item(), 1.0))

    paths_per_step = width * height * 4 * samps
    value = paths_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers -------------------------------------------------
    # every step: scene + camera from host memory -> device, render, reduce, image -> host memory
    host_img = np.empty((height, width, 3), dtype=np.float64)
    pinned = torch.empty((height, width, 3), dtype=torch.float64).pin_memory().numpy() if rank == 0 else None
    sph_bytes = np.ascontiguousarray(spheres).view(np.uint8)
    h2d = int(sph_bytes.nbytes + cam.nbytes)
    d2h = int(host_img.nbytes)

    def e2e_step():
        r.upload_scene(spheres)
        set_camera(cam)
        dr.step(SEED, samps, flags, resolve=False)
        if rank == 0:
            r.resolve_into(pinned)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_ev0, e2e_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ev0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_ev1.record()
    barrier()
    e2e_ms = torch.tensor([max(e2e_ev0.elapsed_time(e2e_ev1), 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * args.steps / (float(e2e_ms.item()) * 1e-3) / 1e6
    del t0

    if rank == 0:
        jit = r.jit_info()
        peak_tflops = r.measure_fp32_peak()
        ostats, n_spheres = oracle_statistics(scene, width, height)
        if ostats is None:
            # sandbox scene: same per-operation constants, event counts from the GPU counters of this run
            # (no sky; ~1/3 of the 10 sphere tests reach the root stage, measured on the src/ scenes)
            rpp = rays_per_path
            hits = {k: getattr(st, "hits_" + k) / max(st.paths, 1) for k in ("diffuse", "specular", "dielectric")}
            fpp = (rpp * (FLOP_RAY_FIXED + n_spheres * (FLOP_TEST + 0.66 * FLOP_ROOT + 0.33 * FLOP_ROOT2)) + FLOP_SHADE * rpp
                   + FLOP_DIFFUSE * hits["diffuse"] + FLOP_SPECULAR * hits["specular"]
                   + (FLOP_DIELECTRIC + FLOP_REFRACT) * hits["dielectric"] + FLOP_PRIMARY)
        else:
            fpp = flop_per_path(ostats, n_spheres)
        # the dominant kernel = the megakernel; its average launch duration from the library's own events
        k_ms = sum(kernel_ms) / len(kernel_ms)
        paths_per_launch = width * height * 4 * my_count
        achieved = paths_per_launch * fpp / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{args.config}_{args.variant}")
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.config}: {scene} {width}x{height} @ {spp} spp "
                            + (f"(BASELINE.json configs[{int(args.config[1]) - 1}])" if args.config[0] == "C" else "(sandbox/main.cpp, SURVEY 8 f-1)"),
                "variant": args.variant, "samples_per_subpixel": samps, "spheres": int(len(spheres)),
                "codegen": ("run-time compiled for this scene (NVRTC, %d ms once)" % jit["compile_ms"]) if jit["last_launch_jit"]
                else "precompiled",
                "partition": f"samples of every sub-pixel split over {world} GPU(s); NCCL sum-reduce to rank 0" if world > 1
                else "single GPU",
                "l2": "every step zeroes the accumulation buffer (133 MB at 1080p > 126 MB L2) and the kernel's inputs are "
                      "~1 KB of constants: compute-bound, L2 state does not matter",
                "seed": SEED,
            },
            "Mrays_per_s": value * rays_per_path, "rays_per_path": rays_per_path,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": traffic, "flop_per_path": fpp, "kernel_ms": k_ms,
                "peak_source": "FFMA loop measured live on this GPU (ptb_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.45",
            },
        }
        if scene == "spheres10k":
            # flop_per_path is the REFERENCE's algorithm (a 10 001-sphere linear scan per ray);
            # the kernel answers the same queries through a bounding-volume hierarchy (~10 box pairs + ~0.6 sphere tests
            # per ray), so "achieved" is an equivalent-work rate and may exceed the hardware peak -- not a pipe fraction
            line["roofline"]["frac"] = None
            line["roofline"]["equivalent_scan_frac"] = achieved / peak_tflops
            line["roofline"]["note"] = ("closest hit through a hierarchy (PTB_ACCEL_AUTO): algorithmic flops are the reference's "
                                        "linear scan, achieved/peak > 1 means work skipped, not a pipe utilisation")
        if not args.no_cpu_baseline:
            bw, bh = (width, height) if scene != "spheres10k" else (width // 8, height // 8)
            bs = 4 if scene != "spheres10k" else 1
            v, kind, cores, dt = cpu_reference_rate(scene, bw, bh, bs)
            line["cpu_baseline"] = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": kind,
                                    "sample": f"{scene} {bw}x{bh} at {4 * bs} spp ({bw * bh * 4 * bs / 1e6:.1f} Mpaths, {dt:.1f} s), "
                                              f"all {cores} host threads, OpenMP dynamic rows"}
        _emit(line)
    dr.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
