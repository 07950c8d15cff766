"""ctypes loader for the CHECKERS under oracle/ (test infrastructure, never the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Two families of libraries share one call surface:

  * ``Oracle("port")``      -> oracle/libpt_oracle.so, the plain-C restatement
                               (pt_oracle.c), buildable anywhere gcc exists.
  * ``Oracle("ref_ctr")``   -> oracle/_ref/libptref_ctr.so, the reference's own
                               object code with the counter stream injected.
  * ``Oracle("ref_stock")`` -> oracle/_ref/libptref_stock.so, the reference's own
                               object code with its stock mt19937 stream.
  * ``Oracle("ref_sandbox")`` -> oracle/_ref/libsbref.so, sandbox/main.cpp's own object
                               code (the stand-alone smallpt fork).

oracle/_ref/*.so can only be BUILT where /root/reference exists; the built files
travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SPHERE_BYTES = 88
CAMERA_BYTES = 176
CAMERA_CONFIG_BYTES = 112

STAT_NAMES = (
    "paths", "rays", "sphere_tests", "disc_nonneg", "second_root", "hit_diffuse", "hit_specular",
    "hit_dielectric", "dielectric_reflect", "rr_draws", "rr_kills", "misses", "depth_limit", "draws",
)

_PATHS = {
    "port": os.path.join(HERE, "libpt_oracle.so"),
    "ref_ctr": os.path.join(HERE, "_ref", "libptref_ctr.so"),
    "ref_stock": os.path.join(HERE, "_ref", "libptref_stock.so"),
    "ref_sandbox": os.path.join(HERE, "_ref", "libsbref.so"),
}

_vp = ctypes.c_void_p
_u32p = ctypes.POINTER(ctypes.c_uint32)


def build(ref: bool = True) -> None:
    """Compile the C restatement and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Oracle:
    def __init__(self, kind: str = "port"):
        if kind == "port" and not os.path.exists(_PATHS[kind]):
            build(ref=False)
        self.kind = kind
        self.lib = ctypes.CDLL(_PATHS[kind])
        self.prefix = "orc_" if kind == "port" else "ptref_"

    # ---- helpers -----------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- scene data (reference builds only) -----------------------------------------
    def scene(self, name: str, w: int, h: int):
        """(spheres[n,88] uint8, camera_config[112] uint8, camera[176] uint8) from the reference's builders."""
        assert self.kind != "port", "scene builders live in the reference builds"
        sph = np.zeros((64, SPHERE_BYTES), dtype=np.uint8)
        cfg = np.zeros(CAMERA_CONFIG_BYTES, dtype=np.uint8)
        cam = np.zeros(CAMERA_BYTES, dtype=np.uint8)
        n = ctypes.c_int(0)
        rc = self.lib.ptref_scene(name.encode(), int(w), int(h), _ptr(sph), 64, ctypes.byref(n), _ptr(cfg), _ptr(cam))
        if rc != 0:
            raise ValueError(f"ptref_scene({name!r}) -> {rc}")
        return sph[: n.value].copy(), cfg, cam

    def camera_with_config(self, cfg: np.ndarray) -> np.ndarray:
        cam = np.zeros(CAMERA_BYTES, dtype=np.uint8)
        cfg = np.ascontiguousarray(cfg).view(np.uint8)
        assert cfg.size == CAMERA_CONFIG_BYTES
        self._fn("camera_with_config")(_ptr(cfg), _ptr(cam))
        return cam

    def color_to_int(self, v: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(v.size, dtype=np.int32)
        self._fn("color_to_int")(_ptr(v), int(v.size), _ptr(out))
        return out.reshape(v.shape)

    def intersect(self, spheres, origin, direction):
        spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        t = ctypes.c_double(0.0)
        fn = self._fn("intersect")
        fn.restype = ctypes.c_int
        idx = fn(_ptr(spheres), len(spheres), _ptr(o), _ptr(d), ctypes.byref(t))
        return idx, t.value

    # ---- counter-stream entry points (port, ref_ctr) ------------------------------------
    def samples(self, spheres, camera, width, height, nsub, seed, xs, ys, sxs, sys_, samples):
        assert self.kind in ("port", "ref_ctr")
        spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
        camera = np.ascontiguousarray(camera).view(np.uint8)
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, sxs, sys_, samples)]
        count = arrs[0].size
        hit = np.zeros(count, dtype=np.int32)
        rad = np.zeros((count, 3), dtype=np.float64)
        ray = np.zeros((count, 6), dtype=np.float64)
        draws = np.zeros(count, dtype=np.uint64)
        self._fn("ctr_samples" if self.kind == "ref_ctr" else "samples")(
            _ptr(spheres), len(spheres), _ptr(camera), int(width), int(height), int(nsub), ctypes.c_uint64(seed),
            *[_ptr(a) for a in arrs], int(count), _ptr(hit), _ptr(rad), _ptr(ray), _ptr(draws))
        return hit, rad, ray, draws

    def trails(self, spheres, camera, width, height, nsub, seed, xs, ys, sxs, sys_, samples, trail_len=32):
        """[count, trail_len] sphere index at each depth of each sample's path (-1 sky, -2 path over); port only."""
        assert self.kind == "port"
        spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
        camera = np.ascontiguousarray(camera).view(np.uint8)
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, sxs, sys_, samples)]
        count = arrs[0].size
        trail = np.zeros((count, trail_len), dtype=np.int32)
        self.lib.orc_trails(_ptr(spheres), len(spheres), _ptr(camera), int(width), int(height), int(nsub),
                            ctypes.c_uint64(seed), *[_ptr(a) for a in arrs], int(count), int(trail_len), _ptr(trail))
        return trail

    def render(self, spheres, camera, width, height, samps, nsub=2, seed=1, first_sample=0, nthreads=0,
               want_sums=False):
        assert self.kind in ("port", "ref_ctr")
        spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
        camera = np.ascontiguousarray(camera).view(np.uint8)
        img = np.zeros((height, width, 3), dtype=np.float64)
        sums = np.zeros((height * width * nsub * nsub, 3), dtype=np.float64) if want_sums else None
        self._fn("ctr_render" if self.kind == "ref_ctr" else "render")(
            _ptr(spheres), len(spheres), _ptr(camera), int(width), int(height), int(samps), int(nsub),
            ctypes.c_uint64(seed), ctypes.c_uint32(first_sample), _ptr(img), _ptr(sums), int(nthreads))
        return (img, sums) if want_sums else img

    # ---- stock mt19937 stream (port: reproducible modes; ref_stock: all modes) ---------------
    def mt_render(self, spheres, camera, width, height, samps, nsub=2, seed_mode=1, y0=0, y1=None, nthreads=0):
        spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
        camera = np.ascontiguousarray(camera).view(np.uint8)
        y1 = height if y1 is None else y1
        img = np.zeros((height, width, 3), dtype=np.float64)
        if self.kind == "port":
            assert seed_mode == 1, "the C restatement only has the reproducible stock mode"
            self.lib.orc_mt_render(_ptr(spheres), len(spheres), _ptr(camera), int(width), int(height), int(samps),
                                   int(nsub), 1, None, int(y0), int(y1), _ptr(img), int(nthreads))
        else:
            assert self.kind == "ref_stock"
            self.lib.ptref_stock_render(_ptr(spheres), len(spheres), _ptr(camera), int(width), int(height),
                                        int(samps), int(nsub), int(seed_mode), int(y0), int(y1), _ptr(img),
                                        int(nthreads))
        return img

    @staticmethod
    def reference_main(spp: int, workdir: str) -> int:
        """Run the reference PROGRAM (src/main.cpp:199-248, built unmodified as
        oracle/_ref/cpu_path_tracer) in workdir -> workdir/image.ppm."""
        exe = os.path.join(HERE, "_ref", "cpu_path_tracer")
        return subprocess.run([exe, str(int(spp))], cwd=workdir, stderr=subprocess.DEVNULL).returncode

    # ---- sandbox/main.cpp, the stand-alone smallpt fork (port, ref_sandbox) -------------------------------
    def sb_scene(self):
        """(spheres[n,88] uint8, cam8 float64[8]) exactly as sandbox/main.cpp defines them."""
        assert self.kind == "ref_sandbox"
        sph = np.zeros((32, SPHERE_BYTES), dtype=np.uint8)
        n = self.lib.sbref_scene(_ptr(sph), 32)
        cam = np.zeros(8, dtype=np.float64)
        self.lib.sbref_camera(_ptr(cam))
        return sph[:n].copy(), cam

    def sb_render(self, spheres, cam8, width, height, samps, mode=1, seed=1, first_sample=0, y0=0, y1=None, nthreads=0):
        """mode 0: the program's erand48 stream (Xi = {0,0,y^3} per row); mode 1: counter stream."""
        assert self.kind in ("port", "ref_sandbox")
        y1 = height if y1 is None else y1
        img = np.zeros((height, width, 3), dtype=np.float64)
        args = [int(width), int(height), int(samps), int(mode), ctypes.c_uint64(seed), ctypes.c_uint32(first_sample),
                int(y0), int(y1), _ptr(img), int(nthreads)]
        if self.kind == "port":
            spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
            cam8 = np.ascontiguousarray(cam8, dtype=np.float64)
            self.lib.orc_sb_render(_ptr(spheres), len(spheres), _ptr(cam8), *args)
        else:
            self.lib.sbref_render(*args)
        return img

    def sb_samples(self, spheres, cam8, width, height, seed, xs, ys, sxs, sys_, samples):
        assert self.kind in ("port", "ref_sandbox")
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, sxs, sys_, samples)]
        count = arrs[0].size
        hit = np.zeros(count, dtype=np.int32)
        rad = np.zeros((count, 3), dtype=np.float64)
        ray = np.zeros((count, 6), dtype=np.float64)
        draws = np.zeros(count, dtype=np.uint64)
        args = [int(width), int(height), ctypes.c_uint64(seed), *[_ptr(a) for a in arrs], int(count), _ptr(hit), _ptr(rad),
                _ptr(ray), _ptr(draws)]
        if self.kind == "port":
            spheres = np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, SPHERE_BYTES)
            cam8 = np.ascontiguousarray(cam8, dtype=np.float64)
            self.lib.orc_sb_samples(_ptr(spheres), len(spheres), _ptr(cam8), *args)
        else:
            self.lib.sbref_samples(*args)
        return hit, rad, ray, draws

    def sb_to_int(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(v.size, dtype=np.int32)
        (self.lib.orc_sb_to_int if self.kind == "port" else self.lib.sbref_to_int)(_ptr(v), int(v.size), _ptr(out))
        return out.reshape(v.shape)

    @staticmethod
    def sandbox_program(spp: int, workdir: str) -> int:
        """Run the sandbox PROGRAM (oracle/_ref/smallpt, built as sandbox/run.sh:3) -> workdir/image.ppm."""
        exe = os.path.join(HERE, "_ref", "smallpt")
        return subprocess.run([exe, str(int(spp))], cwd=workdir, stderr=subprocess.DEVNULL).returncode

    # ---- statistics (port only) -------------------------------------------------------
    def stats_reset(self):
        assert self.kind == "port"
        self.lib.orc_stats_reset()

    def stats(self) -> dict:
        assert self.kind == "port"
        out = (ctypes.c_uint64 * len(STAT_NAMES))()
        self.lib.orc_stats_get(out)
        return dict(zip(STAT_NAMES, [int(v) for v in out]))
