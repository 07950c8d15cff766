"""cpu-path-tracing_b200 -- Python binding (ctypes) of the C ABI in include/ptb200.h.

This is plumbing for tests/ and bench.py: the product is libptb200.so (hand-written
sm_100a CUDA behind a C ABI) driven from C++ (host/pt.hpp, host/ptb_main.cpp).  There
is no CPU or PyTorch fallback anywhere in this package: if libptb200.so is missing,
importing fails; if no CUDA device is usable, Renderer() raises.

Load it with ``__graft_entry__.load_package()`` (the directory name has a hyphen).

Mirrors the host flow of the reference's main (/root/reference/src/main.cpp:199-248):

    scene  = pt::box_scene(w, h)                    -> builtin_scene("box_mirror", w, h)
    cam    = pt::camera::with_config(scene.camera)  -> camera_with_config(cfg)
    image  = vector<vec3>(w*h)                      -> Renderer.set_image(w, h, 2)
    executor.run(taskflow).wait()                   -> Renderer.render(seed, first, samps)
    (render_subpixel clamp/accumulate)              -> Renderer.resolve()
    PPM output                                      -> write_ppm(path, rgb)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTB200_LIB", os.path.join(HERE, "libptb200.so"))  # override: development A/B builds only

SPHERE_BYTES = 88
CAMERA_BYTES = 176
CAMERA_CONFIG_BYTES = 112

VARIANT_MEGAKERNEL = 0x0
VARIANT_WAVEFRONT = 0x1
VARIANT_MEGAKERNEL_SORTED = 0x2
ACCEL_AUTO = 0x0000
ACCEL_SCAN = 0x1000
CODEGEN_AUTO = 0x00000
CODEGEN_PRECOMPILED = 0x10000
PRECISION_FP32 = 0x00
PRECISION_FP64 = 0x10
INTEGRATOR_PT = 0x000
INTEGRATOR_SMALLPT = 0x100

DIFFUSE, SPECULAR, DIELECTRIC = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {HERE}` (or __graft_entry__.build()). "
        "There is no CPU fallback.")

# The multi-GPU entry points load libnccl.so.2 at run time.  In a Python process that may also import torch, it has to be
# the SAME file torch's libtorch_cuda.so is linked against (the pip package nvidia-nccl), not an older system copy: the
# dynamic loader keeps one object per soname, and torch fails to import over a libnccl that lacks its newer symbols.
if "PTB_NCCL_LIB" not in os.environ:
    try:
        import importlib.util as _ilu

        _spec = _ilu.find_spec("nvidia.nccl")
        for _d in (_spec.submodule_search_locations if _spec is not None else []):
            _cand = os.path.join(_d, "lib", "libnccl.so.2")
            if os.path.exists(_cand):
                os.environ["PTB_NCCL_LIB"] = _cand
                break
    except (ImportError, ValueError, AttributeError):
        pass

_lib = ctypes.CDLL(LIB_PATH)

_vp = ctypes.c_void_p
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_sz = ctypes.c_size_t


class PtbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ptb status {code}: {message}")
        self.code = code


class _Stats(ctypes.Structure):
    _fields_ = [
        ("paths", _u64), ("rays", _u64), ("last_render_ms", ctypes.c_double), ("total_render_ms", ctypes.c_double),
        ("kernel_launches", _u64), ("hits_diffuse", _u64), ("hits_specular", _u64), ("hits_dielectric", _u64),
        ("last_resolve_ms", ctypes.c_double),
    ]


@dataclass
class Stats:
    paths: int
    rays: int
    last_render_ms: float
    total_render_ms: float
    kernel_launches: int
    hits_diffuse: int
    hits_specular: int
    hits_dielectric: int
    last_resolve_ms: float = 0.0


def _sig(name, restype, *argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_ptb_abi_version = _sig("ptb_abi_version", ctypes.c_int)
_ptb_device_count = _sig("ptb_device_count", ctypes.c_int)
_ptb_last_error = _sig("ptb_last_error", ctypes.c_char_p, _vp)
_ptb_create = _sig("ptb_create", ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp))
_ptb_destroy = _sig("ptb_destroy", None, _vp)
_ptb_set_stream = _sig("ptb_set_stream", ctypes.c_int, _vp, _vp)
_ptb_reset_stream = _sig("ptb_reset_stream", ctypes.c_int, _vp)
_ptb_synchronize = _sig("ptb_synchronize", ctypes.c_int, _vp)
_ptb_upload_scene = _sig("ptb_upload_scene", ctypes.c_int, _vp, _vp, _sz, _sz)
_ptb_set_camera = _sig("ptb_set_camera", ctypes.c_int, _vp, _vp, _sz)
_ptb_set_smallpt_camera = _sig("ptb_set_smallpt_camera", ctypes.c_int, _vp, _vp)
_ptb_set_image = _sig("ptb_set_image", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int)
_ptb_clear = _sig("ptb_clear", ctypes.c_int, _vp)
_ptb_render = _sig("ptb_render", ctypes.c_int, _vp, _u64, _u32, _u32, _u32)
_ptb_resolve = _sig("ptb_resolve", ctypes.c_int, _vp, _vp)
_ptb_resolve_rgb8 = _sig("ptb_resolve_rgb8", ctypes.c_int, _vp, _vp)
_ptb_resolve_device = _sig("ptb_resolve_device", ctypes.c_int, _vp, ctypes.POINTER(_vp))
_ptb_measure_fp32_peak = _sig("ptb_measure_fp32_peak", ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_double))
_ptb_accum_buffer = _sig("ptb_accum_buffer", ctypes.c_int, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_sz))
_ptb_set_accum_buffer = _sig("ptb_set_accum_buffer", ctypes.c_int, _vp, _vp, _sz)
_ptb_download_accum = _sig("ptb_download_accum", ctypes.c_int, _vp, _vp, _sz)
_ptb_get_stats = _sig("ptb_get_stats", ctypes.c_int, _vp, ctypes.POINTER(_Stats))
_ptb_scene_layout = _sig("ptb_scene_layout", ctypes.c_int, _vp, _vp)
_ptb_jit_info = _sig("ptb_jit_info", ctypes.c_int, _vp, _vp)
_ptb_jit_last_error = _sig("ptb_jit_last_error", ctypes.c_char_p, _vp)
_ptb_trace_samples = _sig("ptb_trace_samples", ctypes.c_int, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _sz, _u32, _vp, _vp,
                          _vp, _vp)
_ptb_trace_paths = _sig("ptb_trace_paths", ctypes.c_int, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _sz, ctypes.c_int, _vp)
_ptb_rng_draws = _sig("ptb_rng_draws", ctypes.c_int, _vp, _u64, _vp, _vp, _sz, ctypes.c_int, _vp)
_ptb_camera_with_config = _sig("ptb_camera_with_config", ctypes.c_int, _vp, _vp)
_ptb_builtin_scene = _sig("ptb_builtin_scene", ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, _vp, _sz,
                          ctypes.POINTER(_sz), _vp)
_ptb_write_ppm = _sig("ptb_write_ppm", ctypes.c_int, ctypes.c_char_p, _vp, ctypes.c_int, ctypes.c_int)
_ptb_write_ppm_smallpt = _sig("ptb_write_ppm_smallpt", ctypes.c_int, ctypes.c_char_p, _vp, ctypes.c_int, ctypes.c_int)
_ptb_builtin_smallpt_scene = _sig("ptb_builtin_smallpt_scene", ctypes.c_int, _vp, _sz, ctypes.POINTER(_sz), _vp)
_ptb_upload_accum = _sig("ptb_upload_accum", ctypes.c_int, _vp, _vp, _sz)
_ptb_download_accum64 = _sig("ptb_download_accum64", ctypes.c_int, _vp, _vp, _sz)
_ptb_upload_accum64 = _sig("ptb_upload_accum64", ctypes.c_int, _vp, _vp, _sz)
_ptb_write_ppm_rgb8 = _sig("ptb_write_ppm_rgb8", ctypes.c_int, ctypes.c_char_p, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int)
_ptb_create_multi = _sig("ptb_create_multi", ctypes.c_int, _vp, ctypes.c_int, ctypes.POINTER(_vp))
_ptb_comm_unique_id = _sig("ptb_comm_unique_id", ctypes.c_int, _vp)
_ptb_comm_init_rank = _sig("ptb_comm_init_rank", ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int)
_ptb_comm_set_transport = _sig("ptb_comm_set_transport", ctypes.c_int, _vp, ctypes.c_int)
_ptb_comm_info = _sig("ptb_comm_info", ctypes.c_int, _vp, _vp)
_ptb_sample_share = _sig("ptb_sample_share", ctypes.c_int, _u32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_u32),
                         ctypes.POINTER(_u32))

TRANSPORT_AUTO, TRANSPORT_NCCL, TRANSPORT_PEER = 0, 1, 2
COMM_ID_BYTES = 128

# every symbol include/ptb200.h declares (tests check the header against this list and the .so)
EXPORTED_SYMBOLS = (
    "ptb_abi_version", "ptb_device_count", "ptb_last_error", "ptb_create", "ptb_destroy", "ptb_set_stream",
    "ptb_reset_stream", "ptb_synchronize", "ptb_upload_scene", "ptb_set_camera", "ptb_set_image", "ptb_clear", "ptb_render",
    "ptb_resolve", "ptb_resolve_rgb8", "ptb_resolve_device", "ptb_measure_fp32_peak", "ptb_accum_buffer", "ptb_set_accum_buffer", "ptb_download_accum",
    "ptb_get_stats", "ptb_scene_layout", "ptb_trace_samples", "ptb_rng_draws", "ptb_camera_with_config", "ptb_builtin_scene",
    "ptb_write_ppm", "ptb_write_ppm_smallpt", "ptb_builtin_smallpt_scene", "ptb_set_smallpt_camera",
    "ptb_jit_info", "ptb_jit_last_error",
    "ptb_upload_accum", "ptb_download_accum64", "ptb_upload_accum64", "ptb_write_ppm_rgb8",
    "ptb_create_multi", "ptb_comm_unique_id", "ptb_comm_init_rank", "ptb_comm_set_transport", "ptb_comm_info",
    "ptb_sample_share", "ptb_trace_paths",
)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def abi_version() -> int:
    return _ptb_abi_version()


def device_count() -> int:
    return _ptb_device_count()


# ---- host-side scene layer (no GPU needed) ---------------------------------------------------
SPHERE_DTYPE = np.dtype([
    ("radius", "<f8"), ("position", "<f8", 3), ("emission", "<f8", 3), ("color", "<f8", 3), ("reflection", "<i4"),
    ("pad", "<i4"),
])
assert SPHERE_DTYPE.itemsize == SPHERE_BYTES
CAMERA_CONFIG_DTYPE = np.dtype([
    ("position", "<f8", 3), ("direction", "<f8", 3), ("up", "<f8", 3), ("aspect_ratio", "<f8"),
    ("vertical_fov_radians", "<f8"), ("focal_length", "<f8"), ("aperture", "<f8"), ("focus_distance", "<f8"),
])
assert CAMERA_CONFIG_DTYPE.itemsize == CAMERA_CONFIG_BYTES
CAMERA_DTYPE = np.dtype([
    ("position", "<f8", 3), ("lower_left_corner", "<f8", 3), ("cam_x_axis", "<f8", 3), ("cam_y_axis", "<f8", 3),
    ("u", "<f8", 3), ("v", "<f8", 3), ("w", "<f8", 3), ("lens_radius", "<f8"),
])
assert CAMERA_DTYPE.itemsize == CAMERA_BYTES

BUILTIN_SCENES = ("simple", "box", "box_mirror", "dof_glass", "spheres10k")


def builtin_scene(name: str, width: int, height: int):
    """(spheres: SPHERE_DTYPE[n], camera_config: CAMERA_CONFIG_DTYPE[1]) of a built-in scene."""
    n = _sz(0)
    rc = _ptb_builtin_scene(name.encode(), width, height, None, 0, ctypes.byref(n), None)
    if rc != 0:
        raise PtbError(rc, f"unknown built-in scene {name!r} or bad size")
    spheres = np.zeros(n.value, dtype=SPHERE_DTYPE)
    cfg = np.zeros(1, dtype=CAMERA_CONFIG_DTYPE)
    rc = _ptb_builtin_scene(name.encode(), width, height, _ptr(spheres), n.value, ctypes.byref(n), _ptr(cfg))
    if rc != 0:
        raise PtbError(rc, "ptb_builtin_scene")
    return spheres, cfg


def builtin_smallpt_scene():
    """(spheres: SPHERE_DTYPE[10], cam8: float64[8]) of sandbox/main.cpp."""
    n = _sz(0)
    _ptb_builtin_smallpt_scene(None, 0, ctypes.byref(n), None)
    spheres = np.zeros(n.value, dtype=SPHERE_DTYPE)
    cam8 = np.zeros(8, dtype=np.float64)
    rc = _ptb_builtin_smallpt_scene(_ptr(spheres), n.value, ctypes.byref(n), _ptr(cam8))
    if rc != 0:
        raise PtbError(rc, "ptb_builtin_smallpt_scene")
    return spheres, cam8


def camera_with_config(cfg: np.ndarray) -> np.ndarray:
    cfg = np.ascontiguousarray(cfg)
    assert cfg.nbytes == CAMERA_CONFIG_BYTES
    cam = np.zeros(1, dtype=CAMERA_DTYPE)
    rc = _ptb_camera_with_config(_ptr(cfg), _ptr(cam))
    if rc != 0:
        raise PtbError(rc, "ptb_camera_with_config")
    return cam


def write_ppm(path: str, rgb: np.ndarray, smallpt: bool = False) -> None:
    rgb = np.ascontiguousarray(rgb, dtype=np.float64)
    h, w, c = rgb.shape
    assert c == 3
    rc = (_ptb_write_ppm_smallpt if smallpt else _ptb_write_ppm)(os.fsencode(path), _ptr(rgb), w, h)
    if rc != 0:
        raise PtbError(rc, f"ptb_write_ppm({path})")


def write_ppm_rgb8(path: str, rgb8: np.ndarray, binary: bool = True) -> None:
    """The 8-bit image of Renderer.resolve_rgb8 as "P6" (binary) or the reference's "P3" token layout."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, c = rgb8.shape
    assert c == 3
    rc = _ptb_write_ppm_rgb8(os.fsencode(path), _ptr(rgb8), w, h, 1 if binary else 0)
    if rc != 0:
        raise PtbError(rc, f"ptb_write_ppm_rgb8({path})")


def sample_share(total: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """(first, count): the contiguous share of `total` samples that member `rank` of `n_ranks` traces."""
    first, count = _u32(0), _u32(0)
    rc = _ptb_sample_share(total, n_ranks, rank, ctypes.byref(first), ctypes.byref(count))
    if rc != 0:
        raise PtbError(rc, "ptb_sample_share: bad (total, n_ranks, rank)")
    return first.value, count.value


def comm_unique_id() -> bytes:
    """128 opaque bytes made on rank 0 of a one-process-per-GPU job; the launcher's side channel carries them."""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = _ptb_comm_unique_id(buf)
    if rc != 0:
        raise PtbError(rc, "ptb_comm_unique_id: libnccl.so.2 could not be loaded")
    return buf.raw


# ---- the GPU renderer ------------------------------------------------------------------------------
class Renderer:
    """One context = one GPU.  Thin, 1:1 over the C ABI."""

    def __init__(self, device=0):
        """device: a GPU index, or a list of them -- then ONE context spans those GPUs (ptb_create_multi): every call
        below applies to all of them, render() splits the samples, resolve*() sum across the GPUs."""
        self._ctx = _vp(None)
        if isinstance(device, (list, tuple)):
            devs = (ctypes.c_int * len(device))(*device)
            rc = _ptb_create_multi(devs, len(device), ctypes.byref(self._ctx))
        else:
            rc = _ptb_create(device, ctypes.byref(self._ctx))
        if rc != 0:
            msg = _ptb_last_error(None).decode()
            self._ctx = _vp(None)
            raise PtbError(rc, msg or "bad device list")
        self.device = device
        self.width = self.height = self.nsub = 0

    # -- lifetime
    def close(self):
        if self._ctx:
            _ptb_destroy(self._ctx)
            self._ctx = _vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PtbError(rc, _ptb_last_error(self._ctx).decode())

    # -- inputs
    def upload_scene(self, spheres: np.ndarray, stride: int | None = None):
        spheres = np.ascontiguousarray(spheres)
        if spheres.dtype == np.uint8:
            stride = stride or SPHERE_BYTES
            count = spheres.size // stride
        else:
            stride = stride or spheres.dtype.itemsize
            count = spheres.shape[0]
        self._check(_ptb_upload_scene(self._ctx, _ptr(spheres), count, stride))

    def set_camera(self, camera: np.ndarray):
        camera = np.ascontiguousarray(camera)
        self._check(_ptb_set_camera(self._ctx, _ptr(camera), camera.nbytes))

    def set_smallpt_camera(self, cam8: np.ndarray):
        cam8 = np.ascontiguousarray(cam8, dtype=np.float64)
        assert cam8.size == 8
        self._check(_ptb_set_smallpt_camera(self._ctx, _ptr(cam8)))

    def set_image(self, width: int, height: int, num_subpixels: int = 2):
        self._check(_ptb_set_image(self._ctx, width, height, num_subpixels))
        self.width, self.height, self.nsub = width, height, num_subpixels

    def set_stream(self, cuda_stream: int):
        """Run on this cudaStream_t handle; 0 is the legacy default stream (torch's default current stream)."""
        self._check(_ptb_set_stream(self._ctx, _vp(cuda_stream) if cuda_stream else None))

    def reset_stream(self):
        self._check(_ptb_reset_stream(self._ctx))

    def synchronize(self):
        self._check(_ptb_synchronize(self._ctx))

    def clear(self):
        self._check(_ptb_clear(self._ctx))

    # -- hot path
    def render(self, seed: int, first_sample: int, samples_per_subpixel: int, flags: int = 0):
        self._check(_ptb_render(self._ctx, seed, first_sample, samples_per_subpixel, flags))

    def resolve(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.float64)
        self._check(_ptb_resolve(self._ctx, _ptr(out)))
        return out

    def resolve_into(self, out: np.ndarray):
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.height * self.width * 3
        self._check(_ptb_resolve(self._ctx, _ptr(out)))

    def resolve_rgb8_into(self, out: np.ndarray):
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self.height * self.width * 3
        self._check(_ptb_resolve_rgb8(self._ctx, _ptr(out)))

    def resolve_rgb8(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        self._check(_ptb_resolve_rgb8(self._ctx, _ptr(out)))
        return out

    def resolve_device(self) -> int:
        """Resolve without leaving the GPU; returns the device address of the W*H*3 FP64 image."""
        p = _vp(None)
        self._check(_ptb_resolve_device(self._ctx, ctypes.byref(p)))
        return p.value

    def measure_fp32_peak(self) -> float:
        """Measured FFMA rate of this GPU in TFLOP/s (the FP32 roofline denominator)."""
        v = ctypes.c_double(0.0)
        self._check(_ptb_measure_fp32_peak(self._ctx, ctypes.byref(v)))
        return v.value

    # -- multi-GPU plumbing
    def accum_buffer(self):
        p, n = _vp(None), _sz(0)
        self._check(_ptb_accum_buffer(self._ctx, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def set_accum_buffer(self, device_ptr: int | None, nbytes: int = 0):
        self._check(_ptb_set_accum_buffer(self._ctx, _vp(device_ptr) if device_ptr else None, nbytes))

    def download_accum(self) -> np.ndarray:
        out = np.empty((self.height * self.width * self.nsub * self.nsub, 4), dtype=np.float32)
        self._check(_ptb_download_accum(self._ctx, _ptr(out), out.size))
        return out

    def upload_accum(self, accum: np.ndarray):
        """Restore a checkpoint made by download_accum (same image geometry); continue with render(seed, first = samples held)."""
        accum = np.ascontiguousarray(accum, dtype=np.float32)
        self._check(_ptb_upload_accum(self._ctx, _ptr(accum), accum.size))

    def download_accum64(self) -> np.ndarray:
        out = np.empty((self.height * self.width * self.nsub * self.nsub, 4), dtype=np.float64)
        self._check(_ptb_download_accum64(self._ctx, _ptr(out), out.size))
        return out

    def upload_accum64(self, accum: np.ndarray):
        accum = np.ascontiguousarray(accum, dtype=np.float64)
        self._check(_ptb_upload_accum64(self._ctx, _ptr(accum), accum.size))

    # -- several GPUs (one process per GPU: comm_init_rank; one process for all: Renderer([0, 1, ...]))
    def comm_init_rank(self, unique_id: bytes, n_ranks: int, rank: int):
        """Join a job of n_ranks processes (collective).  Afterwards render() traces this rank's share of the samples and
        resolve*() are collective: rank 0 receives the image."""
        assert len(unique_id) == COMM_ID_BYTES
        self._check(_ptb_comm_init_rank(self._ctx, ctypes.create_string_buffer(unique_id, COMM_ID_BYTES), n_ranks, rank))

    def comm_set_transport(self, transport: int):
        self._check(_ptb_comm_set_transport(self._ctx, transport))

    def comm_info(self) -> dict:
        out = np.zeros(6, dtype=np.int32)
        self._check(_ptb_comm_info(self._ctx, _ptr(out)))
        d = dict(zip(("n_gpus", "rank", "last_transport", "peer_mapped", "nccl_version", "mode"), [int(v) for v in out]))
        d["last_transport"] = {0: "none", 1: "nccl", 2: "peer"}[d["last_transport"]]
        d["mode"] = {0: "single", 1: "one process", 2: "process per GPU"}[d["mode"]]
        return d

    def resolve_collective(self):
        """resolve() for ranks other than 0 of a one-process-per-GPU job: takes part, receives nothing."""
        self._check(_ptb_resolve(self._ctx, None))

    # -- introspection
    def stats(self) -> Stats:
        s = _Stats()
        self._check(_ptb_get_stats(self._ctx, ctypes.byref(s)))
        return Stats(*[getattr(s, f[0]) for f in _Stats._fields_])

    def scene_layout(self) -> dict:
        out = np.zeros(10, dtype=np.int32)
        self._check(_ptb_scene_layout(self._ctx, _ptr(out)))
        keys = ("small_near", "small_both", "big_near", "big_both", "big_x", "big_y", "big_z", "bits", "fits_const",
                "specialised")
        d = dict(zip(keys, [int(v) for v in out]))
        bits = d.pop("bits")
        d["uniform_k"], d["embed_ok"] = bits & 1, (bits >> 1) & 1
        return d

    def jit_info(self) -> dict:
        """Run-time code generation of the sorted megakernel (PTB_CODEGEN_AUTO): see ptb_jit_info in ptb200.h."""
        out = np.zeros(5, dtype=np.int32)
        self._check(_ptb_jit_info(self._ctx, _ptr(out)))
        d = dict(zip(("available", "compiled", "failures", "last_launch_jit", "compile_ms"), [int(v) for v in out]))
        d["last_error"] = (_ptb_jit_last_error(self._ctx) or b"").decode()
        return d

    def trace_samples(self, seed, xs, ys, sxs, sys_, samples, flags=PRECISION_FP64):
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, sxs, sys_, samples)]
        count = arrs[0].size
        hit = np.zeros(count, dtype=np.int32)
        rad = np.zeros((count, 3), dtype=np.float64)
        ray = np.zeros((count, 6), dtype=np.float64)
        draws = np.zeros(count, dtype=np.uint32)
        self._check(_ptb_trace_samples(self._ctx, seed, *[_ptr(a) for a in arrs], count, flags, _ptr(hit), _ptr(rad),
                                       _ptr(ray), _ptr(draws)))
        return hit, rad, ray, draws

    def trace_paths(self, seed, xs, ys, sxs, sys_, samples, trail_len=32) -> np.ndarray:
        """[count, trail_len] sphere index hit at each depth (FP64 mode): -1 sky, -2 path over."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, sxs, sys_, samples)]
        count = arrs[0].size
        trail = np.zeros((count, trail_len), dtype=np.int32)
        self._check(_ptb_trace_paths(self._ctx, seed, *[_ptr(a) for a in arrs], count, trail_len, _ptr(trail)))
        return trail

    def rng_draws(self, seed, slots, samples, n_draws) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        out = np.zeros((slots.size, n_draws), dtype=np.float64)
        self._check(_ptb_rng_draws(self._ctx, seed, _ptr(slots), _ptr(samples), slots.size, n_draws, _ptr(out)))
        return out


def render_scene(name: str, width: int, height: int, spp: int, seed: int = 1, device: int = 0, flags: int = 0):
    """The whole of the reference's main() for a built-in scene: returns the W*H*3 FP64 image."""
    spheres, cfg = builtin_scene(name, width, height)
    cam = camera_with_config(cfg)
    with Renderer(device) as r:
        r.upload_scene(spheres)
        r.set_camera(cam)
        r.set_image(width, height, 2)
        r.render(seed, 0, spp // 4, flags)  # main.cpp:206: samps = spp / (2*2)
        return r.resolve()
