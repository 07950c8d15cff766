"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes over libptb200.so).  Built on demand, never faked."""
    from __graft_entry__ import PKG_DIR, load_package

    if not os.path.exists(os.path.join(PKG_DIR, "libptb200.so")):
        import subprocess

        subprocess.run(["make", "-s", "-C", PKG_DIR], check=True)
    return load_package()


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import Oracle

    return Oracle("port")


def _ref(kind):
    from oracle import Oracle, available

    if not available(kind):
        pytest.skip(f"oracle/_ref ({kind}) not built here (needs /root/reference at build time)")
    return Oracle(kind)


@pytest.fixture(scope="session")
def ref_ctr():
    return _ref("ref_ctr")


@pytest.fixture(scope="session")
def ref_stock():
    return _ref("ref_stock")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))

    return load


@pytest.fixture(scope="session")
def golden_scene(golden):
    """(spheres[n,88] u8, config[112] u8, camera[176] u8) exactly as the reference's builders produced them."""
    z = golden("scenes.npz")

    def get(name, w, h):
        k = f"{name}_{w}x{h}"
        return z[k + "_spheres"], z[k + "_config"], z[k + "_camera"]

    return get


@pytest.fixture(scope="session")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.skip("no CUDA device")
    return pkg


def make_renderer(pkg, spheres, camera, w, h, nsub=2, device=0):
    r = pkg.Renderer(device)
    r.upload_scene(spheres)
    r.set_camera(camera)
    r.set_image(w, h, nsub)
    return r


def classify_outliers(pkg, orc, r, spheres, camera, W, H, nsub, seed, xs, ys, sx, sy, ss, rad, orad, tol=1e-4, trail_len=64):
    """north_star (a) asks for per-sample radiance within 1e-4 of the reference.  The CUDA and the host arithmetic agree
    bit for bit on + - * / sqrt, so a path stays identical for as long as it only meets mirrors; sin / cos (diffuse_ray,
    main.cpp:47-52) and pow (Schlick, main.cpp:99-102) differ by an ulp between CUDA's and glibc's libm, and a chain of
    convex-mirror bounces afterwards amplifies that ulp geometrically.  This makes the explanation a CHECK: every sample
    outside `tol` must visit the same spheres as the oracle's path up to and including a diffuse or glass bounce --
    i.e. the deviation cannot have started on a mirror-only prefix.
    Returns (number of samples outside tol, number of those whose sphere sequence later diverges)."""
    import numpy as np

    rel = np.abs(rad - orad).max(axis=1) / np.maximum(np.abs(orad).max(axis=1), 1e-12)
    bad = np.nonzero(~(rel <= tol))[0]
    if bad.size == 0:
        return 0, 0
    sel = [np.asarray(a)[bad] for a in (xs, ys, sx, sy, ss)]
    gt = r.trace_paths(seed, *sel, trail_len=trail_len)
    ot = orc.trails(spheres, camera, W, H, nsub, seed, *sel, trail_len=trail_len)
    refl = np.asarray(spheres)["reflection"] if getattr(spheres, "dtype", None) is not None and spheres.dtype.names else \
        np.ascontiguousarray(spheres).view(np.uint8).reshape(-1, 88)[:, 80:84].copy().view(np.int32).ravel()
    diverged = 0
    for k in range(bad.size):
        same = gt[k] == ot[k]
        first_diff = int(np.argmin(same)) if not same.all() else trail_len
        prefix = ot[k][:first_diff]
        prefix = prefix[prefix >= 0]
        rough = np.isin(refl[prefix], (0, 2))  # diffuse or dielectric bounces inside the common prefix
        assert rough.any(), (f"sample {bad[k]} differs by {rel[bad[k]]:.3g} although both paths are identical and mirror-only "
                             f"up to depth {first_diff}: {gt[k][:8]} vs {ot[k][:8]}")
        diverged += first_diff < trail_len
    return int(bad.size), diverged
