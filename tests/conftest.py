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
