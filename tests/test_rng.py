"""The counter-based stream: CPU statement (oracle/ptb_rng.h) and device statement (csrc/ptb_rng.cuh)."""
import numpy as np
import pytest

MASK64 = (1 << 64) - 1


def mix64(z):
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & MASK64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & MASK64
    z ^= z >> 31
    return z


def py_draws(seed, slot, sample, n):
    """Pure-Python restatement of the stream definition in oracle/ptb_rng.h (small cases only)."""
    k = mix64((seed + 0x9E3779B97F4A7C15) & MASK64)
    h = mix64(k ^ ((slot << 32) | sample))
    state, inc = h & 0xFFFFFFFF, ((h >> 32) | 1) & 0xFFFFFFFF
    out = []
    for _ in range(n):
        old = state
        state = (old * 747796405 + inc) & 0xFFFFFFFF
        word = (((old >> ((old >> 28) + 4)) ^ old) * 277803737) & 0xFFFFFFFF
        r = (word >> 22) ^ word
        out.append((r >> 9) / 8388608.0)
    return out


def test_python_restatement_known_answers():
    # frozen known answers: any change of the stream definition must be deliberate
    d = py_draws(1, 0, 0, 4)
    assert d == [0.6672176122665405, 0.4307444095611572, 0.3898298740386963, 0.4154045581817627]
    assert py_draws(0xDEADBEEFCAFEF00D, 33177599, 4095, 3) == [0.9312283992767334, 0.7858893871307373, 0.12594234943389893]
    assert py_draws(1, 0, 0, 4) != py_draws(1, 0, 1, 4) != py_draws(1, 1, 0, 4) != py_draws(2, 0, 0, 4)
    # 23-bit values: exactly representable in binary32
    assert all(np.float32(v) == v for v in py_draws(77, 123456, 999, 64))


def test_oracle_jitter_uses_the_stream(oracle_port, pkg):
    """The first two uniforms of a sample are the sub-pixel jitter (main.cpp:187-188): recover them from the ray."""
    W, H = 64, 48
    sph, cfg = pkg.builtin_scene("simple", W, H)
    cfg = cfg.copy()
    cfg["aperture"] = 0.0  # no lens offset: direction = llc + X s + Y t - position, exactly
    cam = pkg.camera_with_config(cfg)
    xs, ys, sx, sy, ss = [np.array(v) for v in ([5, 17, 63], [3, 40, 47], [0, 1, 1], [1, 0, 1], [0, 7, 123456])]
    _, _, ray, _ = oracle_port.samples(sph, cam, W, H, 2, 42, xs, ys, sx, sy, ss)
    X, Y, llc, pos = cam["cam_x_axis"][0], cam["cam_y_axis"][0], cam["lower_left_corner"][0], cam["position"][0]
    for i in range(3):
        slot = ((int(ys[i]) * W + int(xs[i])) * 2 + int(sy[i])) * 2 + int(sx[i])
        u0, u1 = py_draws(42, slot, int(ss[i]), 2)
        s = (xs[i] + sx[i] * 0.5 + (0.0 + 0.5 * u0)) / W
        t = (ys[i] + sy[i] * 0.5 + (0.0 + 0.5 * u1)) / H
        d = llc + X * s + Y * t - pos
        assert np.allclose(ray[i, 3:], d, rtol=0, atol=1e-12)


def test_uniformity_and_independence():
    # cheap sanity on the python restatement: mean/variance of 1st draws across consecutive samples and slots
    a = np.array([py_draws(9, 0, s, 1)[0] for s in range(4000)])
    b = np.array([py_draws(9, s, 0, 1)[0] for s in range(4000)])
    for v in (a, b):
        assert abs(v.mean() - 0.5) < 0.02 and abs(v.var() - 1 / 12) < 0.01
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.05 and abs(np.corrcoef(a, b)[0, 1]) < 0.05


@pytest.mark.gpu
def test_device_stream_equals_cpu_statement(gpu):
    pkg = gpu
    rng = np.random.default_rng(0)
    n = 512
    slots = rng.integers(0, 1 << 25, n).astype(np.uint32)
    samples = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    slots[:3] = (0, 1, 0xFFFFFFFF)
    samples[:3] = (0, 0xFFFFFFFF, 5)
    for seed in (0, 1, 0xDEADBEEFCAFEF00D):
        with pkg.Renderer(0) as r:
            got = r.rng_draws(seed, slots, samples, 12)
        for i in range(0, n, 7):
            assert got[i].tolist() == py_draws(seed, int(slots[i]), int(samples[i]), 12)
