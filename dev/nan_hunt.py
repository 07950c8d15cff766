"""Find samples whose FP32 radiance is non-finite and report the bounce that broke (needs the PTB_DEBUG_NAN build):
   PTB200_LIB=cpu-path-tracing_b200/build/libptb200_nan.so python dev/nan_hunt.py spheres10k 64 36 3"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
name, W, H, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sph, cfg = pkg.builtin_scene(name, W, H)
cam = pkg.camera_with_config(cfg)
n = W * H * 4 * S
idx = np.arange(n)
s = idx % S; slot = idx // S
sx = slot & 1; sy = (slot >> 1) & 1; pix = slot >> 2
y = pix // W; x = pix % W
with pkg.Renderer(0) as r:
    r.upload_scene(sph); r.set_camera(cam); r.set_image(W, H, 2)
    hit, rad, ray, draws = r.trace_samples(23, x, y, sx, sy, s, pkg.PRECISION_FP32)
    bad = np.nonzero(draws >= 1000)[0]
    print("samples", n, "flagged", bad.size, "non-finite radiance", int((~np.isfinite(rad)).any(axis=1).sum()))
    sp = np.frombuffer(sph.tobytes(), dtype=np.float64).reshape(-1, 11)
    for i in bad[:6]:
        print(f" sample x={x[i]} y={y[i]} sx={sx[i]} sy={sy[i]} s={s[i]}: bounce {rad[i,0]:.0f} hit listpos {rad[i,1]:.0f} t={rad[i,2]!r} prev listpos {int(draws[i])-1001}")
        print("   origin", ray[i, :3], "dir", ray[i, 3:], "|d| before", np.linalg.norm(ray[i, 3:]), "|d|^2 after", np.array([hit[i]], dtype=np.int32).view(np.float32)[0])
        lp = int(rad[i, 1])
        o, d = ray[i, :3], ray[i, 3:]
        P = o + rad[i, 2] * d
        print("   hit point", P)
    h64, rad64, _, _ = r.trace_samples(23, x[bad[:6]], y[bad[:6]], sx[bad[:6]], sy[bad[:6]], s[bad[:6]], pkg.PRECISION_FP64)
    print(" fp64 radiance of those:", rad64)
