"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (this container): it drives oracle/_ref/*.so,
which are the reference's own unmodified sources compiled by oracle/Makefile
(`make -C oracle ref`).  The fixtures are what travels to the GPU box.

    python tests/golden/make_golden.py

Fixtures
  scenes.npz        pt::simple_scene / box_scene (both headers) sphere records, camera_config
                    and pt::camera::with_config output, at 1024x768 and 1920x1080
  samples_<s>.npz   per-sample (primary hit id, radiance, camera ray, draw count) from the
                    reference's intersect/radiance/get_ray driven by the injected counter stream
  mtrows_<s>.npz    image rows rendered by the reference's render_subpixel with its STOCK
                    mt19937 stream seeded 0 (the only reproducible seed, random_state.cpp:5)
  refmain_rows.npz  rows y%64==0 of image.ppm written by the reference PROGRAM (main.cpp:199-248,
                    1024x768 box_mirror, 4 spp): bit-reproducible because (unsigned short)(y^3)==0
  image_<s>.npz     small whole images + un-clamped per-sub-pixel sums with the counter stream
  sandbox.npz       sandbox/main.cpp (stand-alone smallpt): its scene and camera constants, per-sample
                    vectors from its own radiance() with the counter stream, small images in both streams,
                    and 16 rows of the image.ppm its PROGRAM writes at 4 spp (deterministic: erand48)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

SCENES = ("simple", "box", "box_mirror")


def main():
    stock = Oracle("ref_stock")
    ctr = Oracle("ref_ctr")

    out = {}
    for name in SCENES:
        for (w, h) in ((1024, 768), (1920, 1080), (256, 192)):
            sph, cfg, cam = stock.scene(name, w, h)
            out[f"{name}_{w}x{h}_spheres"] = sph
            out[f"{name}_{w}x{h}_config"] = cfg
            out[f"{name}_{w}x{h}_camera"] = cam
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **out)

    rng = np.random.default_rng(20261018)
    W, H, NS, SEED, N = 256, 192, 2, 7, 4000
    for name in SCENES:
        sph, cfg, cam = stock.scene(name, W, H)
        xs, ys = rng.integers(0, W, N), rng.integers(0, H, N)
        sx, sy = rng.integers(0, NS, N), rng.integers(0, NS, N)
        ss = rng.integers(0, 1 << 24, N)
        hit, rad, ray, draws = ctr.samples(sph, cam, W, H, NS, SEED, xs, ys, sx, sy, ss)
        np.savez_compressed(os.path.join(HERE, f"samples_{name}.npz"), width=W, height=H, nsub=NS, seed=SEED,
                            x=xs.astype(np.uint32), y=ys.astype(np.uint32), sx=sx.astype(np.uint32),
                            sy=sy.astype(np.uint32), sample=ss.astype(np.uint32), hit=hit, radiance=rad, ray=ray,
                            draws=draws.astype(np.uint32))

        # stock mt19937{0} stream, two rows of a small image, 2 samples per sub-pixel
        w2, h2 = 160, 120
        sph2, _, cam2 = stock.scene(name, w2, h2)
        img = stock.mt_render(sph2, cam2, w2, h2, 2, 2, seed_mode=1)
        keep = [0, 37, 119]
        np.savez_compressed(os.path.join(HERE, f"mtrows_{name}.npz"), width=w2, height=h2, samps=2, nsub=2,
                            rows_y=np.array(keep), rows=np.stack([img[h2 - 1 - y] for y in keep]))

        # counter stream, whole small image + raw sums
        w3, h3 = 48, 36
        sph3, _, cam3 = stock.scene(name, w3, h3)
        img3, sums3 = ctr.render(sph3, cam3, w3, h3, 6, 2, seed=11, first_sample=3, want_sums=True)
        np.savez_compressed(os.path.join(HERE, f"image_{name}.npz"), width=w3, height=h3, samps=6, nsub=2, seed=11,
                            first_sample=3, image=img3, sums=sums3)

    # the reference program itself
    with tempfile.TemporaryDirectory() as tmp:
        rc = stock.reference_main(4, tmp)
        assert rc == 0
        with open(os.path.join(tmp, "image.ppm")) as f:
            tok = f.read().split()
    assert tok[0] == "P3" and int(tok[1]) == 1024 and int(tok[2]) == 768 and int(tok[3]) == 255
    px = np.array(tok[4:], dtype=np.int32).reshape(768, 1024, 3)
    ys = np.arange(0, 768, 64)
    np.savez_compressed(os.path.join(HERE, "refmain_rows.npz"), spp=4, rows_y=ys,
                        rows=np.stack([px[768 - 1 - y] for y in ys]).astype(np.uint8))
    # ---- sandbox/main.cpp (the stand-alone smallpt fork) ------------------------------------------------
    sb = Oracle("ref_sandbox")
    sph, cam8 = sb.sb_scene()
    W, H, N = 200, 150, 4000
    xs, ys = rng.integers(0, W, N), rng.integers(0, H, N)
    sx, sy = rng.integers(0, 2, N), rng.integers(0, 2, N)
    ss = rng.integers(0, 1 << 24, N)
    hit, rad, ray, draws = sb.sb_samples(sph, cam8, W, H, 13, xs, ys, sx, sy, ss)
    img_stock = sb.sb_render(sph, cam8, 96, 72, 2, mode=0)
    img_ctr = sb.sb_render(sph, cam8, 96, 72, 3, mode=1, seed=21, first_sample=5)
    with tempfile.TemporaryDirectory() as tmp:
        assert sb.sandbox_program(4, tmp) == 0
        with open(os.path.join(tmp, "image.ppm")) as f:
            tok = f.read().split()
    assert tok[:4] == ["P3", "1024", "768", "255"]
    px = np.array(tok[4:], dtype=np.int32).reshape(768, 1024, 3)
    rows = np.arange(0, 768, 48)
    np.savez_compressed(os.path.join(HERE, "sandbox.npz"), spheres=sph, cam8=cam8, width=W, height=H, seed=13,
                        x=xs.astype(np.uint32), y=ys.astype(np.uint32), sx=sx.astype(np.uint32), sy=sy.astype(np.uint32),
                        sample=ss.astype(np.uint32), hit=hit, radiance=rad, ray=ray, draws=draws.astype(np.uint32),
                        img_stock=img_stock, img_ctr=img_ctr, program_rows_y=rows,
                        program_rows=np.stack([px[768 - 1 - y] for y in rows]).astype(np.uint8))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
