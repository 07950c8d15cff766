// sandbox_driver.cpp -- C entry points around the UNMODIFIED sandbox/main.cpp (the reference's
// stand-alone smallpt fork).
//
// TEST INFRASTRUCTURE, built only where /root/reference exists (oracle/Makefile -> oracle/_ref/).
// The sandbox's own Vec / Ray / Sphere / spheres[] / intersect / radiance are compiled as they are
// (`main` renamed, never called; the program itself is built separately as oracle/_ref/smallpt).
// Two things are re-stated here because they are locals of its main(): the camera set-up
// (sandbox/main.cpp:235-238) and the pixel / sub-pixel / sample loop (:241-269).
// `erand48` is renamed to a shim that either forwards to libc's erand48 (stock stream, Xi as the
// program seeds it) or draws from the counter stream of oracle/ptb_rng.h (what the CUDA kernels use).
#include <cstdint>
#include <cstring>
#include <stdlib.h>

#include "../ptb_rng.h"

namespace {
double (*const libc_erand48)(unsigned short*) = erand48;
thread_local bool tl_counter_mode = false;
thread_local ptb_rng tl_rng;
double sb_erand48(unsigned short* xi)
{
    return tl_counter_mode ? ptb_rng_uniform(&tl_rng) : libc_erand48(xi);
}
} // namespace

#define erand48 sb_erand48
#define main sandbox_main
#include "main.cpp" // /root/reference/sandbox/main.cpp via -I
#undef main
#undef erand48

#include <omp.h>

static_assert(sizeof(Sphere) == 88, "sandbox Sphere has the 88-byte layout of pt::sphere (sandbox/main.cpp:62-66)");

extern "C" {

int sbref_scene(void* spheres_out, int max_spheres)
{
    int const n = static_cast<int>(sizeof(spheres) / sizeof(Sphere));
    if(n > max_spheres) {
        return -n;
    }
    std::memcpy(spheres_out, static_cast<void const*>(spheres), sizeof(spheres));
    return n;
}

// camera constants of sandbox/main.cpp:235 (position, direction, field-of-view factor, push-forward)
void sbref_camera(double* out8)
{
    double const v[8] = { 50, 52, 295.6, 0, -0.042612, -1, .5135, 140 };
    std::memcpy(out8, v, sizeof(v));
}

// The loop nest of sandbox/main.cpp:241-269 with run-time w, h.
//   mode 0: the program's stream -- erand48 with Xi = {0, 0, (unsigned short)(y*y*y)} per row
//   mode 1: counter stream re-keyed per sample with (seed, slot, sample), slot as in ptb_rng.h
// image_out: w*h*3 doubles, rows [y0, y1) are written (accumulated into, caller zeroes).
void sbref_render(int const w, int const h, int const samps, int const mode, std::uint64_t const seed,
                  std::uint32_t const first_sample, int const y0, int const y1, double* image_out, int const nthreads)
{
    Ray cam(Vec(50, 52, 295.6), Vec(0, -0.042612, -1).norm());
    Vec const cx = Vec(w * .5135 / h);
    Vec const cy = (cx % cam.d).norm() * .5135;
    Vec r;
    Vec* c = reinterpret_cast<Vec*>(image_out);
    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();

#pragma omp parallel for schedule(dynamic, 1) private(r) num_threads(threads)
    for(int y = y0; y < y1; y++) {
        tl_counter_mode = mode == 1;
        unsigned short Xi[3] = { 0, 0, (unsigned short)(y * y * y) };

        for(int x = 0; x < w; x++) {
            int i = (h - y - 1) * w + x;

            for(int sy = 0; sy < 2; sy++) {
                for(int sx = 0; sx < 2; sx++) {
                    for(int s = 0; s < samps; s++) {
                        if(mode == 1) {
                            std::uint32_t const slot =
                                ((static_cast<std::uint32_t>(y) * static_cast<std::uint32_t>(w) + static_cast<std::uint32_t>(x)) * 2u +
                                 static_cast<std::uint32_t>(sy)) * 2u + static_cast<std::uint32_t>(sx);
                            ptb_rng_key(&tl_rng, seed, slot, first_sample + static_cast<std::uint32_t>(s));
                        }
                        double const r1 = 2 * sb_erand48(Xi);
                        double const dx = r1 < 1 ? sqrt(r1) - 1 : 1 - sqrt(2 - r1);
                        double const r2 = 2 * sb_erand48(Xi);
                        double const dy = r2 < 1 ? sqrt(r2) - 1 : 1 - sqrt(2 - r2);

                        Vec d =
                            cx * (((sx + .5 + dx) / 2 + x) / w - .5) + cy * (((sy + .5 + dy) / 2 + y) / h - .5) + cam.d;

                        r = r + radiance(Ray(cam.o + d * 140, d.norm()), 0, Xi) * (1. / samps);
                    }
                    c[i] = c[i] + Vec(clamp(r.x), clamp(r.y), clamp(r.z)) * .25;
                    r = Vec();
                }
            }
        }
    }
}

// Per-sample probe with the counter stream: primary-hit index (-1 = miss), radiance, camera ray.
void sbref_samples(int const w, int const h, std::uint64_t const seed, std::uint32_t const* xs, std::uint32_t const* ys,
                   std::uint32_t const* sxs, std::uint32_t const* sys, std::uint32_t const* samples, int const count,
                   std::int32_t* primary_hit, double* radiance_out, double* ray_out, std::uint64_t* draws_out)
{
    Ray cam(Vec(50, 52, 295.6), Vec(0, -0.042612, -1).norm());
    Vec const cx = Vec(w * .5135 / h);
    Vec const cy = (cx % cam.d).norm() * .5135;
#pragma omp parallel for schedule(static)
    for(int k = 0; k < count; ++k) {
        tl_counter_mode = true;
        unsigned short Xi[3] = { 0, 0, 0 };
        int const x = static_cast<int>(xs[k]), y = static_cast<int>(ys[k]);
        int const sx = static_cast<int>(sxs[k]), sy = static_cast<int>(sys[k]);
        std::uint32_t const slot =
            ((ys[k] * static_cast<std::uint32_t>(w) + xs[k]) * 2u + sys[k]) * 2u + sxs[k];
        ptb_rng_key(&tl_rng, seed, slot, samples[k]);
        tl_rng.draws = 0;
        double const r1 = 2 * sb_erand48(Xi);
        double const dx = r1 < 1 ? sqrt(r1) - 1 : 1 - sqrt(2 - r1);
        double const r2 = 2 * sb_erand48(Xi);
        double const dy = r2 < 1 ? sqrt(r2) - 1 : 1 - sqrt(2 - r2);
        Vec d = cx * (((sx + .5 + dx) / 2 + x) / w - .5) + cy * (((sy + .5 + dy) / 2 + y) / h - .5) + cam.d;
        Ray const ray(cam.o + d * 140, d.norm());
        double t;
        int id = 0;
        primary_hit[k] = intersect(ray, t, id) ? id : -1;
        Vec const L = radiance(ray, 0, Xi);
        radiance_out[3 * k + 0] = L.x;
        radiance_out[3 * k + 1] = L.y;
        radiance_out[3 * k + 2] = L.z;
        if(ray_out != nullptr) {
            double const v[6] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z };
            std::memcpy(ray_out + 6 * k, v, sizeof(v));
        }
        if(draws_out != nullptr) {
            draws_out[k] = tl_rng.draws;
        }
    }
}

// toInt of sandbox/main.cpp:130-133
void sbref_to_int(double const* v, int n, int* out)
{
    for(int i = 0; i < n; ++i) {
        out[i] = toInt(v[i]);
    }
}

} // extern "C"
