// ref_driver.cpp -- C entry points around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE.  Built only where /root/reference exists (this container)
// by oracle/Makefile into oracle/_ref/; never shipped, never on the product path.
// Nothing here re-implements reference arithmetic: intersect / radiance /
// render_subpixel / camera::get_ray / the scene builders are the reference's own
// object code, pulled in by including src/main.cpp with `main` renamed
// (SURVEY.md section 8c).  What IS re-stated here is loop plumbing only:
//   * the row loop of src/main.cpp:217-233 (so that width/height/scene, which
//     are constexpr in the reference's main, can be chosen at run time), and
//   * in the PTREF_CTR build, the 8-line sample loop of src/main.cpp:184-193,
//     so that the injected counter-based stream can be re-keyed per sample.
//
// Two builds of this file:
//   libptref_stock.so : stock pt::rand_state (mt19937, random_state.cpp)
//   libptref_ctr.so   : -DPTREF_CTR -include inject_rng.hpp (counter stream)

#include <cstdint>
#include <cstring>

// `main` is renamed only to reach the file-scope functions next to it; reference_main is
// never called (a renamed main has no implicit `return 0`).  The reference PROGRAM is built
// as its own executable, oracle/_ref/cpu_path_tracer, by the Makefile.
#define main reference_main
#include "main.cpp" // /root/reference/src/main.cpp via -I
#undef main

// The three scene headers share one include guard and two of them define the
// same pt::box_scene: main.cpp took box_mirror_scene.hpp;
// re-open the guard and rename for the other two.
#undef PT_SMALLPT_SCENE_HPP
#include "simple_scene.hpp"
#undef PT_SMALLPT_SCENE_HPP
#define box_scene box_scene_plain
#include "box_scene.hpp"
#undef box_scene

#include <omp.h>

static_assert(sizeof(pt::sphere) == 88, "pt::sphere layout (SURVEY 8a row a3)");
static_assert(sizeof(pt::camera) == 176, "pt::camera layout (SURVEY 8a row a12)");
static_assert(sizeof(pt::camera_config) == 112, "pt::camera_config layout");

namespace {

auto make_scene(void const* spheres, int const n) -> pt::scene
{
    pt::scene scn{};
    scn.spheres.resize(static_cast<std::size_t>(n));
    std::memcpy(static_cast<void*>(scn.spheres.data()), spheres, sizeof(pt::sphere) * static_cast<std::size_t>(n));
    return scn;
}

auto make_camera(void const* camera) -> pt::camera
{
    pt::camera cam{};
    std::memcpy(static_cast<void*>(&cam), camera, sizeof(pt::camera));
    return cam;
}

} // namespace

extern "C" {

// name: "simple" | "box" | "box_mirror".  Returns 0, or -1 unknown name, -2 capacity.
int ptref_scene(char const* name,
                int const w,
                int const h,
                void* spheres_out,
                int const max_spheres,
                int* n_out,
                void* camera_config_out,
                void* camera_out)
{
    pt::scene scn{};
    if(std::strcmp(name, "simple") == 0) {
        scn = pt::simple_scene(w, h);
    }
    else if(std::strcmp(name, "box") == 0) {
        scn = pt::box_scene_plain(w, h);
    }
    else if(std::strcmp(name, "box_mirror") == 0) {
        scn = pt::box_scene(w, h);
    }
    else {
        return -1;
    }
    int const n = static_cast<int>(scn.spheres.size());
    *n_out = n;
    if(n > max_spheres) {
        return -2;
    }
    std::memcpy(spheres_out, scn.spheres.data(), sizeof(pt::sphere) * scn.spheres.size());
    if(camera_config_out != nullptr) {
        std::memcpy(camera_config_out, &scn.camera_parameters, sizeof(pt::camera_config));
    }
    if(camera_out != nullptr) {
        auto const cam = pt::camera::with_config(scn.camera_parameters);
        std::memcpy(camera_out, &cam, sizeof(pt::camera));
    }
    return 0;
}

// pt::camera::with_config on a caller-supplied 112-byte camera_config.
void ptref_camera_with_config(void const* camera_config, void* camera_out)
{
    pt::camera_config cfg{};
    std::memcpy(static_cast<void*>(&cfg), camera_config, sizeof(cfg));
    auto const cam = pt::camera::with_config(cfg);
    std::memcpy(camera_out, &cam, sizeof(cam));
}

// pt::color_to_int (utils.cpp:11-16) over an array.
void ptref_color_to_int(double const* v, int const n, int* out)
{
    for(int i = 0; i < n; ++i) {
        out[i] = pt::color_to_int(v[i]);
    }
}

// One closest-hit query through the reference's intersect (main.cpp:30-42).
int ptref_intersect(void const* spheres, int const n, double const* origin, double const* direction, double* t_out)
{
    auto const scn = make_scene(spheres, n);
    pt::ray const r{ pt::vec3{ origin[0], origin[1], origin[2] }, pt::vec3{ direction[0], direction[1], direction[2] } };
    double t = 0.0;
    std::size_t id = 0;
    bool const hit = intersect(scn, r, t, id);
    *t_out = t;
    return hit ? static_cast<int>(id) : -1;
}

#ifndef PTREF_CTR

// Row loop of main.cpp:217-233 with run-time width/height/scene; every sample goes
// through the reference's own render_subpixel (main.cpp:179-197) and stock mt19937
// stream.  seed_mode 0: the reference's seed (unsigned short)(y*y*y), which
// random_state.cpp:5 multiplies by random_device() -> only rows with y%64==0 are
// reproducible.  seed_mode 1: seed 0 on every row -> mt19937{0} per row, the whole
// image is reproducible.  image: W*H*3 doubles, accumulated into (caller zeroes).
void ptref_stock_render(void const* spheres,
                        int const n,
                        void const* camera,
                        int const width,
                        int const height,
                        int const samps,
                        int const num_subpixels,
                        int const seed_mode,
                        int const y0,
                        int const y1,
                        double* image_out,
                        int const nthreads)
{
    auto const scn = make_scene(spheres, n);
    auto const cam = make_camera(camera);
    std::vector<pt::vec3> image{};
    image.resize(static_cast<std::size_t>(width) * static_cast<std::size_t>(height), pt::vec3{ 0, 0, 0 });

    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for(int y = y0; y < y1; y++) {
        auto const seed = seed_mode == 0 ? static_cast<unsigned short>(y * y * y) : static_cast<unsigned short>(0);
        auto rng = pt::rand_state::default_with_seed(seed);
        render_state state{ scn, cam, rng, image, width, height, samps, num_subpixels };

        for(int x = 0; x < width; x++) {
            for(int sy = 0; sy < num_subpixels; sy++) {
                for(int sx = 0; sx < num_subpixels; sx++) {
                    render_subpixel(x, y, sx, sy, state);
                }
            }
        }
    }

    for(std::size_t i = 0; i < image.size(); ++i) {
        image_out[3 * i + 0] += image[i].x;
        image_out[3 * i + 1] += image[i].y;
        image_out[3 * i + 2] += image[i].z;
    }
}

#else // PTREF_CTR

// Per-sample probe: for each (x, y, sx, sy, sample) re-key the injected stream,
// then follow main.cpp:186-191 -- jitter, camera::get_ray, radiance -- with the
// reference's own functions.  primary_hit = id from intersect() on the camera ray
// (-1 on miss), radiance = 3 doubles, ray = origin+direction (6 doubles, optional).
void ptref_ctr_samples(void const* spheres,
                       int const n,
                       void const* camera,
                       int const width,
                       int const height,
                       int const num_subpixels,
                       std::uint64_t const seed,
                       std::uint32_t const* xs,
                       std::uint32_t const* ys,
                       std::uint32_t const* sxs,
                       std::uint32_t const* sys,
                       std::uint32_t const* samples,
                       int const count,
                       std::int32_t* primary_hit,
                       double* radiance_out,
                       double* ray_out,
                       std::uint64_t* draws_out)
{
    auto const scn = make_scene(spheres, n);
    auto const cam = make_camera(camera);
    auto const ns = static_cast<std::uint32_t>(num_subpixels);

#pragma omp parallel for schedule(static)
    for(int i = 0; i < count; ++i) {
        auto rng = pt::rand_state::default_with_seed(0);
        std::uint32_t const x = xs[i], y = ys[i], sx = sxs[i], sy = sys[i];
        std::uint32_t const slot = ((y * static_cast<std::uint32_t>(width) + x) * ns + sy) * ns + sx;
        rng.key(seed, slot, samples[i]);
        rng.g.draws = 0;

        double const subpixel_length = 1.0 / num_subpixels;
        double const x_in_subpixel = (x + sx * subpixel_length + rng.generate_between(0.0, subpixel_length));
        double const y_in_subpixel = (y + sy * subpixel_length + rng.generate_between(0.0, subpixel_length));
        pt::ray const new_ray = cam.get_ray(x_in_subpixel / width, y_in_subpixel / height, rng);

        double t = 0.0;
        std::size_t id = 0;
        primary_hit[i] = intersect(scn, new_ray, t, id) ? static_cast<std::int32_t>(id) : -1;

        pt::vec3 const contributor = radiance(scn, new_ray, rng);
        radiance_out[3 * i + 0] = contributor.x;
        radiance_out[3 * i + 1] = contributor.y;
        radiance_out[3 * i + 2] = contributor.z;
        if(ray_out != nullptr) {
            ray_out[6 * i + 0] = new_ray.origin.x;
            ray_out[6 * i + 1] = new_ray.origin.y;
            ray_out[6 * i + 2] = new_ray.origin.z;
            ray_out[6 * i + 3] = new_ray.direction.x;
            ray_out[6 * i + 4] = new_ray.direction.y;
            ray_out[6 * i + 5] = new_ray.direction.z;
        }
        if(draws_out != nullptr) {
            draws_out[i] = rng.g.draws;
        }
    }
}

// Whole image with the counter stream.  Loop nest of main.cpp:217-232 and sample
// loop / clamp / accumulate of main.cpp:184-196.  Renders samples
// [first_sample, first_sample+samps) of each sub-pixel; the mean is over `samps`.
// sums_out (optional): un-clamped per-sub-pixel radiance SUMS, layout
// [slot][3] with slot as in ptb_rng.h -- what the GPU accumulation buffer holds.
void ptref_ctr_render(void const* spheres,
                      int const n,
                      void const* camera,
                      int const width,
                      int const height,
                      int const samps,
                      int const num_subpixels,
                      std::uint64_t const seed,
                      std::uint32_t const first_sample,
                      double* image_out,
                      double* sums_out,
                      int const nthreads)
{
    auto const scn = make_scene(spheres, n);
    auto const cam = make_camera(camera);
    auto const ns = static_cast<std::uint32_t>(num_subpixels);
    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();

#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for(int y = 0; y < height; y++) {
        auto rng = pt::rand_state::default_with_seed(0);
        for(int x = 0; x < width; x++) {
            auto const row = static_cast<std::size_t>((height - y - 1) * width) + static_cast<std::size_t>(x);
            pt::vec3 pixel{ 0, 0, 0 };
            for(int sy = 0; sy < num_subpixels; sy++) {
                for(int sx = 0; sx < num_subpixels; sx++) {
                    std::uint32_t const slot = ((static_cast<std::uint32_t>(y) * static_cast<std::uint32_t>(width) +
                                                 static_cast<std::uint32_t>(x)) *
                                                    ns +
                                                static_cast<std::uint32_t>(sy)) *
                                                   ns +
                                               static_cast<std::uint32_t>(sx);
                    pt::vec3 r{ 0, 0, 0 };
                    pt::vec3 sum{ 0, 0, 0 };
                    for(int s = 0; s < samps; s++) {
                        rng.key(seed, slot, first_sample + static_cast<std::uint32_t>(s));
                        double const subpixel_length = 1.0 / num_subpixels;
                        double const x_in_subpixel =
                            (x + sx * subpixel_length + rng.generate_between(0.0, subpixel_length));
                        double const y_in_subpixel =
                            (y + sy * subpixel_length + rng.generate_between(0.0, subpixel_length));
                        pt::ray const new_ray = cam.get_ray(x_in_subpixel / width, y_in_subpixel / height, rng);
                        pt::vec3 const contributor = radiance(scn, new_ray, rng);
                        r = r + contributor * (1.0 / samps);
                        sum = sum + contributor;
                    }
                    pt::vec3 const subpixel_color = pt::vec3{ pt::clamp(r.x), pt::clamp(r.y), pt::clamp(r.z) };
                    pixel = pixel + subpixel_color * (1.0 / (num_subpixels * num_subpixels));
                    if(sums_out != nullptr) {
                        sums_out[3 * static_cast<std::size_t>(slot) + 0] = sum.x;
                        sums_out[3 * static_cast<std::size_t>(slot) + 1] = sum.y;
                        sums_out[3 * static_cast<std::size_t>(slot) + 2] = sum.z;
                    }
                }
            }
            image_out[3 * row + 0] = pixel.x;
            image_out[3 * row + 1] = pixel.y;
            image_out[3 * row + 2] = pixel.z;
        }
    }
}

#endif // PTREF_CTR

} // extern "C"
