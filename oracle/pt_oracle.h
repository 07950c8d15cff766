/* pt_oracle.h -- CPU restatement of the reference's per-pixel render loop.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker.  The product (cpu-path-tracing_b200/) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so this restatement is pinned against the reference
 * ITSELF, compiled unmodified into oracle/_ref/ (see ref/ref_driver.cpp):
 *   - bit-exact against the reference's stock mt19937 stream with seed 0 (the
 *     reproducible rows y%64==0 of the shipped program, and whole images in
 *     seed_mode 1), and
 *   - bit-exact against the reference's intersect/radiance/get_ray run with the
 *     injected counter stream,
 * both live (tests/test_oracle_vs_reference.py, where /root/reference exists) and
 * through fixtures committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * Record layouts are the reference's own ([probe]):
 *   sphere  88 B : radius@0  position@8  emission@32  color@56  reflection(int32)@80
 *                  (src/sphere.hpp:10-17; identical to sandbox/main.cpp:62-66)
 *   camera 176 B : position@0 lower_left_corner@24 cam_x_axis@48 cam_y_axis@72
 *                  u@96 v@120 w@144 lens_radius@168   (src/camera.hpp:23-32)
 *   camera_config 112 B : position@0 direction@24 up@48 aspect_ratio@72
 *                  vertical_fov_radians@80 focal_length@88 aperture@96
 *                  focus_distance@104                  (src/camera.hpp:11-21)
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_vec3
{
    double x, y, z;
} orc_vec3;

typedef struct orc_sphere
{
    double radius;
    orc_vec3 position;
    orc_vec3 emission;
    orc_vec3 color;
    int32_t reflection; /* 0 diffuse, 1 specular, 2 dielectric (src/reflection.hpp:7-12) */
    int32_t pad_;
} orc_sphere;

typedef struct orc_camera
{
    orc_vec3 position;
    orc_vec3 lower_left_corner;
    orc_vec3 cam_x_axis;
    orc_vec3 cam_y_axis;
    orc_vec3 u, v, w;
    double lens_radius;
} orc_camera;

typedef struct orc_camera_config
{
    orc_vec3 position;
    orc_vec3 direction; /* a look-at POINT (src/camera.cpp:8) */
    orc_vec3 up;
    double aspect_ratio;
    double vertical_fov_radians;
    double focal_length;
    double aperture;
    double focus_distance;
} orc_camera_config;

/* statistics slots filled by the render / sample entry points */
enum
{
    ORC_STAT_PATHS = 0,
    ORC_STAT_RAYS,          /* calls of intersect (main.cpp:30) */
    ORC_STAT_SPHERE_TESTS,  /* calls of sphere::intersect */
    ORC_STAT_DISC_NONNEG,   /* tests reaching the sqrt (sphere.cpp:18) */
    ORC_STAT_SECOND_ROOT,   /* tests evaluating the far root (sphere.cpp:22) */
    ORC_STAT_HIT_DIFFUSE,
    ORC_STAT_HIT_SPECULAR,
    ORC_STAT_HIT_DIELECTRIC,
    ORC_STAT_DIELECTRIC_REFLECT,
    ORC_STAT_RR_DRAWS,
    ORC_STAT_RR_KILLS,
    ORC_STAT_MISSES,
    ORC_STAT_DEPTH_LIMIT,
    ORC_STAT_DRAWS,
    ORC_STAT_COUNT
};

void orc_stats_reset(void);
void orc_stats_get(uint64_t out[ORC_STAT_COUNT]);

void orc_camera_with_config(void const* camera_config, void* camera_out);
void orc_color_to_int(double const* v, int n, int* out);
int orc_intersect(void const* spheres, int n, double const* origin, double const* direction, double* t_out);

/* Same signatures and meaning as ptref_ctr_samples / ptref_ctr_render /
 * ptref_stock_render in ref/ref_driver.cpp. */
void orc_samples(void const* spheres, int n, void const* camera, int width, int height, int num_subpixels,
                 uint64_t seed, uint32_t const* xs, uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys,
                 uint32_t const* samples, int count, int32_t* primary_hit, double* radiance_out, double* ray_out,
                 uint64_t* draws_out);

/* Sphere index hit at each depth of each sample's path: trail_out[count*trail_len], -1 = sky, -2 = path over. */
void orc_trails(void const* spheres, int n, void const* camera, int width, int height, int num_subpixels, uint64_t seed,
                uint32_t const* xs, uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys, uint32_t const* samples,
                int count, int trail_len, int32_t* trail_out);
void orc_render(void const* spheres, int n, void const* camera, int width, int height, int samps,
                int num_subpixels, uint64_t seed, uint32_t first_sample, double* image_out, double* sums_out,
                int nthreads);

/* seed_mode 1: mt19937{0} per row (the only reproducible mode of the stock stream).
 * seed_mode 2: mt19937{row_seed[y]} per row (an explicit 32-bit engine seed per row). */
void orc_mt_render(void const* spheres, int n, void const* camera, int width, int height, int samps,
                   int num_subpixels, int seed_mode, uint32_t const* row_seed, int y0, int y1, double* image_out,
                   int nthreads);

/* ---- sandbox/main.cpp (the stand-alone smallpt fork), pt_oracle_sandbox.c -------------------------
 * cam8 = camera position(3), direction(3, un-normalised), field-of-view factor (.5135), push (140):
 * the constants of sandbox/main.cpp:235-237,260.  mode 0 = the program's erand48 stream,
 * 1 = counter stream.  Same signatures as sbref_render / sbref_samples in ref/sandbox_driver.cpp,
 * plus the scene, which the sandbox keeps in a global array. */
void orc_sb_render(void const* spheres, int n, double const* cam8, int w, int h, int samps, int mode, uint64_t seed,
                   uint32_t first_sample, int y0, int y1, double* image_out, int nthreads);
void orc_sb_samples(void const* spheres, int n, double const* cam8, int w, int h, uint64_t seed, uint32_t const* xs,
                    uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys, uint32_t const* samples, int count,
                    int32_t* primary_hit, double* radiance_out, double* ray_out, uint64_t* draws_out);
void orc_sb_to_int(double const* v, int n, int* out);

#ifdef __cplusplus
}
#endif

#endif /* PT_ORACLE_H */
