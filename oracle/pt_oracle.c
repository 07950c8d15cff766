/* pt_oracle.c -- plain-C, FP64 restatement of the reference's per-pixel render loop.
 *
 * TEST INFRASTRUCTURE -- see pt_oracle.h for who may load this and how it is
 * pinned against the reference itself.  Every function cites the reference lines
 * it follows (paths relative to /root/reference).  Operation ORDER is kept exactly
 * (left-to-right association of the reference's expressions), because the bar is
 * bit-exact agreement with the compiled reference under -ffp-contract=off.
 */
#include "pt_oracle.h"
#include "ptb_rng.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

/* src/constants.hpp:7-10 */
#define ORC_EPSILON 1e-4
#define ORC_PI 3.14159265358979323846
#define ORC_INF 1e20
#define ORC_DEPTH_LIMIT 100

/* ------------------------------------------------------------------ vec3 */
/* src/vec.cpp:15-69.  dot is (x*bx + y*by) + z*bz; norm is v * (1/sqrt(.)). */
typedef orc_vec3 v3;

static inline v3 v3_make(double x, double y, double z)
{
    v3 r = { x, y, z };
    return r;
}
static inline v3 v3_add(v3 a, v3 b) /* vec.cpp:15-18 */
{
    return v3_make(a.x + b.x, a.y + b.y, a.z + b.z);
}
static inline v3 v3_sub(v3 a, v3 b) /* vec.cpp:20-23 */
{
    return v3_make(a.x - b.x, a.y - b.y, a.z - b.z);
}
static inline v3 v3_scale(v3 a, double b) /* vec.cpp:25-28 */
{
    return v3_make(a.x * b, a.y * b, a.z * b);
}
static inline v3 v3_blend(v3 a, v3 b) /* vec.cpp:30-33 */
{
    return v3_make(a.x * b.x, a.y * b.y, a.z * b.z);
}
static inline double v3_dot(v3 a, v3 b) /* vec.cpp:40-43 */
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
static inline v3 v3_norm(v3 a) /* vec.cpp:35-38 */
{
    return v3_scale(a, 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z));
}
static inline v3 v3_cross(v3 a, v3 b) /* vec.cpp:45-48 */
{
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

typedef struct ray
{
    v3 o, d;
} ray;

/* ------------------------------------------------------------------ RNG */
/* Two interchangeable streams behind generate()/generate_between()
 * (src/random_state.cpp:9-17):
 *   counter: oracle/ptb_rng.h (what the CUDA kernels use)
 *   mt     : std::mt19937 + libstdc++ uniform_real_distribution<double>(0,1)
 *            = generate_canonical<double,53>: two 32-bit engine outputs,
 *            (u1 + u2*2^32) / 2^64 in double, clamped below 1
 *            (src/random_state.hpp:14-15; libstdc++ bits/random.tcc). */
typedef struct mt19937
{
    uint32_t mt[624];
    int idx;
} mt19937;

static void mt_seed(mt19937* m, uint32_t s)
{
    m->mt[0] = s;
    for(int i = 1; i < 624; ++i) {
        m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    }
    m->idx = 624;
}

static uint32_t mt_next(mt19937* m)
{
    if(m->idx >= 624) {
        for(int i = 0; i < 624; ++i) {
            uint32_t const y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7fffffffu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

typedef struct rng
{
    int use_mt;
    ptb_rng ctr;
    mt19937 mt;
    uint64_t draws;
} rng;

static inline double rng_generate(rng* g) /* random_state.cpp:9-12 */
{
    g->draws++;
    if(!g->use_mt) {
        return ptb_rng_uniform(&g->ctr);
    }
    double sum = 0.0;
    double tmp = 1.0;
    sum += (double)mt_next(&g->mt) * tmp;
    tmp *= 4294967296.0;
    sum += (double)mt_next(&g->mt) * tmp;
    tmp *= 4294967296.0;
    double ret = sum / tmp;
    if(ret >= 1.0) {
        ret = nextafter(1.0, 0.0);
    }
    return ret;
}

static inline double rng_between(rng* g, double mn, double mx) /* random_state.cpp:14-17 */
{
    return mn + (mx - mn) * rng_generate(g);
}

/* ------------------------------------------------------------------ stats */
static uint64_t g_stats[ORC_STAT_COUNT];

typedef struct stats
{
    uint64_t c[ORC_STAT_COUNT];
    int32_t* trail; /* orc_trails only: sphere hit at each depth < trail_len (-1 = sky); NULL otherwise */
    int trail_len;
} stats;

static void stats_merge(stats const* s)
{
#pragma omp critical(orc_stats)
    for(int i = 0; i < ORC_STAT_COUNT; ++i) {
        g_stats[i] += s->c[i];
    }
}

void orc_stats_reset(void)
{
    memset(g_stats, 0, sizeof(g_stats));
}

void orc_stats_get(uint64_t out[ORC_STAT_COUNT])
{
    memcpy(out, g_stats, sizeof(g_stats));
}

/* ------------------------------------------------------------------ geometry */
/* src/sphere.cpp:6-30 */
static inline double sphere_intersect(orc_sphere const* s, ray const* r, stats* st)
{
    v3 const oc = v3_sub(r->o, s->position);
    double const a = v3_dot(r->d, r->d);
    double const half_b = v3_dot(oc, r->d);
    double const c = v3_dot(oc, oc) - s->radius * s->radius;
    double const discriminant = half_b * half_b - a * c;
    st->c[ORC_STAT_SPHERE_TESTS]++;

    if(discriminant < 0) {
        return 0.0;
    }
    st->c[ORC_STAT_DISC_NONNEG]++;

    double const sqrtd = sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;

    if(root < ORC_EPSILON) {
        st->c[ORC_STAT_SECOND_ROOT]++;
        root = (-half_b + sqrtd) / a;
        if(root < ORC_EPSILON) {
            return 0.0;
        }
    }
    return root;
}

/* src/main.cpp:30-42: ascending index, strict '<' so the lowest index wins ties */
static inline int scene_intersect(orc_sphere const* sph, int n, ray const* r, double* t, int* id, stats* st)
{
    *t = ORC_INF;
    st->c[ORC_STAT_RAYS]++;
    for(int i = 0; i < n; i++) {
        double const d = sphere_intersect(&sph[i], r, st);
        if(d > 0 && d < *t) {
            *t = d;
            *id = i;
        }
    }
    return *t < ORC_INF;
}

/* src/hit_record.hpp:11-18, hit_record.cpp:3-12 */
typedef struct hit_record
{
    ray original_ray;
    v3 hit_point;
    v3 outward_normal;
    v3 normal;
    int front_facing;
} hit_record;

static inline hit_record get_hit_record_at(orc_sphere const* s, ray const* r, double t)
{
    hit_record h;
    h.original_ray = *r;
    h.hit_point = v3_add(r->o, v3_scale(r->d, t)); /* ray.cpp:3-6 */
    h.outward_normal = v3_norm(v3_sub(h.hit_point, s->position));
    h.front_facing = v3_dot(h.outward_normal, r->d) < 0;
    h.normal = h.front_facing ? h.outward_normal : v3_scale(h.outward_normal, -1);
    return h;
}

/* ------------------------------------------------------------------ scattering */
/* src/main.cpp:44-58 */
static inline ray diffuse_ray(hit_record const* rec, rng* g)
{
    double const phi = 2 * ORC_PI * rng_generate(g);
    double const random_angle = rng_generate(g);
    double const sin_theta = sqrt(random_angle);
    double const cos_theta = sqrt(1.0 - random_angle);

    v3 const w = rec->normal;
    v3 const u = v3_norm(v3_cross(fabs(w.x) > 0.1 ? v3_make(0, 1, 0) : v3_make(1, 0, 0), w));
    v3 const v = v3_cross(w, u);
    v3 const nd = v3_norm(v3_add(v3_add(v3_scale(v3_scale(u, cos(phi)), sin_theta), v3_scale(v3_scale(v, sin(phi)), sin_theta)),
                                 v3_scale(w, cos_theta)));
    ray out = { rec->hit_point, nd };
    return out;
}

/* src/main.cpp:60-67: outward normal, un-normalised incoming direction, one dead draw */
static inline ray specular_ray(hit_record const* rec, rng* g)
{
    double const fuzziness = 0.0;
    v3 const d = rec->original_ray.d;
    v3 const reflected = v3_sub(d, v3_scale(v3_scale(rec->outward_normal, 2.0), v3_dot(rec->outward_normal, d)));
    double const factor = rng_generate(g) * fuzziness;
    ray out = { rec->hit_point, v3_add(reflected, v3_make(factor, factor, factor)) };
    return out;
}

/* src/main.cpp:79-84 */
static inline double reflectance(double cosine, double ref_idx)
{
    double r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
    r0 *= r0;
    return r0 + (1.0 - r0) * pow(1.0 - cosine, 5);
}

/* src/main.cpp:69-97 */
static inline ray dielectric_ray(hit_record const* rec, rng* g, stats* st)
{
    double const refraction_index = 2.0;
    double const refraction_ratio = rec->front_facing ? (1.0 / refraction_index) : refraction_index;

    v3 const unit_direction = v3_norm(rec->original_ray.d);

    double const cos_theta = fmin(v3_dot(v3_scale(unit_direction, -1.0), rec->normal), 1.0);
    double const sin_theta = sqrt(1.0 - cos_theta * cos_theta);

    int const cannot_refract = refraction_ratio * sin_theta > 1.0;

    if(cannot_refract || reflectance(cos_theta, refraction_ratio) > rng_generate(g)) {
        st->c[ORC_STAT_DIELECTRIC_REFLECT]++;
        return specular_ray(rec, g);
    }

    v3 const r_out_perp = v3_scale(v3_add(unit_direction, v3_scale(rec->normal, cos_theta)), refraction_ratio);
    v3 const r_out_parallel = v3_scale(rec->normal, -sqrt(fabs(1.0 - v3_dot(r_out_perp, r_out_perp))));
    ray out = { rec->hit_point, v3_add(r_out_perp, r_out_parallel) };
    return out;
}

/* ------------------------------------------------------------------ integrator */
/* src/main.cpp:104-158 */
static v3 radiance(orc_sphere const* sph, int n, ray const* primary, rng* g, stats* st)
{
    int const russian_roulette_threshold = 4;
    v3 accumulated_emission = v3_make(0.0, 0.0, 0.0);
    v3 accumulated_reflectance = v3_make(1, 1, 1);
    ray r = *primary;

    for(int depth = 0; depth < ORC_DEPTH_LIMIT; ++depth) {
        double closest_distance = 0.0;
        int object_index = 0;

        int const any_hit = scene_intersect(sph, n, &r, &closest_distance, &object_index, st);
        if(st->trail != NULL && depth < st->trail_len) {
            st->trail[depth] = any_hit ? object_index : -1;
        }
        if(!any_hit) {
            v3 const unit_direction = v3_norm(r.d);
            double const t = 0.5 * (unit_direction.y + 1.0);
            v3 const background = v3_add(v3_scale(v3_make(1.0, 1.0, 1.0), 1.0 - t), v3_scale(v3_make(0.5, 0.7, 1.0), t));
            st->c[ORC_STAT_MISSES]++;
            return v3_add(accumulated_emission, v3_blend(accumulated_reflectance, background));
        }

        orc_sphere const* obj = &sph[object_index];
        hit_record const record = get_hit_record_at(obj, &r, closest_distance);
        v3 color = obj->color;

        accumulated_emission = v3_add(accumulated_emission, v3_blend(accumulated_reflectance, obj->emission));

        double const probability = fmax(fmax(color.x, color.y), color.z);

        if(depth > russian_roulette_threshold) {
            st->c[ORC_STAT_RR_DRAWS]++;
            if(rng_generate(g) < probability) {
                color = v3_scale(color, 1.0 / probability);
            }
            else {
                st->c[ORC_STAT_RR_KILLS]++;
                return accumulated_emission;
            }
        }

        accumulated_reflectance = v3_blend(accumulated_reflectance, color);

        switch(obj->reflection) {
        case 0:
            st->c[ORC_STAT_HIT_DIFFUSE]++;
            r = diffuse_ray(&record, g);
            break;
        case 1:
            st->c[ORC_STAT_HIT_SPECULAR]++;
            r = specular_ray(&record, g);
            break;
        case 2:
            st->c[ORC_STAT_HIT_DIELECTRIC]++;
            r = dielectric_ray(&record, g, st);
            break;
        default:
            break; /* the reference's switch has no default either: ray unchanged */
        }
    }

    st->c[ORC_STAT_DEPTH_LIMIT]++;
    return accumulated_emission;
}

/* ------------------------------------------------------------------ camera */
/* src/camera.cpp:3-17 */
void orc_camera_with_config(void const* camera_config, void* camera_out)
{
    orc_camera_config cfg;
    memcpy(&cfg, camera_config, sizeof(cfg));
    double const viewport_height = 2.0 * tan(0.5 * cfg.vertical_fov_radians);
    double const viewport_width = cfg.aspect_ratio * viewport_height;

    v3 const w = v3_norm(v3_sub(cfg.position, cfg.direction));
    v3 const u = v3_norm(v3_cross(cfg.up, w));
    v3 const v = v3_cross(w, u);

    v3 const cam_x_axis = v3_scale(v3_scale(u, viewport_width), cfg.focus_distance);
    v3 const cam_y_axis = v3_scale(v3_scale(v, viewport_height), cfg.focus_distance);
    v3 const llc = v3_sub(v3_sub(v3_sub(cfg.position, v3_scale(cam_x_axis, 0.5)), v3_scale(cam_y_axis, 0.5)),
                          v3_scale(w, cfg.focus_distance));

    orc_camera cam = { cfg.position, llc, cam_x_axis, cam_y_axis, u, v, w, cfg.aperture / 2.0 };
    memcpy(camera_out, &cam, sizeof(cam));
}

/* src/camera.cpp:19-30: rejection sampling, x drawn before y (brace-init order) */
static inline v3 random_in_unit_disk(rng* g)
{
    for(;;) {
        double const px = rng_between(g, -1.0, 1.0);
        double const py = rng_between(g, -1.0, 1.0);
        v3 const point = v3_make(px, py, 0.0);
        if(v3_dot(point, point) >= 1.0) {
            continue;
        }
        return point;
    }
}

/* src/camera.cpp:32-38: note offset = rd*s + rd*t (not u*rd.x + v*rd.y) and the
 * direction is NOT normalised */
static inline ray camera_get_ray(orc_camera const* cam, double s, double t, rng* g)
{
    v3 const rd = v3_scale(random_in_unit_disk(g), cam->lens_radius);
    v3 const offset = v3_add(v3_scale(rd, s), v3_scale(rd, t));
    v3 const direction = v3_sub(
        v3_sub(v3_add(v3_add(cam->lower_left_corner, v3_scale(cam->cam_x_axis, s)), v3_scale(cam->cam_y_axis, t)), cam->position),
        offset);
    ray out = { v3_add(cam->position, offset), direction };
    return out;
}

/* src/main.cpp:186-190: jittered position inside the stratum -> camera ray */
static inline ray primary_ray(orc_camera const* cam, int x, int y, int sx, int sy, int width, int height,
                              int num_subpixels, rng* g)
{
    double const subpixel_length = 1.0 / num_subpixels;
    double const x_in_subpixel = (x + sx * subpixel_length + rng_between(g, 0.0, subpixel_length));
    double const y_in_subpixel = (y + sy * subpixel_length + rng_between(g, 0.0, subpixel_length));
    return camera_get_ray(cam, x_in_subpixel / width, y_in_subpixel / height, g);
}

/* src/utils.cpp:6-9 */
static inline double clamp01(double x)
{
    return x < 0.0 ? 0.0 : (1.0 < x ? 1.0 : x);
}

/* src/utils.cpp:11-16 */
void orc_color_to_int(double const* v, int n, int* out)
{
    for(int i = 0; i < n; ++i) {
        double const corrected = pow(clamp01(v[i]), 1.0 / 2.2);
        out[i] = (int)round(corrected * 255.0);
    }
}

int orc_intersect(void const* spheres, int n, double const* origin, double const* direction, double* t_out)
{
    stats st;
    memset(&st, 0, sizeof(st));
    ray const r = { v3_make(origin[0], origin[1], origin[2]), v3_make(direction[0], direction[1], direction[2]) };
    double t = 0.0;
    int id = 0;
    int const hit = scene_intersect((orc_sphere const*)spheres, n, &r, &t, &id, &st);
    *t_out = t;
    return hit ? id : -1;
}

static inline uint32_t slot_of(int x, int y, int sx, int sy, int width, int ns)
{
    return (((uint32_t)y * (uint32_t)width + (uint32_t)x) * (uint32_t)ns + (uint32_t)sy) * (uint32_t)ns + (uint32_t)sx;
}

/* ------------------------------------------------------------------ entry points */
void orc_samples(void const* spheres, int n, void const* camera, int width, int height, int num_subpixels,
                 uint64_t seed, uint32_t const* xs, uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys,
                 uint32_t const* samples, int count, int32_t* primary_hit, double* radiance_out, double* ray_out,
                 uint64_t* draws_out)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    orc_camera cam;
    memcpy(&cam, camera, sizeof(cam));

#pragma omp parallel
    {
        stats st;
        memset(&st, 0, sizeof(st));
        rng g;
        memset(&g, 0, sizeof(g));
#pragma omp for schedule(static)
        for(int i = 0; i < count; ++i) {
            ptb_rng_key(&g.ctr, seed, slot_of((int)xs[i], (int)ys[i], (int)sxs[i], (int)sys[i], width, num_subpixels),
                        samples[i]);
            g.draws = 0;
            ray const pr = primary_ray(&cam, (int)xs[i], (int)ys[i], (int)sxs[i], (int)sys[i], width, height,
                                       num_subpixels, &g);
            double t = 0.0;
            int id = 0;
            stats dummy;
            memset(&dummy, 0, sizeof(dummy));
            primary_hit[i] = scene_intersect(sph, n, &pr, &t, &id, &dummy) ? id : -1;

            st.c[ORC_STAT_PATHS]++;
            v3 const c = radiance(sph, n, &pr, &g, &st);
            radiance_out[3 * i + 0] = c.x;
            radiance_out[3 * i + 1] = c.y;
            radiance_out[3 * i + 2] = c.z;
            if(ray_out) {
                ray_out[6 * i + 0] = pr.o.x;
                ray_out[6 * i + 1] = pr.o.y;
                ray_out[6 * i + 2] = pr.o.z;
                ray_out[6 * i + 3] = pr.d.x;
                ray_out[6 * i + 4] = pr.d.y;
                ray_out[6 * i + 5] = pr.d.z;
            }
            if(draws_out) {
                draws_out[i] = g.draws;
            }
            st.c[ORC_STAT_DRAWS] += g.draws;
        }
        stats_merge(&st);
    }
}

/* Which sphere each sample's path hits at depth 0 .. trail_len-1 (-1 = sky, -2 = path over): the same calls as
 * orc_samples, recording object_index of src/main.cpp:114.  For classifying samples that differ from the CUDA path. */
void orc_trails(void const* spheres, int n, void const* camera, int width, int height, int num_subpixels, uint64_t seed,
                uint32_t const* xs, uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys, uint32_t const* samples,
                int count, int trail_len, int32_t* trail_out)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    orc_camera cam;
    memcpy(&cam, camera, sizeof(cam));
#pragma omp parallel
    {
        stats st;
        rng g;
        memset(&g, 0, sizeof(g));
#pragma omp for schedule(static)
        for(int i = 0; i < count; ++i) {
            memset(&st, 0, sizeof(st));
            st.trail = trail_out + (size_t)i * (size_t)trail_len;
            st.trail_len = trail_len;
            for(int d = 0; d < trail_len; ++d) {
                st.trail[d] = -2;
            }
            ptb_rng_key(&g.ctr, seed, slot_of((int)xs[i], (int)ys[i], (int)sxs[i], (int)sys[i], width, num_subpixels),
                        samples[i]);
            g.draws = 0;
            ray const pr = primary_ray(&cam, (int)xs[i], (int)ys[i], (int)sxs[i], (int)sys[i], width, height,
                                       num_subpixels, &g);
            (void)radiance(sph, n, &pr, &g, &st);
        }
    }
}

/* Loop nest src/main.cpp:217-232; sample loop, clamp and accumulate main.cpp:184-196 */
void orc_render(void const* spheres, int n, void const* camera, int width, int height, int samps,
                int num_subpixels, uint64_t seed, uint32_t first_sample, double* image_out, double* sums_out,
                int nthreads)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    orc_camera cam;
    memcpy(&cam, camera, sizeof(cam));
    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();

#pragma omp parallel num_threads(threads)
    {
        stats st;
        memset(&st, 0, sizeof(st));
        rng g;
        memset(&g, 0, sizeof(g));
#pragma omp for schedule(dynamic, 1)
        for(int y = 0; y < height; y++) {
            for(int x = 0; x < width; x++) {
                size_t const row = (size_t)((height - y - 1) * width) + (size_t)x;
                v3 pixel = v3_make(0, 0, 0);
                for(int sy = 0; sy < num_subpixels; sy++) {
                    for(int sx = 0; sx < num_subpixels; sx++) {
                        uint32_t const slot = slot_of(x, y, sx, sy, width, num_subpixels);
                        v3 r = v3_make(0, 0, 0);
                        v3 sum = v3_make(0, 0, 0);
                        for(int s = 0; s < samps; s++) {
                            ptb_rng_key(&g.ctr, seed, slot, first_sample + (uint32_t)s);
                            g.draws = 0;
                            ray const pr = primary_ray(&cam, x, y, sx, sy, width, height, num_subpixels, &g);
                            st.c[ORC_STAT_PATHS]++;
                            v3 const c = radiance(sph, n, &pr, &g, &st);
                            st.c[ORC_STAT_DRAWS] += g.draws;
                            r = v3_add(r, v3_scale(c, 1.0 / samps));
                            sum = v3_add(sum, c);
                        }
                        v3 const sub = v3_make(clamp01(r.x), clamp01(r.y), clamp01(r.z));
                        pixel = v3_add(pixel, v3_scale(sub, 1.0 / (num_subpixels * num_subpixels)));
                        if(sums_out) {
                            sums_out[3 * (size_t)slot + 0] = sum.x;
                            sums_out[3 * (size_t)slot + 1] = sum.y;
                            sums_out[3 * (size_t)slot + 2] = sum.z;
                        }
                    }
                }
                image_out[3 * row + 0] = pixel.x;
                image_out[3 * row + 1] = pixel.y;
                image_out[3 * row + 2] = pixel.z;
            }
        }
        stats_merge(&st);
    }
}

/* The stock program's structure: one sequential mt19937 stream per image row
 * (src/main.cpp:217-233), render_subpixel per stratum (main.cpp:179-197). */
void orc_mt_render(void const* spheres, int n, void const* camera, int width, int height, int samps,
                   int num_subpixels, int seed_mode, uint32_t const* row_seed, int y0, int y1, double* image_out,
                   int nthreads)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    orc_camera cam;
    memcpy(&cam, camera, sizeof(cam));
    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();

#pragma omp parallel num_threads(threads)
    {
        stats st;
        memset(&st, 0, sizeof(st));
        rng* g = (rng*)calloc(1, sizeof(rng));
        g->use_mt = 1;
#pragma omp for schedule(dynamic, 1)
        for(int y = y0; y < y1; y++) {
            mt_seed(&g->mt, seed_mode == 2 ? row_seed[y] : 0u);
            for(int x = 0; x < width; x++) {
                size_t const row = (size_t)((height - y - 1) * width) + (size_t)x;
                for(int sy = 0; sy < num_subpixels; sy++) {
                    for(int sx = 0; sx < num_subpixels; sx++) {
                        v3 r = v3_make(0, 0, 0);
                        for(int s = 0; s < samps; s++) {
                            ray const pr = primary_ray(&cam, x, y, sx, sy, width, height, num_subpixels, g);
                            st.c[ORC_STAT_PATHS]++;
                            v3 const c = radiance(sph, n, &pr, g, &st);
                            r = v3_add(r, v3_scale(c, 1.0 / samps));
                        }
                        v3 const sub = v3_make(clamp01(r.x), clamp01(r.y), clamp01(r.z));
                        v3 const add = v3_scale(sub, 1.0 / (num_subpixels * num_subpixels));
                        image_out[3 * row + 0] = image_out[3 * row + 0] + add.x;
                        image_out[3 * row + 1] = image_out[3 * row + 1] + add.y;
                        image_out[3 * row + 2] = image_out[3 * row + 2] + add.z;
                    }
                }
            }
        }
        st.c[ORC_STAT_DRAWS] += g->draws;
        free(g);
        stats_merge(&st);
    }
}
