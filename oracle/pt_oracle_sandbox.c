/* pt_oracle_sandbox.c -- plain-C, FP64 restatement of the reference's stand-alone smallpt fork,
 * /root/reference/sandbox/main.cpp: reverse-order closest hit,
 * unit directions, recursive radiance with Russian roulette after depth 5 returning the emission,
 * glass (IOR 1.5) that SPLITS into reflection + refraction while depth <= 2, black on a miss,
 * tent-filtered pinhole camera pushed 140 units forward.
 *
 * TEST INFRASTRUCTURE (see pt_oracle.h).  Pinned against the reference itself: bit-identical to
 * oracle/_ref/libsbref.so (the sandbox's own object code) in both streams, and to the image.ppm the
 * sandbox PROGRAM writes (oracle/_ref/smallpt, built as sandbox/run.sh:3 says) -- the program is
 * deterministic: erand48 seeded {0, 0, (unsigned short)(y^3)} per row (sandbox/main.cpp:245).
 *
 * Two expressions of the sandbox have an evaluation order C++ leaves unspecified and that CHANGES
 * the result; the orders below are what g++ 13 -O3 produced for the pinned build (verified by the
 * bit-exact tests, tests/test_sandbox.py):
 *   Ray(cam.o + d * 140, d.norm())                      -> d.norm() runs FIRST (arguments right to left),
 *                                                          so the origin is pushed along the UNIT direction
 *   radiance(reflRay) * Re + radiance(Ray(x, tdir)) * Tr -> the RIGHT operand (refraction) is traced first,
 *                                                          so it consumes the stream before the reflection
 */
#include "pt_oracle.h"
#include "ptb_rng.h"

#include <math.h>
#include <omp.h>
#include <string.h>

typedef orc_vec3 v3;

static inline v3 mk(double x, double y, double z)
{
    v3 r = { x, y, z };
    return r;
}
static inline v3 add(v3 a, v3 b)
{
    return mk(a.x + b.x, a.y + b.y, a.z + b.z);
}
static inline v3 sub(v3 a, v3 b)
{
    return mk(a.x - b.x, a.y - b.y, a.z - b.z);
}
static inline v3 scale(v3 a, double b)
{
    return mk(a.x * b, a.y * b, a.z * b);
}
static inline v3 mult(v3 a, v3 b)
{
    return mk(a.x * b.x, a.y * b.y, a.z * b.z);
}
static inline v3 norm(v3 a) /* sandbox/main.cpp:32-35 */
{
    return scale(a, 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z));
}
static inline double dot(v3 a, v3 b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
static inline v3 cross(v3 a, v3 b) /* operator%, sandbox/main.cpp:40-43 */
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

typedef struct sb_rng
{
    int counter_mode;
    ptb_rng ctr;
    uint64_t x48; /* erand48 state */
    uint64_t draws;
} sb_rng;

/* libc erand48: X <- (0x5DEECE66D * X + 0xB) mod 2^48, result X / 2^48 */
static inline double sb_rand(sb_rng* g)
{
    g->draws++;
    if(g->counter_mode) {
        return ptb_rng_uniform(&g->ctr);
    }
    g->x48 = (0x5DEECE66Dull * g->x48 + 0xBull) & 0xFFFFFFFFFFFFull;
    return ldexp((double)g->x48, -48);
}

typedef struct sb_ray
{
    v3 o, d;
} sb_ray;

/* Sphere::intersect, sandbox/main.cpp:75-92 */
static inline double sb_sphere_intersect(orc_sphere const* s, sb_ray const* r)
{
    v3 const op = sub(s->position, r->o);
    double t;
    double const eps = 1e-4;
    double const b = dot(op, r->d);
    double det = b * b - dot(op, op) + s->radius * s->radius;
    if(det < 0) {
        return 0;
    }
    det = sqrt(det);
    return (t = b - det) > eps ? t : ((t = b + det) > eps ? t : 0);
}

/* intersect, sandbox/main.cpp:135-147: REVERSE index order, strict '<' */
static inline int sb_intersect(orc_sphere const* sph, int n, sb_ray const* r, double* t, int* id)
{
    double d;
    double const inf = *t = 1e20;
    for(int i = n; i--;) {
        if((d = sb_sphere_intersect(&sph[i], r)) != 0 && d < *t) {
            *t = d;
            *id = i;
        }
    }
    return *t < inf;
}

#ifndef SB_SPLIT_REFLECT_FIRST
#define SB_SPLIT_REFLECT_FIRST 0 /* the pinned build evaluates the RIGHT operand (refraction) first */
#endif

/* radiance, sandbox/main.cpp:149-227 (recursive, like the reference) */
static v3 sb_radiance(orc_sphere const* sph, int n, sb_ray const* r, int depth, sb_rng* g)
{
    double t;
    int id = 0;
    if(!sb_intersect(sph, n, r, &t, &id)) {
        return mk(0, 0, 0);
    }
    orc_sphere const* obj = &sph[id];
    v3 const x = add(r->o, scale(r->d, t));
    v3 const nn = norm(sub(x, obj->position));
    v3 const nl = dot(nn, r->d) < 0 ? nn : scale(nn, -1);
    v3 f = obj->color;
    double const p = fmax(fmax(f.x, f.y), f.z);

    if(++depth > 5) {
        if(sb_rand(g) < p) {
            f = scale(f, 1.0 / p);
        }
        else {
            return obj->emission;
        }
    }

    if(obj->reflection == 0) { /* DIFF */
        double const r1 = 2 * M_PI * sb_rand(g);
        double const r2 = sb_rand(g);
        double const r2s = sqrt(r2);
        v3 const w = nl;
        v3 const u = norm(cross(fabs(w.x) > .1 ? mk(0, 1, 0) : mk(1, 0, 0), w));
        v3 const v = cross(w, u);
        v3 const d = norm(add(add(scale(scale(u, cos(r1)), r2s), scale(scale(v, sin(r1)), r2s)), scale(w, sqrt(1 - r2))));
        sb_ray const nr = { x, d };
        return add(obj->emission, mult(f, sb_radiance(sph, n, &nr, depth, g)));
    }
    else if(obj->reflection == 1) { /* SPEC */
        sb_ray const nr = { x, sub(r->d, scale(scale(nn, 2), dot(nn, r->d))) };
        return add(obj->emission, mult(f, sb_radiance(sph, n, &nr, depth, g)));
    }

    sb_ray const refl_ray = { x, sub(r->d, scale(scale(nn, 2), dot(nn, r->d))) };
    int const into = dot(nn, nl) > 0;
    double const nc = 1;
    double const nt = 1.5;
    double const nnt = into ? nc / nt : nt / nc;
    double const ddn = dot(r->d, nl);
    double cos2t;
    if((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0) {
        return add(obj->emission, mult(f, sb_radiance(sph, n, &refl_ray, depth, g)));
    }
    v3 const tdir = norm(sub(scale(r->d, nnt), scale(nn, (into ? 1 : -1) * (ddn * nnt + sqrt(cos2t)))));
    double const a = nt - nc, b = nt + nc;
    double const R0 = a * a / (b * b);
    double const c = 1 - (into ? -ddn : dot(tdir, nn));
    double const Re = R0 + (1 - R0) * c * c * c * c * c;
    double const Tr = 1 - Re;
    double const P = .25 + .5 * Re;
    double const RP = Re / P;
    double const TP = Tr / (1 - P);
    sb_ray const refr_ray = { x, tdir };
    v3 inner;
    if(depth > 2) {
        if(sb_rand(g) < P) {
            inner = scale(sb_radiance(sph, n, &refl_ray, depth, g), RP);
        }
        else {
            inner = scale(sb_radiance(sph, n, &refr_ray, depth, g), TP);
        }
    }
    else {
#if SB_SPLIT_REFLECT_FIRST
        v3 const a1 = scale(sb_radiance(sph, n, &refl_ray, depth, g), Re);
        v3 const a2 = scale(sb_radiance(sph, n, &refr_ray, depth, g), Tr);
#else
        v3 const a2 = scale(sb_radiance(sph, n, &refr_ray, depth, g), Tr);
        v3 const a1 = scale(sb_radiance(sph, n, &refl_ray, depth, g), Re);
#endif
        inner = add(a1, a2);
    }
    return add(obj->emission, mult(f, inner));
}

typedef struct sb_camera
{
    v3 o, d, cx, cy;
    double push;
} sb_camera;

/* sandbox/main.cpp:235-237; cam8 = position, direction (un-normalised), fov factor, push */
static sb_camera sb_make_camera(double const* cam8, int w, int h)
{
    sb_camera c;
    c.o = mk(cam8[0], cam8[1], cam8[2]);
    c.d = norm(mk(cam8[3], cam8[4], cam8[5]));
    c.cx = mk(w * cam8[6] / h, 0, 0);
    c.cy = scale(norm(cross(c.cx, c.d)), cam8[6]);
    c.push = cam8[7];
    return c;
}

/* sandbox/main.cpp:253-261 */
static inline sb_ray sb_camera_ray(sb_camera const* c, int x, int y, int sx, int sy, int w, int h, sb_rng* g)
{
    double const r1 = 2 * sb_rand(g);
    double const dx = r1 < 1 ? sqrt(r1) - 1 : 1 - sqrt(2 - r1);
    double const r2 = 2 * sb_rand(g);
    double const dy = r2 < 1 ? sqrt(r2) - 1 : 1 - sqrt(2 - r2);
    v3 d = add(add(scale(c->cx, ((sx + .5 + dx) / 2 + x) / w - .5), scale(c->cy, ((sy + .5 + dy) / 2 + y) / h - .5)), c->d);
#ifndef SB_CAM_NORM_FIRST
#define SB_CAM_NORM_FIRST 1
#endif
#if SB_CAM_NORM_FIRST
    d = norm(d); /* d.norm() is evaluated before cam.o + d * 140 in the pinned build */
    sb_ray r = { add(c->o, scale(d, c->push)), d };
#else
    sb_ray r = { add(c->o, scale(d, c->push)), norm(d) };
#endif
    return r;
}

static inline double sb_clamp(double x)
{
    return x < 0 ? 0 : x > 1 ? 1 : x;
}

void orc_sb_to_int(double const* v, int n, int* out) /* toInt, sandbox/main.cpp:130-133 */
{
    for(int i = 0; i < n; ++i) {
        out[i] = (int)(pow(sb_clamp(v[i]), 1 / 2.2) * 255 + .5);
    }
}

static inline uint32_t sb_slot(int x, int y, int sx, int sy, int w)
{
    return (((uint32_t)y * (uint32_t)w + (uint32_t)x) * 2u + (uint32_t)sy) * 2u + (uint32_t)sx;
}

/* loop nest of sandbox/main.cpp:241-269; mode 0 = the program's erand48 stream, 1 = counter stream */
void orc_sb_render(void const* spheres, int n, double const* cam8, int w, int h, int samps, int mode, uint64_t seed,
                   uint32_t first_sample, int y0, int y1, double* image_out, int nthreads)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    sb_camera const cam = sb_make_camera(cam8, w, h);
    int const threads = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for(int y = y0; y < y1; y++) {
        sb_rng g;
        memset(&g, 0, sizeof(g));
        g.counter_mode = mode == 1;
        g.x48 = (uint64_t)(unsigned short)(y * y * y) << 32; /* Xi = {0, 0, y^3}: x[2] is the high word */
        for(int x = 0; x < w; x++) {
            size_t const i = (size_t)(h - y - 1) * (size_t)w + (size_t)x;
            for(int sy = 0; sy < 2; sy++) {
                for(int sx = 0; sx < 2; sx++) {
                    v3 r = mk(0, 0, 0);
                    for(int s = 0; s < samps; s++) {
                        if(mode == 1) {
                            ptb_rng_key(&g.ctr, seed, sb_slot(x, y, sx, sy, w), first_sample + (uint32_t)s);
                        }
                        sb_ray const ray = sb_camera_ray(&cam, x, y, sx, sy, w, h, &g);
                        r = add(r, scale(sb_radiance(sph, n, &ray, 0, &g), 1. / samps));
                    }
                    v3 const add25 = scale(mk(sb_clamp(r.x), sb_clamp(r.y), sb_clamp(r.z)), .25);
                    image_out[3 * i + 0] = image_out[3 * i + 0] + add25.x;
                    image_out[3 * i + 1] = image_out[3 * i + 1] + add25.y;
                    image_out[3 * i + 2] = image_out[3 * i + 2] + add25.z;
                }
            }
        }
    }
}

void orc_sb_samples(void const* spheres, int n, double const* cam8, int w, int h, uint64_t seed, uint32_t const* xs,
                    uint32_t const* ys, uint32_t const* sxs, uint32_t const* sys, uint32_t const* samples, int count,
                    int32_t* primary_hit, double* radiance_out, double* ray_out, uint64_t* draws_out)
{
    orc_sphere const* sph = (orc_sphere const*)spheres;
    sb_camera const cam = sb_make_camera(cam8, w, h);
#pragma omp parallel for schedule(static)
    for(int k = 0; k < count; ++k) {
        sb_rng g;
        memset(&g, 0, sizeof(g));
        g.counter_mode = 1;
        ptb_rng_key(&g.ctr, seed, sb_slot((int)xs[k], (int)ys[k], (int)sxs[k], (int)sys[k], w), samples[k]);
        sb_ray const ray = sb_camera_ray(&cam, (int)xs[k], (int)ys[k], (int)sxs[k], (int)sys[k], w, h, &g);
        double t;
        int id = 0;
        primary_hit[k] = sb_intersect(sph, n, &ray, &t, &id) ? id : -1;
        v3 const L = sb_radiance(sph, n, &ray, 0, &g);
        radiance_out[3 * k + 0] = L.x;
        radiance_out[3 * k + 1] = L.y;
        radiance_out[3 * k + 2] = L.z;
        if(ray_out) {
            ray_out[6 * k + 0] = ray.o.x;
            ray_out[6 * k + 1] = ray.o.y;
            ray_out[6 * k + 2] = ray.o.z;
            ray_out[6 * k + 3] = ray.d.x;
            ray_out[6 * k + 4] = ray.d.y;
            ray_out[6 * k + 5] = ray.d.z;
        }
        if(draws_out) {
            draws_out[k] = g.draws;
        }
    }
}
