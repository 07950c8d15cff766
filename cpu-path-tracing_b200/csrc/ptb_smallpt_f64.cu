// ptb_smallpt_f64.cu -- deterministic parity kernels for the reference's stand-alone smallpt fork,
// /root/reference/sandbox/main.cpp.  FP64, compiled with -fmad=false,
// reference operation order inside every expression.
//
// What differs from the src/ integrator (ptb_f64.cu), with the sandbox lines:
//   Sphere::intersect :75-92   unit directions, det = b*b - op.op + r*r, roots compared with '> eps'
//   intersect :135-147         REVERSE index order, strict '<' (the highest index wins ties)
//   radiance :149-227          black on a miss; Russian roulette after depth 5 that returns the
//                              emission; glass IOR 1.5; while depth <= 2 a glass hit SPLITS into
//                              reflection and refraction (the pinned build traces the refraction
//                              first); no depth limit
//   camera :235-261            pinhole, tent filter over the 2x2 sub-pixels, ray pushed 140 units
//                              along the unit direction
// The recursion is unrolled into a loop with a two-entry stack (a split can only happen at
// depth 1 and 2).  Radiance is accumulated forward, L += W * e, instead of bottom-up
// e + f * (child): the same sum in a different rounding order, hence a tolerance (1e-4 relative
// per north star; observed ~1e-15) instead of bit equality on radiance.  Hit indices and camera
// rays are bit-exact.
#include "ptb_kernels.h"
#include "ptb_rng.cuh"

namespace ptb {

namespace {

constexpr double kPi = 3.14159265358979323846; // M_PI
constexpr int kSafetyBounces = 1 << 20;        // the sandbox has no depth limit; this only guards against a hang

struct V3
{
    double x, y, z;
};
__device__ __forceinline__ V3 mk(double x, double y, double z)
{
    V3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
__device__ __forceinline__ V3 operator+(V3 a, V3 b)
{
    return mk(a.x + b.x, a.y + b.y, a.z + b.z);
}
__device__ __forceinline__ V3 operator-(V3 a, V3 b)
{
    return mk(a.x - b.x, a.y - b.y, a.z - b.z);
}
__device__ __forceinline__ V3 operator*(V3 a, double b)
{
    return mk(a.x * b, a.y * b, a.z * b);
}
__device__ __forceinline__ V3 mult(V3 a, V3 b)
{
    return mk(a.x * b.x, a.y * b.y, a.z * b.z);
}
__device__ __forceinline__ double dot(V3 a, V3 b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
__device__ __forceinline__ V3 norm(V3 a)
{
    return a * (1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z));
}
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct Ray64
{
    V3 o, d;
};
struct Rng64
{
    Rng g;
    uint32_t draws;
};
__device__ __forceinline__ double rnd(Rng64& r)
{
    r.draws++;
    return rng_uniform_f64(r.g);
}

__device__ __forceinline__ double sphere_intersect(RawSphere const& s, Ray64 const& r)
{
    V3 const op = mk(s.px, s.py, s.pz) - r.o;
    double t;
    double const eps = 1e-4;
    double const b = dot(op, r.d);
    double det = b * b - dot(op, op) + s.radius * s.radius;
    if(det < 0) {
        return 0;
    }
    det = sqrt(det);
    return (t = b - det) > eps ? t : ((t = b + det) > eps ? t : 0);
}

__device__ __forceinline__ bool scene_intersect(RawSphere const* __restrict__ sph, int n, Ray64 const& r, double& t, int& id)
{
    double d;
    double const inf = t = 1e20;
    for(int i = n; i--;) {
        if((d = sphere_intersect(sph[i], r)) != 0 && d < t) {
            t = d;
            id = i;
        }
    }
    return t < inf;
}

struct Pending
{
    Ray64 r;
    V3 w;
    int depth;
};

struct Counters64
{
    uint32_t rays, diffuse, specular, dielectric;
};

__device__ V3 radiance(RawSphere const* __restrict__ sph, int n, Ray64 r, Rng64& rng, Counters64& cnt)
{
    V3 L = mk(0, 0, 0);
    V3 W = mk(1, 1, 1);
    int depth = 0;
    Pending stack[2];
    int sp = 0;
    for(int guard = 0; guard < kSafetyBounces; ++guard) {
        double t;
        int id = 0;
        cnt.rays++;
        bool segment_done = false;
        if(!scene_intersect(sph, n, r, t, id)) {
            segment_done = true; // black on a miss
        }
        else {
            RawSphere const obj = sph[id];
            V3 const e = mk(obj.er, obj.eg, obj.eb);
            V3 const x = r.o + r.d * t;
            V3 const nn = norm(x - mk(obj.px, obj.py, obj.pz));
            V3 const nl = dot(nn, r.d) < 0 ? nn : nn * -1.0;
            V3 f = mk(obj.cr, obj.cg, obj.cb);
            double const p = fmax(fmax(f.x, f.y), f.z);
            L = L + mult(W, e);
            bool killed = false;
            if(++depth > 5) {
                if(rnd(rng) < p) {
                    f = f * (1.0 / p);
                }
                else {
                    killed = true; // returns obj.e: already added
                }
            }
            if(killed) {
                segment_done = true;
            }
            else if(obj.reflection == 0) {
                cnt.diffuse++;
                double const r1 = 2 * kPi * rnd(rng);
                double const r2 = rnd(rng);
                double const r2s = sqrt(r2);
                V3 const w = nl;
                V3 const u = norm(cross(fabs(w.x) > .1 ? mk(0, 1, 0) : mk(1, 0, 0), w));
                V3 const v = cross(w, u);
                V3 const d = norm((u * cos(r1)) * r2s + (v * sin(r1)) * r2s + w * sqrt(1 - r2));
                r.o = x;
                r.d = d;
                W = mult(W, f);
            }
            else if(obj.reflection == 1) {
                cnt.specular++;
                r.d = r.d - (nn * 2.0) * dot(nn, r.d);
                r.o = x;
                W = mult(W, f);
            }
            else {
                cnt.dielectric++;
                Ray64 refl;
                refl.o = x;
                refl.d = r.d - (nn * 2.0) * dot(nn, r.d);
                bool const into = dot(nn, nl) > 0;
                double const nc = 1;
                double const nt = 1.5;
                double const nnt = into ? nc / nt : nt / nc;
                double const ddn = dot(r.d, nl);
                double const cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
                if(cos2t < 0) {
                    r = refl;
                    W = mult(W, f);
                }
                else {
                    V3 const tdir = norm(r.d * nnt - nn * ((into ? 1 : -1) * (ddn * nnt + sqrt(cos2t))));
                    double const a = nt - nc, b = nt + nc;
                    double const R0 = a * a / (b * b);
                    double const c = 1 - (into ? -ddn : dot(tdir, nn));
                    double const Re = R0 + (1 - R0) * c * c * c * c * c;
                    double const Tr = 1 - Re;
                    double const P = .25 + .5 * Re;
                    double const RP = Re / P;
                    double const TP = Tr / (1 - P);
                    Ray64 refr;
                    refr.o = x;
                    refr.d = tdir;
                    V3 const wf = mult(W, f);
                    if(depth > 2) {
                        if(rnd(rng) < P) {
                            r = refl;
                            W = wf * RP;
                        }
                        else {
                            r = refr;
                            W = wf * TP;
                        }
                    }
                    else {
                        // split: the refraction subtree is traced first, the reflection waits on the stack
                        stack[sp].r = refl;
                        stack[sp].w = wf * Re;
                        stack[sp].depth = depth;
                        ++sp;
                        r = refr;
                        W = wf * Tr;
                    }
                }
            }
        }
        if(segment_done) {
            if(sp == 0) {
                break;
            }
            --sp;
            r = stack[sp].r;
            W = stack[sp].w;
            depth = stack[sp].depth;
        }
    }
    return L;
}

struct Cam64
{
    V3 o, d, cx, cy;
    double push;
};

__device__ __forceinline__ Cam64 make_camera(double const* cam8, uint32_t w, uint32_t h)
{
    Cam64 c;
    c.o = mk(cam8[0], cam8[1], cam8[2]);
    c.d = norm(mk(cam8[3], cam8[4], cam8[5]));
    c.cx = mk(static_cast<int>(w) * cam8[6] / static_cast<int>(h), 0, 0);
    c.cy = norm(cross(c.cx, c.d)) * cam8[6];
    c.push = cam8[7];
    return c;
}

__device__ __forceinline__ Ray64 camera_ray(Cam64 const& c, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy, uint32_t w,
                                            uint32_t h, Rng64& rng)
{
    double const r1 = 2 * rnd(rng);
    double const dx = r1 < 1 ? sqrt(r1) - 1 : 1 - sqrt(2 - r1);
    double const r2 = 2 * rnd(rng);
    double const dy = r2 < 1 ? sqrt(r2) - 1 : 1 - sqrt(2 - r2);
    V3 d = c.cx * (((static_cast<int>(sx) + .5 + dx) / 2 + static_cast<int>(x)) / static_cast<int>(w) - .5) +
           c.cy * (((static_cast<int>(sy) + .5 + dy) / 2 + static_cast<int>(y)) / static_cast<int>(h) - .5) + c.d;
    d = norm(d); // d.norm() runs before cam.o + d * 140 in the pinned build (argument evaluation order)
    Ray64 r;
    r.o = c.o + d * c.push;
    r.d = d;
    return r;
}

__global__ void smallpt_probe_f64_kernel(ProbeParams const q, RawSphere const* __restrict__ sph, int n,
                                         double const* __restrict__ cam8)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= q.count) {
        return;
    }
    Cam64 const cam = make_camera(cam8, q.width, q.height);
    uint32_t const x = q.x[i], y = q.y[i], sx = q.sx[i], sy = q.sy[i];
    uint32_t const slot = ((y * q.width + x) * 2u + sy) * 2u + sx;
    Rng64 rng;
    rng.g = rng_open(q.key, slot, q.sample[i]);
    rng.draws = 0;
    Ray64 const pr = camera_ray(cam, x, y, sx, sy, q.width, q.height, rng);
    if(q.ray != nullptr) {
        q.ray[6 * i + 0] = pr.o.x;
        q.ray[6 * i + 1] = pr.o.y;
        q.ray[6 * i + 2] = pr.o.z;
        q.ray[6 * i + 3] = pr.d.x;
        q.ray[6 * i + 4] = pr.d.y;
        q.ray[6 * i + 5] = pr.d.z;
    }
    double t;
    int id = 0;
    q.primary_hit[i] = scene_intersect(sph, n, pr, t, id) ? id : -1;
    Counters64 cnt{ 0, 0, 0, 0 };
    V3 const L = radiance(sph, n, pr, rng, cnt);
    q.radiance[3 * i + 0] = L.x;
    q.radiance[3 * i + 1] = L.y;
    q.radiance[3 * i + 2] = L.z;
    if(q.draws != nullptr) {
        q.draws[i] = rng.draws;
    }
}

__global__ void smallpt_render_f64_kernel(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width,
                                          uint32_t height, RawSphere const* __restrict__ sph, int n,
                                          double const* __restrict__ cam8, double* __restrict__ accum,
                                          DeviceCounters* counters)
{
    uint32_t const slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t const nslots = width * height * 4u;
    Counters64 cnt{ 0, 0, 0, 0 };
    if(slot < nslots) {
        Cam64 const cam = make_camera(cam8, width, height);
        uint32_t const sx = slot & 1u, sy = (slot >> 1) & 1u, pix = slot >> 2;
        uint32_t const y = pix / width, x = pix - y * width;
        double* a = accum + 4 * static_cast<size_t>(slot);
        V3 sum = mk(a[0], a[1], a[2]); // continues the slot's running sum: split or resumed renders add in the same order
        for(uint32_t s = 0; s < samples; ++s) {
            Rng64 rng;
            rng.g = rng_open(key, slot, first_sample + s);
            rng.draws = 0;
            Ray64 const pr = camera_ray(cam, x, y, sx, sy, width, height, rng);
            sum = sum + radiance(sph, n, pr, rng, cnt);
        }
        a[0] = sum.x;
        a[1] = sum.y;
        a[2] = sum.z;
        a[3] += static_cast<double>(samples);
    }
    uint32_t const rays = __reduce_add_sync(0xffffffffu, cnt.rays);
    uint32_t const nd = __reduce_add_sync(0xffffffffu, cnt.diffuse);
    uint32_t const nsp = __reduce_add_sync(0xffffffffu, cnt.specular);
    uint32_t const ndi = __reduce_add_sync(0xffffffffu, cnt.dielectric);
    if((threadIdx.x & 31u) == 0u) {
        atomicAdd(&counters->rays, static_cast<unsigned long long>(rays));
        atomicAdd(&counters->diffuse, static_cast<unsigned long long>(nd));
        atomicAdd(&counters->specular, static_cast<unsigned long long>(nsp));
        atomicAdd(&counters->dielectric, static_cast<unsigned long long>(ndi));
    }
}

} // namespace

cudaError_t launch_smallpt_probe_f64(ProbeParams const& p, RawSphere const* spheres, int n, double const* cam8,
                                     cudaStream_t stream)
{
    if(p.count == 0) {
        return cudaSuccess;
    }
    unsigned const threads = 128;
    smallpt_probe_f64_kernel<<<(p.count + threads - 1) / threads, threads, 0, stream>>>(p, spheres, n, cam8);
    return cudaGetLastError();
}

cudaError_t launch_smallpt_render_f64(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width, uint32_t height,
                                      RawSphere const* spheres, int n, double const* cam8, double* accum64,
                                      DeviceCounters* counters, cudaStream_t stream)
{
    uint32_t const nslots = width * height * 4u;
    unsigned const threads = 128;
    smallpt_render_f64_kernel<<<(nslots + threads - 1) / threads, threads, 0, stream>>>(key, first_sample, samples, width,
                                                                                          height, spheres, n, cam8, accum64,
                                                                                          counters);
    return cudaGetLastError();
}

} // namespace ptb
