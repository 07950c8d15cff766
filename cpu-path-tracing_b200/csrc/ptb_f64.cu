// ptb_f64.cu -- deterministic parity kernels: FP64, reference operation order.
//
// THIS FILE IS COMPILED WITH -fmad=false: no multiply-add contraction, so +,-,*,/
// and sqrt round exactly as the reference's x86-64 build does (no FMA there either:
// the reference adds no -march, CMakeLists.txt:7-8).  Paths that touch only those
// operations (camera rays, sphere hits, mirror bounces, Russian roulette) come out
// BIT-IDENTICAL to the reference; diffuse bounces go through sin/cos and dielectric
// ones through pow, where CUDA's libm and glibc differ by an ulp or two
// (SURVEY.md section 7), hence the 1e-4 relative bar on radiance and the exact bar on
// primary-hit indices.
//
// Reads the reference's own AoS records (pt::sphere 88 B, pt::camera 176 B) directly.
// Every function follows the reference line by line:
//   sphere_intersect  src/sphere.cpp:6-30        scene_intersect  src/main.cpp:30-42
//   hit record        src/hit_record.cpp:3-12    diffuse_ray      src/main.cpp:44-58
//   specular_ray      src/main.cpp:60-67         dielectric_ray   src/main.cpp:69-97
//   radiance          src/main.cpp:104-158       get_ray          src/camera.cpp:19-38
//   sample loop       src/main.cpp:184-193
#include "ptb_kernels.h"
#include "ptb_rng.cuh"

namespace ptb {

namespace {

constexpr double kEps = 1e-4;                  // constants.hpp:7
constexpr double kPi = 3.14159265358979323846; // constants.hpp:8
constexpr double kInf = 1e20;                  // constants.hpp:9
constexpr int kDepthLimit64 = 100;             // constants.hpp:10

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
__device__ __forceinline__ V3 blend(V3 a, V3 b)
{
    return mk(a.x * b.x, a.y * b.y, a.z * b.z);
}
__device__ __forceinline__ double dot(V3 a, V3 b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z; // (x*x' + y*y') + z*z', vec.cpp:40-43
}
__device__ __forceinline__ V3 norm(V3 a)
{
    return a * (1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); // vec.cpp:35-38
}
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); // vec.cpp:45-48
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

__device__ __forceinline__ double gen(Rng64& r)
{
    r.draws++;
    return rng_uniform_f64(r.g);
}
__device__ __forceinline__ double gen_between(Rng64& r, double mn, double mx)
{
    return mn + (mx - mn) * gen(r); // random_state.cpp:14-17
}

__device__ __forceinline__ double sphere_intersect(RawSphere const& s, Ray64 const& r)
{
    V3 const oc = r.o - mk(s.px, s.py, s.pz);
    double const a = dot(r.d, r.d);
    double const half_b = dot(oc, r.d);
    double const c = dot(oc, oc) - s.radius * s.radius;
    double const discriminant = half_b * half_b - a * c;
    if(discriminant < 0) {
        return 0.0;
    }
    double const sqrtd = sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;
    if(root < kEps) {
        root = (-half_b + sqrtd) / a;
        if(root < kEps) {
            return 0.0;
        }
    }
    return root;
}

__device__ __forceinline__ bool scene_intersect(RawSphere const* __restrict__ sph, int n, Ray64 const& r, double& t,
                                                int& id)
{
    t = kInf;
    for(int i = 0; i < n; i++) {
        double const d = sphere_intersect(sph[i], r);
        if(d > 0 && d < t) {
            t = d;
            id = i;
        }
    }
    return t < kInf;
}

struct Hit64
{
    V3 point, outward, normal;
    bool front;
};

__device__ __forceinline__ Ray64 specular_ray(Hit64 const& h, V3 d, Rng64& rng)
{
    V3 const reflected = d - (h.outward * 2.0) * dot(h.outward, d);
    double const factor = gen(rng) * 0.0;
    Ray64 out;
    out.o = h.point;
    out.d = reflected + mk(factor, factor, factor);
    return out;
}

struct Counters64
{
    uint32_t rays, diffuse, specular, dielectric;
    int32_t* trail = nullptr; // parity tooling (ptb_trace_paths): sphere hit at each depth < trail_len, -1 = sky
    int trail_len = 0;
};

__device__ V3 radiance(RawSphere const* __restrict__ sph, int n, Ray64 r, Rng64& rng, Counters64& cnt)
{
    V3 E = mk(0.0, 0.0, 0.0);
    V3 T = mk(1, 1, 1);
    for(int depth = 0; depth < kDepthLimit64; ++depth) {
        double t = 0.0;
        int id = 0;
        cnt.rays++;
        bool const any_hit = scene_intersect(sph, n, r, t, id);
        if(depth < cnt.trail_len) {
            cnt.trail[depth] = any_hit ? id : -1;
        }
        if(!any_hit) {
            V3 const unit = norm(r.d);
            double const tt = 0.5 * (unit.y + 1.0);
            V3 const background = mk(1.0, 1.0, 1.0) * (1.0 - tt) + mk(0.5, 0.7, 1.0) * tt;
            return E + blend(T, background);
        }
        RawSphere const obj = sph[id];
        Hit64 h;
        h.point = r.o + r.d * t;
        h.outward = norm(h.point - mk(obj.px, obj.py, obj.pz));
        h.front = dot(h.outward, r.d) < 0;
        h.normal = h.front ? h.outward : h.outward * -1.0;
        V3 color = mk(obj.cr, obj.cg, obj.cb);

        E = E + blend(T, mk(obj.er, obj.eg, obj.eb));

        double const probability = fmax(fmax(color.x, color.y), color.z);
        if(depth > 4) {
            if(gen(rng) < probability) {
                color = color * (1.0 / probability);
            }
            else {
                return E;
            }
        }
        T = blend(T, color);

        if(obj.reflection == 0) {
            cnt.diffuse++;
            double const phi = 2 * kPi * gen(rng);
            double const random_angle = gen(rng);
            double const sin_theta = sqrt(random_angle);
            double const cos_theta = sqrt(1.0 - random_angle);
            V3 const w = h.normal;
            V3 const u = norm(cross(fabs(w.x) > 0.1 ? mk(0, 1, 0) : mk(1, 0, 0), w));
            V3 const v = cross(w, u);
            V3 const nd = norm(((u * cos(phi)) * sin_theta + (v * sin(phi)) * sin_theta) + w * cos_theta);
            r.o = h.point;
            r.d = nd;
        }
        else if(obj.reflection == 1) {
            cnt.specular++;
            r = specular_ray(h, r.d, rng);
        }
        else if(obj.reflection == 2) {
            cnt.dielectric++;
            double const ratio = h.front ? (1.0 / 2.0) : 2.0;
            V3 const unit = norm(r.d);
            double const cos_theta = fmin(dot(unit * -1.0, h.normal), 1.0);
            double const sin_theta = sqrt(1.0 - cos_theta * cos_theta);
            bool reflect = ratio * sin_theta > 1.0;
            if(!reflect) {
                double r0 = (1.0 - ratio) / (1.0 + ratio);
                r0 *= r0;
                double const refl = r0 + (1.0 - r0) * pow(1.0 - cos_theta, 5.0);
                reflect = refl > gen(rng);
            }
            if(reflect) {
                r = specular_ray(h, r.d, rng);
            }
            else {
                V3 const perp = (unit + h.normal * cos_theta) * ratio;
                V3 const par = h.normal * (-sqrt(fabs(1.0 - dot(perp, perp))));
                r.o = h.point;
                r.d = perp + par;
            }
        }
    }
    return E;
}

__device__ __forceinline__ Ray64 primary_ray(RawCamera const& cam, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy,
                                             uint32_t width, uint32_t height, uint32_t ns, Rng64& rng)
{
    double const len = 1.0 / static_cast<int>(ns);
    double const xs = (static_cast<int>(x) + static_cast<int>(sx) * len + gen_between(rng, 0.0, len));
    double const ys = (static_cast<int>(y) + static_cast<int>(sy) * len + gen_between(rng, 0.0, len));
    double const s = xs / static_cast<int>(width);
    double const t = ys / static_cast<int>(height);

    V3 point = mk(0, 0, 0);
    for(;;) {
        double const px = gen_between(rng, -1.0, 1.0);
        double const py = gen_between(rng, -1.0, 1.0);
        point = mk(px, py, 0.0);
        if(dot(point, point) >= 1.0) {
            continue;
        }
        break;
    }
    V3 const rd = point * cam.lens_radius;
    V3 const offset = rd * s + rd * t;
    V3 const pos = mk(cam.pos[0], cam.pos[1], cam.pos[2]);
    V3 const llc = mk(cam.llc[0], cam.llc[1], cam.llc[2]);
    V3 const ax = mk(cam.ax[0], cam.ax[1], cam.ax[2]);
    V3 const ay = mk(cam.ay[0], cam.ay[1], cam.ay[2]);
    Ray64 out;
    out.d = (((llc + ax * s) + ay * t) - pos) - offset;
    out.o = pos + offset;
    return out;
}

__global__ void probe_f64_kernel(ProbeParams const q, RawSphere const* __restrict__ sph, int n,
                                 RawCamera const* __restrict__ camp)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= q.count) {
        return;
    }
    RawCamera const cam = *camp;
    uint32_t const x = q.x[i], y = q.y[i], sx = q.sx[i], sy = q.sy[i];
    uint32_t const slot = ((y * q.width + x) * q.ns + sy) * q.ns + sx;
    Rng64 rng;
    rng.g = rng_open(q.key, slot, q.sample[i]);
    rng.draws = 0;
    Ray64 const pr = primary_ray(cam, x, y, sx, sy, q.width, q.height, q.ns, rng);
    if(q.ray != nullptr) {
        q.ray[6 * i + 0] = pr.o.x;
        q.ray[6 * i + 1] = pr.o.y;
        q.ray[6 * i + 2] = pr.o.z;
        q.ray[6 * i + 3] = pr.d.x;
        q.ray[6 * i + 4] = pr.d.y;
        q.ray[6 * i + 5] = pr.d.z;
    }
    double t = 0.0;
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

// The sequence of spheres a sample's path visits (depth by depth), for classifying the samples whose radiance differs
// from the oracle's: same arithmetic as probe_f64_kernel, nothing else recorded.
__global__ void trail_f64_kernel(ProbeParams const q, RawSphere const* __restrict__ sph, int n,
                                 RawCamera const* __restrict__ camp, int32_t* __restrict__ trail, int trail_len)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= q.count) {
        return;
    }
    RawCamera const cam = *camp;
    uint32_t const x = q.x[i], y = q.y[i], sx = q.sx[i], sy = q.sy[i];
    uint32_t const slot = ((y * q.width + x) * q.ns + sy) * q.ns + sx;
    Rng64 rng;
    rng.g = rng_open(q.key, slot, q.sample[i]);
    rng.draws = 0;
    Ray64 const pr = primary_ray(cam, x, y, sx, sy, q.width, q.height, q.ns, rng);
    Counters64 cnt{ 0, 0, 0, 0 };
    cnt.trail = trail + static_cast<size_t>(i) * static_cast<size_t>(trail_len);
    cnt.trail_len = trail_len;
    for(int d = 0; d < trail_len; ++d) {
        cnt.trail[d] = -2; // depth never reached
    }
    (void)radiance(sph, n, pr, rng, cnt);
}

// One thread per sub-pixel slot, samples in index order: the accumulation order is
// fixed, so the FP64 image is reproducible run to run.
__global__ void render_f64_kernel(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width, uint32_t height,
                                  uint32_t ns, RawSphere const* __restrict__ sph, int n,
                                  RawCamera const* __restrict__ camp, double* __restrict__ accum,
                                  DeviceCounters* counters)
{
    uint32_t const slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t const nslots = width * height * ns * ns;
    Counters64 cnt{ 0, 0, 0, 0 };
    if(slot < nslots) {
        RawCamera const cam = *camp;
        uint32_t const sx = slot % ns;
        uint32_t const q = slot / ns;
        uint32_t const sy = q % ns;
        uint32_t const pix = q / ns;
        uint32_t const y = pix / width;
        uint32_t const x = pix - y * width;
        // the running sum CONTINUES from what the slot holds (main.cpp:184-191 adds sample after sample): a render split
        // into several calls, or resumed from a checkpoint, performs the very same additions as one call
        double* a = accum + 4 * static_cast<size_t>(slot);
        V3 sum = mk(a[0], a[1], a[2]);
        for(uint32_t s = 0; s < samples; ++s) {
            Rng64 rng;
            rng.g = rng_open(key, slot, first_sample + s);
            rng.draws = 0;
            Ray64 const pr = primary_ray(cam, x, y, sx, sy, width, height, ns, rng);
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

cudaError_t launch_trail_f64(ProbeParams const& p, RawSphere const* spheres, int n, RawCamera const* cam, int32_t* trail,
                             int trail_len, cudaStream_t stream)
{
    if(p.count == 0) {
        return cudaSuccess;
    }
    trail_f64_kernel<<<(p.count + 127) / 128, 128, 0, stream>>>(p, spheres, n, cam, trail, trail_len);
    return cudaGetLastError();
}

cudaError_t launch_probe_f64(ProbeParams const& p, RawSphere const* spheres, int n, RawCamera const* cam,
                             cudaStream_t stream)
{
    if(p.count == 0) {
        return cudaSuccess;
    }
    unsigned const threads = 128;
    probe_f64_kernel<<<(p.count + threads - 1) / threads, threads, 0, stream>>>(p, spheres, n, cam);
    return cudaGetLastError();
}

cudaError_t launch_render_f64(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width, uint32_t height,
                              uint32_t ns, RawSphere const* spheres, int n, RawCamera const* cam, double* accum64,
                              DeviceCounters* counters, cudaStream_t stream)
{
    uint32_t const nslots = width * height * ns * ns;
    unsigned const threads = 128;
    render_f64_kernel<<<(nslots + threads - 1) / threads, threads, 0, stream>>>(key, first_sample, samples, width,
                                                                                  height, ns, spheres, n, cam, accum64,
                                                                                  counters);
    return cudaGetLastError();
}

} // namespace ptb
