// ptb_smallpt_f32.cuh -- FP32 throughput arithmetic of the reference's stand-alone smallpt fork,
// /root/reference/sandbox/main.cpp.  Shares the closest-hit scan and the
// diffuse scatter with the src/ integrator (ptb_path_f32.cuh); everything else follows the sandbox:
//   gen_smallpt     :253-261  tent-filtered pinhole camera, ray pushed 140 units along the unit direction
//   bounce_smallpt  :149-227  black miss, roulette after depth 5, mirror without the src/ "fuzz" draw,
//                             glass IOR 1.5 with reflect+refract SPLITTING while depth <= 2
// The recursion becomes a loop with a two-entry per-thread stack in shared memory (a split can only
// happen at depth 1 and 2); the refraction branch is traced first, the reflection waits on the stack:
// that is the order of the pinned reference build, so the random stream is consumed identically.
#pragma once

#include "ptb_path_f32.cuh"

namespace ptb {

constexpr int kSmallptRoulette = 5;       // sandbox/main.cpp:167
constexpr int kSmallptSplitDepth = 2;     // sandbox/main.cpp:223
constexpr int kSmallptSafetyDepth = 1 << 20; // the sandbox has no depth limit; guards against a hang only
constexpr int kSplitFields = 10;          // o(3) d(3) w(3) depth

__device__ __forceinline__ float tent(float r1)
{
    return r1 < 1.0f ? fast_sqrt(r1) - 1.0f : 1.0f - fast_sqrt(2.0f - r1);
}

__device__ __forceinline__ void gen_smallpt(PathF32& p, SmallptCamF32 const& cam, uint32_t x, uint32_t y, uint32_t sx,
                                            uint32_t sy)
{
    float const fx = tent(2.0f * rng_uniform_f32(p.rng));
    float const fy = tent(2.0f * rng_uniform_f32(p.rng));
    float const a = fmaf((static_cast<float>(sx) + 0.5f + fx), 0.5f, static_cast<float>(x)) * cam.inv_w - 0.5f;
    float const b = fmaf((static_cast<float>(sy) + 0.5f + fy), 0.5f, static_cast<float>(y)) * cam.inv_h - 0.5f;
    float const dx = fmaf(cam.cyx, b, fmaf(cam.cxx, a, cam.dx));
    float const dy = fmaf(cam.cyy, b, cam.dy);
    float const dz = fmaf(cam.cyz, b, cam.dz);
    float const inv = fast_rsqrt(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    p.dx = dx * inv;
    p.dy = dy * inv;
    p.dz = dz * inv;
    p.ox = fmaf(p.dx, cam.push, cam.ox);
    p.oy = fmaf(p.dy, cam.push, cam.oy);
    p.oz = fmaf(p.dz, cam.push, cam.oz);
    p.len = 1.0f;
    p.tr = p.tg = p.tb = 1.0f;
    p.er = p.eg = p.eb = 0.0f;
    p.depth = 0;
    p.last = -1;
}

// per-thread split stack in shared memory: field-major so that a warp touches 32 consecutive words
struct SplitStack
{
    float* base;   // [2][kSplitFields][threads]
    int threads;
    int count;

    __device__ __forceinline__ float& at(int level, int field) const
    {
        return base[(level * kSplitFields + field) * threads + static_cast<int>(threadIdx.x)];
    }
};

// One bounce of the sandbox's radiance().  Returns true while there is something left to trace for this
// camera sample (the current segment or a branch waiting on the stack); p.er/eg/eb accumulate W * e.
template<bool kCount>
__device__ __forceinline__ bool bounce_smallpt(PathF32& p, bool hit, float t, int id, ShadePlanes const& sp,
                                               BounceCounters& cnt, SplitStack& st)
{
    bool segment_done = !hit; // black on a miss (:154-156)
    if(hit) {
        float4 const sa = sp.a[id];
        float4 const sb = sp.b[id];
        float const hx = fmaf(p.dx, t, p.ox);
        float const hy = fmaf(p.dy, t, p.oy);
        float const hz = fmaf(p.dz, t, p.oz);
        float nx, ny, nz; // n = norm(x - p), outward (sandbox/main.cpp:160)
        unit_normal(hx, hy, hz, sa, nx, ny, nz);
        // every visited hit contributes W * e, whether the roulette then kills the path (it returns
        // obj.e) or not (obj.e + f * child)
        p.er = fmaf(p.tr, sb.x, p.er);
        p.eg = fmaf(p.tg, sb.y, p.eg);
        p.eb = fmaf(p.tb, sb.z, p.eb);
        p.depth++;
        float4 col;
        if(p.depth > kSmallptRoulette) {
            float4 const sc = sp.c[id];
            if(!(rng_uniform_f32(p.rng) < sc.w) || p.depth > kSmallptSafetyDepth) {
                segment_done = true;
            }
            col = sp.d[id];
        }
        else {
            col = sp.c[id];
        }
        if(!segment_done) {
            p.ox = hx;
            p.oy = hy;
            p.oz = hz;
            p.last = id;
            int const refl = __float_as_int(sb.w) & 0xff;
            float const dn = fmaf(nx, p.dx, fmaf(ny, p.dy, nz * p.dz));
            if(refl == 1) { // SPEC :196-198
                if(kCount) {
                    cnt.specular++;
                }
                p.tr *= col.x;
                p.tg *= col.y;
                p.tb *= col.z;
                float const k = -2.0f * dn;
                p.dx = fmaf(k, nx, p.dx);
                p.dy = fmaf(k, ny, p.dy);
                p.dz = fmaf(k, nz, p.dz);
            }
            else if(refl == 0) { // DIFF :176-186
                if(kCount) {
                    cnt.diffuse++;
                }
                p.tr *= col.x;
                p.tg *= col.y;
                p.tb *= col.z;
                bool const front = dn < 0.0f;
                scatter_diffuse(p, front ? nx : -nx, front ? ny : -ny, front ? nz : -nz);
            }
            else { // REFR :200-226
                if(kCount) {
                    cnt.dielectric++;
                }
                float const wr = p.tr * col.x, wg = p.tg * col.y, wb = p.tb * col.z;
                bool const into = dn < 0.0f; // n.nl > 0  <=>  the ray faces the outward normal
                float const nnt = into ? (1.0f / 1.5f) : 1.5f;
                float const ddn = -fabsf(dn); // r.d . nl
                float const cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
                float const k = -2.0f * dn; // reflection: d - 2 (n.d) n
                float const rx = fmaf(k, nx, p.dx), ry = fmaf(k, ny, p.dy), rz = fmaf(k, nz, p.dz);
                if(cos2t < 0.0f) { // total internal reflection
                    p.tr = wr;
                    p.tg = wg;
                    p.tb = wb;
                    p.dx = rx;
                    p.dy = ry;
                    p.dz = rz;
                }
                else {
                    float const s = (into ? 1.0f : -1.0f) * fmaf(ddn, nnt, fast_sqrt(cos2t));
                    float tx = fmaf(p.dx, nnt, -nx * s), ty = fmaf(p.dy, nnt, -ny * s), tz = fmaf(p.dz, nnt, -nz * s);
                    float const inv = fast_rsqrt(fmaf(tx, tx, fmaf(ty, ty, tz * tz)));
                    tx *= inv;
                    ty *= inv;
                    tz *= inv;
                    float const c = 1.0f - (into ? -ddn : fmaf(tx, nx, fmaf(ty, ny, tz * nz)));
                    float const c2 = c * c;
                    float const re = fmaf(0.96f, c2 * c2 * c, 0.04f); // R0 = (0.5/2.5)^2
                    float const tr = 1.0f - re;
                    if(p.depth > kSmallptSplitDepth) {
                        float const pp = fmaf(0.5f, re, 0.25f);
                        if(rng_uniform_f32(p.rng) < pp) {
                            float const rp = re * fast_rcp(pp);
                            p.tr = wr * rp;
                            p.tg = wg * rp;
                            p.tb = wb * rp;
                            p.dx = rx;
                            p.dy = ry;
                            p.dz = rz;
                        }
                        else {
                            float const tp = tr * fast_rcp(1.0f - pp);
                            p.tr = wr * tp;
                            p.tg = wg * tp;
                            p.tb = wb * tp;
                            p.dx = tx;
                            p.dy = ty;
                            p.dz = tz;
                        }
                    }
                    else {
                        // split: park the reflection, go on with the refraction
                        int const l = st.count++;
                        st.at(l, 0) = hx;
                        st.at(l, 1) = hy;
                        st.at(l, 2) = hz;
                        st.at(l, 3) = rx;
                        st.at(l, 4) = ry;
                        st.at(l, 5) = rz;
                        st.at(l, 6) = wr * re;
                        st.at(l, 7) = wg * re;
                        st.at(l, 8) = wb * re;
                        st.at(l, 9) = __int_as_float(p.depth | ((id + 1) << 8)); // depth <= 2 here
                        p.tr = wr * tr;
                        p.tg = wg * tr;
                        p.tb = wb * tr;
                        p.dx = tx;
                        p.dy = ty;
                        p.dz = tz;
                    }
                }
            }
        }
    }
    if(segment_done) {
        if(st.count == 0) {
            return false;
        }
        int const l = --st.count;
        p.ox = st.at(l, 0);
        p.oy = st.at(l, 1);
        p.oz = st.at(l, 2);
        p.dx = st.at(l, 3);
        p.dy = st.at(l, 4);
        p.dz = st.at(l, 5);
        p.tr = st.at(l, 6);
        p.tg = st.at(l, 7);
        p.tb = st.at(l, 8);
        int const packed = __float_as_int(st.at(l, 9));
        p.depth = packed & 0xFF;
        p.last = (packed >> 8) - 1;
    }
    return true;
}

} // namespace ptb
