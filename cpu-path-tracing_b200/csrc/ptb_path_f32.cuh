// ptb_path_f32.cuh -- the FP32 throughput arithmetic of one path: camera sample,
// closest hit, shading, scattering.  Shared by the megakernel, the wavefront
// kernels and the FP32 probe so that all three trace bit-identical paths.
//
// What each piece replaces in the reference (/root/reference/src):
//   gen_primary   main.cpp:186-190 jitter + camera.cpp:19-38 thin lens (incl. the
//                 offset = rd*s + rd*t quirk and the un-normalised direction)
//   closest_hit   main.cpp:30-42 linear scan + sphere.cpp:6-30 quadratic
//   shade_bounce  main.cpp:115-154 sky / emission / Russian roulette / throughput,
//                 hit_record.cpp:3-12, diffuse_ray :44-58, specular_ray :60-67,
//                 dielectric_ray :69-97
// The draw ORDER of the random stream is the reference's, so a
// sample consumes the same uniforms here, in the FP64 parity kernels and in the
// oracle; only the rounding differs.
#pragma once

#include "ptb_rng.cuh"
#include "ptb_scene.cuh"

namespace ptb {

#ifdef PTB_EPS
constexpr float kEpsilon = PTB_EPS;
#else
constexpr float kEpsilon = 1e-4f;      // constants.hpp:7
#endif
constexpr int kDepthLimit = 100;       // constants.hpp:10
constexpr int kRouletteThreshold = 4;  // main.cpp:106
constexpr uint32_t kNoHitBits = 0x7F800000u; // +inf: also what "t < inf" (main.cpp:41) becomes
#ifdef PTB_ID_BITS
constexpr int kIdBits = PTB_ID_BITS;
#else
constexpr int kIdBits = 4;                    // specialised kernels: up to 16 spheres
#endif

__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// sqrt of the sphere discriminant: MUFU.SQRT; x < 0 -> NaN, which the key ordering treats
// as a miss.
__device__ __forceinline__ float disc_sqrt(float x)
{
    return fast_sqrt(x);
}
// a + sqrt(x) and a - sqrt(x) for a discriminant x.  PTB_DISC_RSQRT (an experiment, off): sqrt(x) = x * rsqrt(x) folded
// into the addition -- MUFU.RSQ + FFMA instead of MUFU.SQRT + FADD, same instruction count, half the XU-pipe time; x = 0
// gives NaN = miss where sqrt gives a grazing hit.  Measured 0.4 % (box_mirror) and 0.75 % (box) SLOWER than MUFU.SQRT:
// the XU pipe is not what limits the scan ("tried and dropped").
__device__ __forceinline__ float add_root(float a, float x)
{
#ifdef PTB_DISC_RSQRT
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(x, y, a);
#else
    return a + disc_sqrt(x);
#endif
}
__device__ __forceinline__ void sub_add_root(float a, float x, float& lo, float& hi)
{
#ifdef PTB_DISC_RSQRT
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    lo = fmaf(-x, y, a);
    hi = fmaf(x, y, a);
#else
    float const sq = disc_sqrt(x);
    lo = a - sq;
    hi = a + sq;
#endif
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- small device utilities shared by the kernels ------------------------------------------------------------------
// One 16-byte vector reduction into a slot of the accumulation buffer (sm_90+): no read-modify-write.
__device__ __forceinline__ void red_add_v4(float4* addr, float x, float y, float z, float w)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :
                 : "l"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
                 : "memory");
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
    return __reduce_add_sync(0xffffffffu, v);
}

struct PathF32
{
    float ox, oy, oz; // ray origin (shifted frame)
    float dx, dy, dz; // ray direction, UNIT length (see `len`)
    float len;        // length the reference's un-normalised direction would have (camera.cpp:36):
                      // the reference compares roots in units of t_ref = t / len against epsilon
    float tr, tg, tb; // accumulated_reflectance (main.cpp:108)
    float er, eg, eb; // accumulated_emission   (main.cpp:107)
    Rng rng;
    int depth;
    int last; // list position of the sphere this ray starts on, -1 for a camera ray (robust shapes only)
};

// slot -> (x, y, sx, sy) in reference loop coordinates; slot = ((y*W+x)*ns+sy)*ns+sx
__device__ __forceinline__ void slot_coords(uint32_t slot, uint32_t width, uint32_t ns, uint32_t& x, uint32_t& y,
                                            uint32_t& sx, uint32_t& sy)
{
    uint32_t pix;
    if(ns == 2u) {
        sx = slot & 1u;
        sy = (slot >> 1) & 1u;
        pix = slot >> 2;
    }
    else {
        sx = slot % ns;
        uint32_t const q = slot / ns;
        sy = q % ns;
        pix = q / ns;
    }
    y = pix / width;
    x = pix - y * width;
}

// main.cpp:186-190 + camera.cpp:19-38.  The reference keeps the camera ray un-normalised
// (|d| ~ focus distance) and a mirror bounce preserves that length, so its epsilon test
// (sphere.cpp:21, in units of t) scales with |d|.  Here the direction is normalised ONCE,
// when the sample is generated, and the original length travels with the path as `len`:
// t_ref < epsilon  <=>  t_unit < epsilon * len.  That removes a = d.d, 1/a and every
// multiplication by a from the sphere tests.
__device__ __forceinline__ void gen_primary(PathF32& p, CameraF32 const& cam, uint32_t x, uint32_t y, uint32_t sx,
                                            uint32_t sy)
{
    float const len = cam.sub_len;
    float const u0 = rng_uniform_f32(p.rng);
    float const u1 = rng_uniform_f32(p.rng);
    // (explicit FMAs: where a * b + c * d leaves the choice of what to fuse to the compiler, two builds of this
    //  header -- nvcc and the run-time compiler, or two kernels that inline it -- may round differently)
    float const xs = fmaf(len, u0, fmaf(static_cast<float>(sx), len, static_cast<float>(x)));
    float const ys = fmaf(len, u1, fmaf(static_cast<float>(sy), len, static_cast<float>(y)));
    float const s = xs * cam.inv_w;
    float const t = ys * cam.inv_h;

    // rejection-sampled unit disk, x drawn before y (camera.cpp:21-29)
    float lx, ly;
    do {
        lx = fmaf(2.0f, rng_uniform_f32(p.rng), -1.0f);
        ly = fmaf(2.0f, rng_uniform_f32(p.rng), -1.0f);
    } while(fmaf(lx, lx, ly * ly) >= 1.0f);

    float const st = s + t; // offset = rd*s + rd*t, z component 0 (camera.cpp:34-35)
    float const offx = lx * cam.lens_radius * st;
    float const offy = ly * cam.lens_radius * st;

    p.ox = cam.px + offx;
    p.oy = cam.py + offy;
    p.oz = cam.pz;
    float const dx = fmaf(cam.bx, t, fmaf(cam.ax, s, cam.rx)) - offx;
    float const dy = fmaf(cam.by, t, fmaf(cam.ay, s, cam.ry)) - offy;
    float const dz = fmaf(cam.bz, t, fmaf(cam.az, s, cam.rz));
    float const d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    float const inv = fast_rsqrt(d2);
    p.dx = dx * inv;
    p.dy = dy * inv;
    p.dz = dz * inv;
    p.len = d2 * inv;
    p.tr = p.tg = p.tb = 1.0f;
    p.er = p.eg = p.eb = 0.0f;
    p.depth = 0;
    p.last = -1;
}

// Per-ray invariants of the sphere tests.
struct RayTerms
{
    float eps;   // epsilon * len : roots are compared in units of t along the UNIT direction
    float od;    // o.d
    float oo;    // o.o
    float o2x, o2y, o2z; // 2*o
    float kod, koo;      // k*od, k*oo for scenes whose big spheres share one radius (k = 1/2R)
};

__device__ __forceinline__ RayTerms ray_terms(PathF32 const& p, float k_uniform = 0.0f)
{
    RayTerms r;
    // __fmul_rn: these products are ADDED to other terms in every sphere test; left as plain `*` the compiler may fuse
    // them into those additions in one build and not in another (constant-bank vs immediate operands), which changes a
    // last bit once in ~3e6 paths
    r.eps = __fmul_rn(kEpsilon, p.len);
    r.od = fmaf(p.ox, p.dx, fmaf(p.oy, p.dy, p.oz * p.dz));
    r.oo = fmaf(p.ox, p.ox, fmaf(p.oy, p.oy, p.oz * p.oz));
    r.o2x = p.ox + p.ox;
    r.o2y = p.oy + p.oy;
    r.o2z = p.oz + p.oz;
    r.kod = __fmul_rn(k_uniform, r.od);
    r.koo = __fmul_rn(k_uniform, r.oo);
    return r;
}

// Candidate key of one sphere: bits of (t - eps) for the smallest root with t >= eps, or
// something >= 0x7F800000 when there is none.  Non-negative floats order like unsigned
// integers; negative values (root < eps) and NaN (discriminant < 0, sphere.cpp:14) order
// ABOVE +inf, so a single unsigned min does "root < eps -> try the far root -> reject"
// (sphere.cpp:21-27).  kBoth = false keeps the near root only (opaque sphere seen from
// outside: the far root can never be the answer).
// kRobust (scenes that are large against epsilon, and every run-time-count scene): a ray that starts ON
// sphere `self` knows one root of that sphere is 0 -- the point it stands on -- so the other one is
// 2*nb exactly; taking it from there instead of from nb +- sqrt(nb^2 - c) with c ~ 0 removes the
// binary32 failure where a ray refracted OUT of a glass ball re-hits it from inside a few 1e-4 further
// on, is totally reflected and then circles inside for ~1000 bounces (seen on the sandbox scene).
// Discriminant of a small sphere for scenes that are LARGE against the sphere: half_b^2 - c with c = |c - o|^2 - r^2
// subtracts two numbers of size |c - o|^2 to get one of size r^2 -- at 70 units from a r = 0.2 sphere binary32 leaves
// an error of 1.5 % of the sphere's cross-section in the hit / miss decision.  r^2 - |(c - o) - ((c - o).d) d|^2, the
// squared distance from the centre to the ray, has no such cancellation (unit d): two more FFMA.
__device__ __forceinline__ float small_disc_far(float cx, float cy, float cz, float nb, float r2, PathF32 const& p)
{
    float const qx = fmaf(-nb, p.dx, cx), qy = fmaf(-nb, p.dy, cy), qz = fmaf(-nb, p.dz, cz);
    return fmaf(-qx, qx, fmaf(-qy, qy, fmaf(-qz, qz, r2)));
}

template<bool kBoth, bool kRobust = false>
__device__ __forceinline__ uint32_t key_small(SmallGeo const& s, PathF32 const& p, RayTerms const& r, bool self = false)
{
#ifdef PTB_JIT_SCENE_INIT
    // run-time build: a coordinate that IS zero (a literal) needs no subtraction -- "0 - o" does not fold by itself
    // (signed zero), written as a negation it disappears into the operand modifiers of the FFMAs that follow
    float const cx = s.cx == 0.0f ? -p.ox : s.cx - p.ox, cy = s.cy == 0.0f ? -p.oy : s.cy - p.oy, cz = s.cz == 0.0f ? -p.oz : s.cz - p.oz;
#else
    float const cx = s.cx - p.ox, cy = s.cy - p.oy, cz = s.cz - p.oz; // c - o = -oc
#endif
    float const nb = fmaf(cx, p.dx, fmaf(cy, p.dy, cz * p.dz));       // -half_b
    float disc;
    if constexpr(kRobust) {
        disc = small_disc_far(cx, cy, cz, nb, s.r2, p);
    }
    else {
        float const cc = fmaf(cx, cx, fmaf(cy, cy, fmaf(cz, cz, -s.r2))); // c (sphere.cpp:11)
        disc = fmaf(nb, nb, -cc);
    }
    float const h = nb - r.eps;
    float tn, tf;
    sub_add_root(h, disc, tn, tf);
    uint32_t key;
    if constexpr(kBoth) {
        key = min(__float_as_uint(tn), __float_as_uint(tf));
    }
    else {
        key = __float_as_uint(tn);
    }
    if constexpr(kRobust) {
        key = self ? __float_as_uint(nb + h) : key; // 2*nb - eps
    }
    return key;
}

// x with its sign flipped when hb >= 0 (sign bit clear): x ^ (~hb & 0x80000000) as ONE three-input logic op.  Left to the
// compiler the expression becomes a negation (FADD) plus a LOP3 per sphere test.
__device__ __forceinline__ float flip_unless_negative(float x, float hb)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD2;" : "=r"(r) : "r"(__float_as_uint(x)), "r"(__float_as_uint(hb)));
    return __uint_as_float(r);
}

// Big-sphere form.  With m = sqrt(disc) + |hb| (no cancellation, m > 0) the two roots are
//   sigma * c / m   and   sigma * m / k,     sigma = -sign(hb)  (+1 when approaching),
// which is the numerically stable pairing for EITHER sign of hb: the earlier c/(sqrt - hb)
// is only safe for hb <= 0; moving away from a sphere (hb > 0) it divides by a difference
// that rounding can turn into a tiny POSITIVE number, i.e. a spurious hit ~1e7 away.
template<bool kBoth, bool kRobust = false>
__device__ __forceinline__ uint32_t key_big(BigGeo const& b, PathF32 const& p, RayTerms const& r, bool self = false)
{
    float const hb = fmaf(p.dx, b.gx, fmaf(p.dy, b.gy, fmaf(p.dz, b.gz, b.k * r.od)));           // half_b / 2R
    float const cp = fmaf(r.o2x, b.gx, fmaf(r.o2y, b.gy, fmaf(r.o2z, b.gz, fmaf(b.k, r.oo, b.K)))); // c / 2R
    float const disc = fmaf(hb, hb, -(b.k * cp));
    float const m = add_root(fabsf(hb), disc);
    float const cs = flip_unless_negative(cp, hb); // sigma = -1 when hb >= 0
    float const t1 = fmaf(cs, fast_rcp(m), -r.eps);
    uint32_t key;
    if constexpr(kBoth) {
        float const ms = flip_unless_negative(m, hb);
        float const t2 = fmaf(ms, b.two_r, -r.eps);
        key = min(__float_as_uint(t1), __float_as_uint(t2));
    }
    else {
        key = __float_as_uint(t1);
    }
    if constexpr(kRobust) {
        // standing on the sphere: roots are 0 and -2 hb / k
        key = self ? __float_as_uint(fmaf(-2.0f * hb, b.two_r, -r.eps)) : key;
    }
    return key;
}

// Near-only big sphere whose centre lies ON a coordinate axis of the shifted frame (the host
// picks the frame so that the R = 1e6 walls of the box scenes, whose centres all sit on the
// three lines through (0,0,-1), qualify): g = -k c has a single non-zero component, so the two
// 3-term dot products shrink to one FFMA each.  kUniformK: all big spheres share k = 1/2R, so
// k*od and k*oo are per-ray values.  Exact -- nothing is approximated.
template<int AXIS, bool kUniformK>
__device__ __forceinline__ uint32_t key_big_axis(float ga, float K, float k, PathF32 const& p, RayTerms const& r)
{
    float const da = AXIS == 0 ? p.dx : (AXIS == 1 ? p.dy : p.dz);
    float const oa2 = AXIS == 0 ? r.o2x : (AXIS == 1 ? r.o2y : r.o2z);
    float const hb = fmaf(da, ga, kUniformK ? r.kod : k * r.od);
#ifdef PTB_JIT_SCENE_INIT
    // run-time build: (2 o_a) * g = o_a * (2 g) bit for bit, and 2 g is a literal -- the doubled origin is never formed
    float const oa = AXIS == 0 ? p.ox : (AXIS == 1 ? p.oy : p.oz);
    float const cp = fmaf(oa, 2.0f * ga, kUniformK ? (K == 0.0f ? r.koo : r.koo + K) : fmaf(k, r.oo, K));
    (void)oa2;
#else
    float const cp = fmaf(oa2, ga, kUniformK ? r.koo + K : fmaf(k, r.oo, K));
#endif
    float const disc = fmaf(hb, hb, -(k * cp));
    float const m = add_root(fabsf(hb), disc);
    return __float_as_uint(fmaf(flip_unless_negative(cp, hb), fast_rcp(m), -r.eps));
}

// Two near-only axis spheres that are mirror images of each other (the left / right and top / bottom walls of
// the box scenes: centres +-C on a frame axis, one radius).  A ray outside a sphere and moving AWAY from its
// centre never meets it, and with |C| ~ 1e6 "towards the +C centre" is the sign of the direction's axis component
// for every ray that could reach the wall before t ~ 1e5.  So the pair costs one test: with s = sign(d_a),
// half_b = k o.d - |d_a| G and c' = k o.o + K - s (2 o_a) G, G = k C, are the selected sphere's coefficients exactly
// (same bits as key_big_axis gives for it).  (ga, K) are the +C sphere's; returns its key with `sel` = 0 for the +C
// sphere, 1 for the -C one.
template<int AXIS>
__device__ __forceinline__ uint32_t key_big_pair(float ga, float K, float k, PathF32 const& p, RayTerms const& r, uint32_t& sel)
{
    float const da = AXIS == 0 ? p.dx : (AXIS == 1 ? p.dy : p.dz);
    float const oa2 = AXIS == 0 ? r.o2x : (AXIS == 1 ? r.o2y : r.o2z);
    float const G = fabsf(ga);
    uint32_t const sign = __float_as_uint(da) & 0x80000000u;
    float const hb = fmaf(-fabsf(da), G, r.kod);
#ifdef PTB_JIT_SCENE_INIT
    float const oa = AXIS == 0 ? p.ox : (AXIS == 1 ? p.oy : p.oz); // (2 o_a) * G = o_a * (2 G), 2 G a literal
    float const cp = fmaf(-__uint_as_float(__float_as_uint(oa) ^ sign), 2.0f * G, r.koo + K);
    (void)oa2;
#else
    float const cp = fmaf(-__uint_as_float(__float_as_uint(oa2) ^ sign), G, r.koo + K);
#endif
    float const disc = fmaf(hb, hb, -(k * cp));
    float const m = add_root(fabsf(hb), disc);
    sel = sign >> 31;
    return __float_as_uint(fmaf(flip_unless_negative(cp, hb), fast_rcp(m), -r.eps));
}

// ---- closest hit through the bounding-volume hierarchy (ptb_bvh.hpp) -------------
// Same answer as the linear scan of main.cpp:30-42 over the small spheres: every candidate goes through the very
// same root computation (key_small, robust form), the smallest key wins, equal keys go to the lower list position.
// The tree only skips spheres whose padded box the ray misses or enters beyond the best root so far.

__device__ __forceinline__ void bvh_closest_hit(GeoLists const& gl, PathF32 const& p, RayTerms const& r, uint32_t& best,
                                                int& id)
{
    // slab test: t_mid = mid * (1/d) - o * (1/d), entry / exit = t_mid -+ half * |1/d| -- three FFMAs per slab and NO min / max
    // to order the two planes (FMNMX issues at half rate, on the pipe this loop loads most: -3.7 % time on config 5 against
    // one FFMA per plane + min + max).  Against (plane - o) * (1/d) the roundings move t by up to ~1e-7 (|o| + |mid|) in
    // space units -- the boxes are padded by 1e-6 |c| + 1e-5 for exactly that (ptb_bvh.hpp).
    // A zero direction component must not become inf (inf - inf): |d_a| is floored at 1e-20, which keeps both plane
    // distances finite and of the right signs (+-1e20 scale: "never reached" or "always inside").
    auto const safe_rcp = [](float d) { return fast_rcp(fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); };
    float const ix = safe_rcp(p.dx), iy = safe_rcp(p.dy), iz = safe_rcp(p.dz);
    float const nx = -p.ox * ix, ny = -p.oy * iy, nz = -p.oz * iz;
    float const jx = fabsf(ix), jy = fabsf(iy), jz = fabsf(iz);
    float tbest = best < kNoHitBits ? __uint_as_float(best) + r.eps : 3.0e38f;
    int stack[kBvhStack];
    int sp = 0;
    int node = gl.bvh_root;
    for(;;) {
        if(node >= 0) {
            float4 const n0 = gl.bvh_nodes[4 * node + 0];
            float4 const n1 = gl.bvh_nodes[4 * node + 1];
            float4 const n2 = gl.bvh_nodes[4 * node + 2];
            float4 const n3 = gl.bvh_nodes[4 * node + 3];
            float const acx = fmaf(n0.x, ix, nx), acy = fmaf(n0.z, iy, ny), acz = fmaf(n2.x, iz, nz);
            float const bcx = fmaf(n1.x, ix, nx), bcy = fmaf(n1.z, iy, ny), bcz = fmaf(n2.z, iz, nz);
            float const amin = fmaxf(fmaxf(fmaf(-n0.y, jx, acx), fmaf(-n0.w, jy, acy)), fmaxf(fmaf(-n2.y, jz, acz), 0.0f));
            float const amax = fminf(fminf(fmaf(n0.y, jx, acx), fmaf(n0.w, jy, acy)), fminf(fmaf(n2.y, jz, acz), tbest));
            float const bmin = fmaxf(fmaxf(fmaf(-n1.y, jx, bcx), fmaf(-n1.w, jy, bcy)), fmaxf(fmaf(-n2.w, jz, bcz), 0.0f));
            float const bmax = fminf(fminf(fmaf(n1.y, jx, bcx), fmaf(n1.w, jy, bcy)), fminf(fmaf(n2.w, jz, bcz), tbest));
            bool const ha = amin <= amax, hb = bmin <= bmax;
            int const ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
            if(ha && hb) {
                bool const a_first = amin <= bmin;
                stack[sp++] = a_first ? cb : ca;
                node = a_first ? ca : cb;
                continue;
            }
            if(ha || hb) {
                node = ha ? ca : cb;
                continue;
            }
        }
        else {
            int const code = ~node;
            int const first = code >> 3, count = (code & 7) + 1;
            // leaves hold ONE sphere (ptb_bvh.hpp; up to 8 only where coincident centres cannot be split): no unrolling,
            // and the 16-byte record comes in one load
#pragma unroll 1
            for(int k = 0; k < count; ++k) {
                float4 const s = __ldg(reinterpret_cast<float4 const*>(gl.bvh_geo) + (first + k)); // cx, cy, cz, r^2
                int const pw = __ldg(gl.bvh_pos + (first + k));
                int const pos = pw & 0x7fffffff;
                float const cx = s.x - p.ox, cy = s.y - p.oy, cz = s.z - p.oz;
                float const nb = fmaf(cx, p.dx, fmaf(cy, p.dy, cz * p.dz));
                float const h = nb - r.eps;
                float tn, tf;
                sub_add_root(h, small_disc_far(cx, cy, cz, nb, s.w, p), tn, tf);
                uint32_t const kn = __float_as_uint(tn);
                uint32_t key = pw < 0 ? min(kn, __float_as_uint(tf)) : kn; // key_small<kBoth>
                key = p.last == pos ? __float_as_uint(nb + h) : key;           // key_small<., kRobust>: standing on it
                if(key < best || (key == best && pos < id)) {
                    best = key;
                    id = pos;
                    tbest = key < kNoHitBits ? __uint_as_float(key) + r.eps : tbest;
                }
            }
        }
        if(sp == 0) {
            break;
        }
        node = stack[--sp];
    }
}

// Compile-time description of a scene's geometry lists; SN < 0 = run-time counts.
//   SN small near-only, SB small both-roots, BN big near-only (of which the first BX / BY / BZ
//   are x- / y- / z-axis spheres), BB big both-roots, UK = big spheres share one radius.
//   EM = the list position may ride in the low mantissa bits of the hit key (see closest_hit);
//   only when the scene is small against epsilon: the 2^-19 relative truncation of t moves the
//   hit point by up to 2e-6 * t, which must stay far below epsilon = 1e-4 (a refracted ray
//   leaving a glass sphere re-hits it from inside otherwise -- seen on the sandbox scene, whose
//   extent is ~300 units).
//   PM = bit a set: axis group a is ONE mirror-image pair (centres +C and -C on the axis, same radius, the +C
//   sphere listed first): a ray moving towards +a can only meet the +C sphere and vice versa, so one test with
//   the signs folded in serves both (key_big_pair).
template<int SN, int SB, int BN, int BB, int BX = 0, int BY = 0, int BZ = 0, bool UK = false, bool EM = true, int PM = 0>
struct SceneShape
{
    static constexpr int pair_mask = PM;
    static_assert(SN < 0 || (((PM & 1) == 0 || BX == 2) && ((PM & 2) == 0 || BY == 2) && ((PM & 4) == 0 || BZ == 2)),
                  "a paired axis group holds exactly two spheres");
    static_assert(PM == 0 || (UK && EM), "pairs need the shared radius and index-in-key");
    static constexpr bool embed = EM;
    static constexpr int small_near = SN, small_both = SB, big_near = BN, big_both = BB;
    static constexpr int big_x = BX, big_y = BY, big_z = BZ;
    static constexpr bool uniform_k = UK;
    static constexpr bool generic = SN < 0;
    static constexpr int total = SN + SB + BN + BB;
    static_assert(SN < 0 || BX + BY + BZ <= BN, "axis spheres are a prefix of the near-only big list");
};
using GenericShape = SceneShape<-1, -1, -1, -1>;

// main.cpp:30-42.  Returns the LIST POSITION of the closest sphere (see ptb_scene.cuh).
//  - specialised shapes: counts known at compile time, geometry read from constant-bank
//    operands, fully unrolled; the list position rides in the low kIdBits mantissa bits of
//    the key, so the running minimum is one LOP3 (immediate) + one VIMNMX per sphere
//    (compare/select ops are half-rate on B200).  Cost: t is truncated by < 2^-19
//    relative, towards the ray origin.  Equal keys fall to the lower list position.
//  - generic shape: run-time counts, geometry from global memory (any scene size).
// kKeepReg: the mantissa mask comes in a REGISTER (keep_reg, loaded once per kernel through an opaque move) so
// that "(key & mask) | position" is a single three-input LOP3 with the position as its immediate; with both as
// immediates the compiler needs two LOP3 per sphere.
// What a run-time compiled kernel (ptb_jit.cpp) reads its sphere coefficients from: the same members closest_hit
// takes from the constant bank, initialised from literals so that they fold into instruction immediates.
template<class Shape>
struct JitSceneT
{
    SmallGeo small_geo[Shape::small_near + Shape::small_both > 0 ? Shape::small_near + Shape::small_both : 1];
    BigGeo big_geo[Shape::big_near + Shape::big_both > 0 ? Shape::big_near + Shape::big_both : 1];
    float axis_coef[Shape::big_near + Shape::big_both > 0 ? 2 * (Shape::big_near + Shape::big_both) : 2];
    int n_small_near, n_small, n_big_near, n_big; // never read: specialised shapes carry their counts as types
};

template<class Shape, bool kKeepReg = false, class Scene = ConstSceneF32>
__device__ __forceinline__ bool closest_hit(Scene const& cs, GeoLists const& gl, PathF32 const& p,
                                            RayTerms const& r, float& t_out, int& id_out, uint32_t keep_reg = 0u)
{
    // index-in-key scan: the running minimum starts ABOVE every key (a miss is any key >= +inf's bits), so that no
    // "min(best, inf)" is needed before the position is taken out of the winner
    uint32_t best = (!Shape::generic && Shape::embed) ? 0xFFFFFFFFu : kNoHitBits;
    int id = -1;
    if constexpr(!Shape::generic) {
        // only the index-in-key kernels are limited by the key's spare bits; full-precision keys carry the id separately
        static_assert(!Shape::embed || Shape::total <= (1 << kIdBits), "list position does not fit the key");
        constexpr uint32_t keep = Shape::embed ? ~((1u << kIdBits) - 1u) : 0xFFFFFFFFu;
        constexpr int NS = Shape::small_near + Shape::small_both;
        constexpr int NB = Shape::big_near + Shape::big_both;
        auto const take = [&](uint32_t k, int pos) {
            if constexpr(Shape::embed) {
                best = min(best, (k & (kKeepReg ? keep_reg : keep)) | static_cast<uint32_t>(pos));
            }
            else if(k < best) { // full-precision key: compare and select (two more half-rate ops per sphere)
                best = k;
                id = pos;
            }
        };
        constexpr bool kRobust = !Shape::embed;
#pragma unroll
        for(int i = 0; i < NS; ++i) {
            uint32_t const k = i < Shape::small_near ? key_small<false, kRobust>(cs.small_geo[i], p, r, p.last == i)
                                                     : key_small<true, kRobust>(cs.small_geo[i], p, r, p.last == i);
            take(k, i);
        }
#pragma unroll
        for(int i = 0; i < NB; ++i) {
            uint32_t k;
            constexpr int x0 = 0, y0 = Shape::big_x, z0 = Shape::big_x + Shape::big_y;
            if(((Shape::pair_mask & 1) != 0 && i == x0) || ((Shape::pair_mask & 2) != 0 && i == y0) ||
               ((Shape::pair_mask & 4) != 0 && i == z0)) {
                uint32_t sel;
                float const ga = cs.axis_coef[2 * i], K = cs.axis_coef[2 * i + 1], kk = cs.big_geo[0].k; // pairs: one radius
                k = i == x0 && (Shape::pair_mask & 1) != 0   ? key_big_pair<0>(ga, K, kk, p, r, sel)
                    : i == y0 && (Shape::pair_mask & 2) != 0 ? key_big_pair<1>(ga, K, kk, p, r, sel)
                                                             : key_big_pair<2>(ga, K, kk, p, r, sel);
                best = min(best, (k & (kKeepReg ? keep_reg : keep)) | (static_cast<uint32_t>(NS + i) + sel));
                continue;
            }
            if(((Shape::pair_mask & 1) != 0 && i == x0 + 1) || ((Shape::pair_mask & 2) != 0 && i == y0 + 1) ||
               ((Shape::pair_mask & 4) != 0 && i == z0 + 1)) {
                continue; // the -C sphere of a pair: answered together with its partner
            }
            float const kb = Shape::uniform_k ? cs.big_geo[0].k : cs.big_geo[i].k;
            if(i < Shape::big_x) {
                k = key_big_axis<0, Shape::uniform_k>(cs.axis_coef[2 * i], cs.axis_coef[2 * i + 1], kb, p, r);
            }
            else if(i < Shape::big_x + Shape::big_y) {
                k = key_big_axis<1, Shape::uniform_k>(cs.axis_coef[2 * i], cs.axis_coef[2 * i + 1], kb, p, r);
            }
            else if(i < Shape::big_x + Shape::big_y + Shape::big_z) {
                k = key_big_axis<2, Shape::uniform_k>(cs.axis_coef[2 * i], cs.axis_coef[2 * i + 1], kb, p, r);
            }
            else if(i < Shape::big_near) {
                k = key_big<false, kRobust>(cs.big_geo[i], p, r, p.last == NS + i);
            }
            else {
                k = key_big<true, kRobust>(cs.big_geo[i], p, r, p.last == NS + i);
            }
            take(k, NS + i);
        }
        if constexpr(Shape::embed) {
            id = static_cast<int>(best & ~keep);
            best &= keep;
        }
    }
    else {
        int const nsn = gl.bvh_nodes != nullptr ? 0 : cs.n_small_near, ns = cs.n_small;
        int const nbn = cs.n_big_near, nb = cs.n_big;
#pragma unroll 4
        for(int i = 0; i < nsn; ++i) {
            uint32_t const k = key_small<false, true>(gl.small_geo[i], p, r, p.last == i);
            if(k < best) {
                best = k;
                id = i;
            }
        }
        for(int i = gl.bvh_nodes != nullptr ? ns : nsn; i < ns; ++i) {
            uint32_t const k = key_small<true, true>(gl.small_geo[i], p, r, p.last == i);
            if(k < best) {
                best = k;
                id = i;
            }
        }
        for(int i = 0; i < nbn; ++i) {
            uint32_t const k = key_big<false, true>(gl.big_geo[i], p, r, p.last == ns + i);
            if(k < best) {
                best = k;
                id = ns + i;
            }
        }
        for(int i = nbn; i < nb; ++i) {
            uint32_t const k = key_big<true, true>(gl.big_geo[i], p, r, p.last == ns + i);
            if(k < best) {
                best = k;
                id = ns + i;
            }
        }
        // the hierarchy LAST: the few huge spheres (a ground, walls) have given an upper bound by now, and the traversal
        // drops every box the ray enters beyond it.  Equal roots still go to the lower list position (the traversal's own rule).
        if(gl.bvh_nodes != nullptr) {
            bvh_closest_hit(gl, p, r, best, id);
        }
    }
    id_out = id;
    t_out = __uint_as_float(best) + r.eps;
    return best < kNoHitBits;
}

struct BounceCounters
{
    uint32_t rays;
    uint32_t diffuse;
    uint32_t specular;
    uint32_t dielectric;
};

// hit_record.cpp:6 -- outward UNIT normal norm(P - centre); sa = (-c/R, 1/R).
// (P - c) / R alone is unit only if P lies exactly on the sphere, and P = o + t d with t solved for |d| = 1: a
// deviation delta of |d|^2 puts P off the surface by t delta / 2, the mirror formula with the resulting non-unit
// normal multiplies delta by ~4 t / R (small sphere hit from afar) or ~17 (inside a glass ball, far root
// 2 (c - o).d), and a chain of such bounces overflowed binary32 within ten bounces on the 10 001-sphere scene.
// The reference normalises here too; with a unit normal the mirror formula preserves |d| and delta stays at
// rounding level for the whole path.  MUFU.RSQ rather than a Newton step about 1: its error does not depend on
// how far off the surface P is, so nothing feeds back.
__device__ __forceinline__ void unit_normal(float px, float py, float pz, float4 const& sa, float& nx, float& ny, float& nz)
{
    nx = fmaf(px, sa.w, sa.x);
    ny = fmaf(py, sa.w, sa.y);
    nz = fmaf(pz, sa.w, sa.z);
    float const inv = fast_rsqrt(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
    nx *= inv;
    ny *= inv;
    nz *= inv;
}

// specular_ray, main.cpp:60-67: mirror about the OUTWARD normal; the length of the
// reference's direction is unchanged by it (p.len stays).  The reference then draws one
// uniform and multiplies it by fuzziness = 0 -- the draw must still advance the stream.
__device__ __forceinline__ void reflect_ray(PathF32& p, float nx, float ny, float nz)
{
    float const dn2 = 2.0f * fmaf(nx, p.dx, fmaf(ny, p.dy, nz * p.dz));
    float const rx = fmaf(-dn2, nx, p.dx);
    float const ry = fmaf(-dn2, ny, p.dy);
    float const rz = fmaf(-dn2, nz, p.dz);
    p.dx = rx;
    p.dy = ry;
    p.dz = rz;
    (void)rng_next32(p.rng);
}

// specular_ray for a ray that stands on its hit point (p.o = P): mirror about m = (P - c) / R WITHOUT normalising it --
// r = d - 2 (m.d) / (m.m) m is the exact reflection whatever |m| is (so nothing feeds a deviation of |m| back into the
// direction, the failure unit_normal guards against), and it costs a reciprocal and two products where the unit normal
// costs a reciprocal square root and three.  Every FP32 variant mirrors through this one function.
__device__ __forceinline__ void mirror_at_hit(PathF32& p, float4 const& sa)
{
    float const mx = fmaf(p.ox, sa.w, sa.x), my = fmaf(p.oy, sa.w, sa.y), mz = fmaf(p.oz, sa.w, sa.z);
    float const mm = fmaf(mx, mx, fmaf(my, my, mz * mz));
    float const md = fmaf(mx, p.dx, fmaf(my, p.dy, mz * p.dz));
    float const sc = -2.0f * md * fast_rcp(mm);
    p.dx = fmaf(sc, mx, p.dx);
    p.dy = fmaf(sc, my, p.dy);
    p.dz = fmaf(sc, mz, p.dz);
    (void)rng_next32(p.rng); // the dead "fuzz" draw of main.cpp:65
}

// ---- one iteration of the bounce loop of main.cpp:111-155 AFTER the closest-hit query ----------
// Split in two so that the wavefront variant can run the halves in different kernels:
//   shade_common : sky on a miss, hit record, emission, Russian roulette, throughput
//   scatter_*    : the three material functions of main.cpp:141-154

// Returns true while the path is alive; on false p.er/eg/eb hold the path's radiance.
// On true: p.o = hit point, (nx,ny,nz) = outward normal, refl = material tag.
__device__ __forceinline__ bool shade_common(PathF32& p, bool hit, float t, int id, ShadePlanes const& sp, float& nx,
                                             float& ny, float& nz, int& refl)
{
    if(!hit) {
        // main.cpp:116-119 sky gradient on the unit direction
        float const tt = 0.5f * (p.dy + 1.0f);
        float const omt = 1.0f - tt;
        p.er = fmaf(p.tr, fmaf(0.5f, tt, omt), p.er);
        p.eg = fmaf(p.tg, fmaf(0.7f, tt, omt), p.eg);
        p.eb = fmaf(p.tb, omt + tt, p.eb);
        return false;
    }

    float4 const sa = sp.a[id];
    float4 const sb = sp.b[id];

    // hit_record.cpp:5-9
    float const hx = fmaf(p.dx, t, p.ox);
    float const hy = fmaf(p.dy, t, p.oy);
    float const hz = fmaf(p.dz, t, p.oz);
    unit_normal(hx, hy, hz, sa, nx, ny, nz);

    // main.cpp:126
    p.er = fmaf(p.tr, sb.x, p.er);
    p.eg = fmaf(p.tg, sb.y, p.eg);
    p.eb = fmaf(p.tb, sb.z, p.eb);

    // main.cpp:128-139 Russian roulette with p = max(color), survivor weight color/p
    float4 col;
    if(p.depth > kRouletteThreshold) {
        float4 const sc = sp.c[id];
        if(!(rng_uniform_f32(p.rng) < sc.w)) {
            return false;
        }
        col = sp.d[id];
    }
    else {
        col = sp.c[id];
    }
    p.tr *= col.x;
    p.tg *= col.y;
    p.tb *= col.z;

    p.ox = hx;
    p.oy = hy;
    p.oz = hz;
    p.last = id;
    refl = __float_as_int(sb.w) & 0xff; // bit 8 flags an emitter (ptb_mega_sorted.cuh)
    return true;
}

// diffuse_ray, main.cpp:44-58: cosine-weighted about the FRONT-FACING normal (fx,fy,fz)
__device__ __forceinline__ void scatter_diffuse(PathF32& p, float fx, float fy, float fz)
{
    float const u1 = rng_uniform_f32(p.rng);
    float const u2 = rng_uniform_f32(p.rng);
    float sphi, cphi;
    __sincosf(6.283185307179586f * u1, &sphi, &cphi);
    float const sin_t = fast_sqrt(u2);
    float const cos_t = fast_sqrt(1.0f - u2);
    // u = norm((|w.x| > 0.1 ? (0,1,0) : (1,0,0)) x w), v = w x u
    // v = w x u written out per branch, the two-term components with the fused operand pinned (see gen_primary)
    float ux, uy, uz, vx, vy, vz;
    if(fabsf(fx) > 0.1f) {
        float const inv = fast_rsqrt(fmaf(fz, fz, __fmul_rn(fx, fx)));
        ux = fz * inv;
        uy = 0.0f;
        uz = -fx * inv;
        vx = fy * uz;
        vy = fmaf(fz, ux, -__fmul_rn(fx, uz));
        vz = -(fy * ux);
    }
    else {
        float const inv = fast_rsqrt(fmaf(fz, fz, __fmul_rn(fy, fy)));
        ux = 0.0f;
        uy = -fz * inv;
        uz = fy * inv;
        vx = fmaf(fy, uz, -__fmul_rn(fz, uy));
        vy = -(fx * uz);
        vz = fx * uy;
    }
    float const cu = cphi * sin_t, cv = sphi * sin_t;
    p.dx = fmaf(ux, cu, fmaf(vx, cv, fx * cos_t));
    p.dy = fmaf(uy, cu, fmaf(vy, cv, fy * cos_t));
    p.dz = fmaf(uz, cu, fmaf(vz, cv, fz * cos_t));
    p.len = 1.0f; // main.cpp:54-55 normalises the new direction
}

// dielectric_ray, main.cpp:69-97, refraction index 2.0; (nx,ny,nz) outward normal, dn = n.d
__device__ __forceinline__ void scatter_dielectric(PathF32& p, float nx, float ny, float nz, float dn)
{
    bool const front = dn < 0.0f; // hit_record.cpp:7
    float const fx = front ? nx : -nx, fy = front ? ny : -ny, fz = front ? nz : -nz;
    float const ratio = front ? 0.5f : 2.0f;
    float const cos_t = fminf(fabsf(dn), 1.0f); // (-unit_d).normal, the normal faces the ray
    float const sin_t = fast_sqrt(fmaxf(0.0f, fmaf(-cos_t, cos_t, 1.0f)));
    bool reflect = ratio * sin_t > 1.0f;
    if(!reflect) {
        // Schlick, r0 = ((1-n)/(1+n))^2 = 1/9 for n = 2 and n = 1/2 alike
        float const m = 1.0f - cos_t;
        float const m2 = m * m;
        float const refl_prob = fmaf(8.0f / 9.0f, m2 * m2 * m, 1.0f / 9.0f);
        reflect = refl_prob > rng_uniform_f32(p.rng);
    }
    if(reflect) {
        reflect_ray(p, nx, ny, nz);
    }
    else {
        float const px = fmaf(fx, cos_t, p.dx) * ratio;
        float const py = fmaf(fy, cos_t, p.dy) * ratio;
        float const pz = fmaf(fz, cos_t, p.dz) * ratio;
        float const par = -fast_sqrt(fabsf(1.0f - fmaf(px, px, fmaf(py, py, pz * pz))));
        p.dx = fmaf(fx, par, px);
        p.dy = fmaf(fy, par, py);
        p.dz = fmaf(fz, par, pz);
        p.len = 1.0f; // r_out_perp + r_out_parallel is a unit vector (main.cpp:93-96)
    }
}

// Both halves back to back: what the megakernel and the probe run.
template<bool kCount>
__device__ __forceinline__ bool shade_bounce(PathF32& p, bool hit, float t, int id, ShadePlanes const& sp,
                                             BounceCounters& cnt)
{
    float nx, ny, nz;
    int refl;
    if(!shade_common(p, hit, t, id, sp, nx, ny, nz, refl)) {
        return false;
    }
    if(refl == 1) {
        if(kCount) {
            cnt.specular++;
        }
        mirror_at_hit(p, sp.a[id]);
    }
    else {
        float const dn = fmaf(nx, p.dx, fmaf(ny, p.dy, nz * p.dz));
        if(refl == 0) {
            if(kCount) {
                cnt.diffuse++;
            }
            bool const front = dn < 0.0f; // hit_record.cpp:7
            scatter_diffuse(p, front ? nx : -nx, front ? ny : -ny, front ? nz : -nz);
        }
        else {
            if(kCount) {
                cnt.dielectric++;
            }
            scatter_dielectric(p, nx, ny, nz, dn);
        }
    }
    p.depth++;
    return p.depth < kDepthLimit; // main.cpp:111
}

} // namespace ptb
