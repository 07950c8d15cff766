// ptb_api.cpp -- the GPU-backed entry points of include/ptb200.h: context, scene
// packing, render / resolve orchestration.  Host C++ only; kernels live in
// ptb_f32.cu (throughput), ptb_f64.cu (parity), ptb_resolve.cu.
//
// There is deliberately no CPU path in this file: every compute entry point needs a
// live context, and ptb_create fails when no CUDA device is usable.
#include "../../include/ptb200.h"
#include "ptb_bvh.hpp"
#include "ptb_context.hpp"
#include "ptb_rng.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

// The PRECOMPILED FP32 kernels read the packed sphere lists from ONE __constant__ symbol per device (ptb_f32.cu: c_scene),
// filled stream-ordered at the start of every render.  Two contexts on the same GPU driven from two host threads would
// overwrite it under each other's running kernel, so those launches take turns (the lock is held until the kernel is done).
// Everything else -- the run-time compiled kernels (coefficients as literals, cameras in the kernel argument) and the FP64
// kernels (scene in per-context global memory) -- shares nothing and does not lock.
static std::mutex& device_render_mutex(int device)
{
    static std::mutex m[64];
    return m[static_cast<unsigned>(device) % 64u];
}

using namespace ptb;


namespace {

thread_local std::string g_create_error;

int fail_cuda(ptb_context* ctx, cudaError_t e, char const* what)
{
    ctx->err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return PTB_ERR_CUDA;
}

int fail(ptb_context* ctx, int code, char const* what)
{
    ctx->err = what;
    return code;
}

#define PTB_CUDA(ctx, call) \
    do { \
        cudaError_t const e_ = (call); \
        if(e_ != cudaSuccess) { \
            return fail_cuda((ctx), e_, #call); \
        } \
    } while(0)

float4* active_accum(ptb_context* ctx)
{
    return ctx->ext_accum != nullptr ? ctx->ext_accum : ctx->d_accum;
}

// Geometry class: spheres whose radius dwarfs the distances that matter need the
// 2R-normalised form in binary32 (ptb_scene.cuh).  Absolute error of the classic
// form near the surface is ~3e-8 * R in t; the reference's own absolute scale is
// epsilon = 1e-4 (constants.hpp:7), so R > 32 moves to the stable form.
constexpr double kBigRadius = 32.0;

// Choose the frame of the FP32 scene.  Default origin: centroid of the ordinary-sized spheres
// (the region rays live in), so binary32 coordinates stay O(scene extent).  Refinement: if, on
// some axis, every big sphere that does NOT extend along that axis has the same coordinate
// (true for the five R = 1e6 walls of the box scenes, whose centres all lie on the three lines
// through (0, 0, -1)), use that coordinate instead: those centres then lie exactly ON the frame
// axes and take the one-component test (key_big_axis).
void pack_scene(ptb_context* ctx)
{
    int const n = ctx->n;
    std::vector<RawSphere> const& s = ctx->h_spheres;

    double c[3] = { 0, 0, 0 };
    int m = 0;
    for(int i = 0; i < n; ++i) {
        if(std::fabs(s[i].radius) <= kBigRadius) {
            c[0] += s[i].px;
            c[1] += s[i].py;
            c[2] += s[i].pz;
            ++m;
        }
    }
    if(m > 0) {
        for(double& v : c) {
            v /= m;
        }
    }
    else if(ctx->have_camera) {
        for(int a = 0; a < 3; ++a) {
            c[a] = ctx->h_camera.pos[a];
        }
    }
    for(int a = 0; a < 3; ++a) {
        bool have = false, same = true;
        double v = 0.0;
        for(int i = 0; i < n; ++i) {
            if(std::fabs(s[i].radius) <= kBigRadius) {
                continue;
            }
            double const coord = a == 0 ? s[i].px : (a == 1 ? s[i].py : s[i].pz);
            if(std::fabs(coord) >= 1e-3 * std::fabs(s[i].radius)) {
                continue; // the sphere extends along this axis
            }
            if(!have) {
                v = coord;
                have = true;
            }
            else if(coord != v) {
                same = false;
            }
        }
        // only if the candidate keeps the frame near the action
        if(have && same && std::fabs(v - c[a]) <= 4.0 * kBigRadius) {
            c[a] = v;
        }
    }
    ctx->shift[0] = c[0];
    ctx->shift[1] = c[1];
    ctx->shift[2] = c[2];
}

struct PackedScene
{
    std::vector<SmallGeo> small_geo; // near-only first, then both-roots
    std::vector<BigGeo> big_geo;     // near-only first, then both-roots
    std::vector<int> order;          // list position -> original index
    std::vector<float4> shade;       // 4 planes, by list position
    SceneCounts counts{};
    BvhTree bvh;                     // over the small spheres, leaf_order in LIST POSITIONS (empty: not built)
    std::vector<SmallGeo> bvh_geo;   // leaf order
    std::vector<int> bvh_pos;        // leaf slot -> list position | both-roots flag
};

// A sphere can only ever be hit at its NEAR root when no ray origin can lie inside it:
// it must be opaque (diffuse / specular scatter back to the outside, main.cpp:44-67) and the
// camera lens (position +- the largest lens offset, camera.cpp:34-35: |rd*(s+t)| <= 2*sqrt(2)*lens_radius)
// must be outside it.  Dielectric spheres are traversed from inside and keep both roots.
// Overlapping spheres (BASELINE config 5 is full of them) change nothing: a hit point on one sphere that lay inside
// an opaque neighbour would have been hidden by that neighbour, so no ray ever STARTS inside an opaque sphere.
bool near_root_only(RawSphere const& s, ptb_context const* ctx)
{
    if(s.reflection == 2 || (!ctx->have_camera && !ctx->have_sbcam)) {
        return false;
    }
    auto outside = [&](double const* pos, double slack) {
        double const dx = pos[0] - s.px, dy = pos[1] - s.py, dz = pos[2] - s.pz;
        return std::sqrt(dx * dx + dy * dy + dz * dz) > s.radius + slack + 1e-9;
    };
    bool ok = true;
    if(ctx->have_camera) {
        ok = ok && outside(ctx->h_camera.pos, 3.0 * std::fabs(ctx->h_camera.lens_radius));
    }
    if(ctx->have_sbcam) {
        // rays start at position + push * unit direction, direction within the field of view: test the
        // whole segment's end region conservatively through both end points and the push distance
        double const* c = ctx->sb_cam8;
        double const len = std::sqrt(c[3] * c[3] + c[4] * c[4] + c[5] * c[5]);
        double const start[3] = { c[0] + c[7] * c[3] / len, c[1] + c[7] * c[4] / len, c[2] + c[7] * c[5] / len };
        // the pushed origins lie within push * (half the image plane extent) of `start`:
        // half width = fov * aspect / 2, half height = fov / 2 (sandbox/main.cpp:236-237,258-260); 20 % margin.
        // The aspect ratio is only known once the image is set: assume a very wide image until then.
        double const aspect = ctx->width > 0 ? static_cast<double>(ctx->width) / ctx->height : 4.0;
        double const half = 0.5 * std::fabs(c[6]) * std::sqrt(aspect * aspect + 1.0);
        ok = ok && outside(start, 1.2 * std::fabs(c[7]) * half);
    }
    return ok;
}

PackedScene pack_geometry(ptb_context* ctx)
{
    int const n = ctx->n;
    // the reference only ever squares the radius (sphere.cpp:11) and normalises P - centre (hit_record.cpp:6): its sign
    // does not matter there, and must not here, where the normal is scaled by 1/R
    std::vector<RawSphere> s = ctx->h_spheres;
    for(RawSphere& q : s) {
        q.radius = std::fabs(q.radius);
    }
    double const* sh = ctx->shift;
    PackedScene out;

    // list order: [small near-only, small both, big near-only (x-axis, y-axis, z-axis, other), big both],
    // original order inside each
    std::vector<int> lists[4];
    std::vector<int> big_near[4]; // x, y, z, other
    double k_first = 0.0;
    bool uniform_k = true, any_big = false;

    // tree over the ordinary-sized spheres, in the shifted FP32 frame the kernels use
    std::vector<int> small_ids;
    std::vector<BvhSphere> small_f32;
    for(int i = 0; i < n; ++i) {
        if(s[i].radius <= kBigRadius) {
            small_ids.push_back(i);
            small_f32.push_back(BvhSphere{ static_cast<float>(s[i].px - sh[0]), static_cast<float>(s[i].py - sh[1]),
                                           static_cast<float>(s[i].pz - sh[2]), static_cast<float>(s[i].radius) });
        }
    }

    for(int i = 0; i < n; ++i) {
        bool const big = s[i].radius > kBigRadius;
        bool const near_only = near_root_only(s[i], ctx);
        if(big) {
            double const k = 1.0 / (2.0 * s[i].radius);
            if(!any_big) {
                k_first = k;
                any_big = true;
            }
            else if(k != k_first) {
                uniform_k = false;
            }
        }
        if(big && near_only) {
            double const x = s[i].px - sh[0], y = s[i].py - sh[1], z = s[i].pz - sh[2];
            int const axis = (y == 0.0 && z == 0.0) ? 0 : ((x == 0.0 && z == 0.0) ? 1 : ((x == 0.0 && y == 0.0) ? 2 : 3));
            big_near[axis].push_back(i);
        }
        else {
            lists[(big ? 2 : 0) + (near_only ? 0 : 1)].push_back(i);
        }
    }
    // Mirror-image pairs: an axis group of exactly two spheres with opposite centres and one radius (checked on the
    // FP32 coefficients the kernel will read); the +C sphere goes first (ptb_path_f32.cuh: key_big_pair).
    int pair_mask = 0;
    for(int a = 0; a < 3; ++a) {
        std::vector<int>& l = big_near[a];
        if(l.size() != 2) {
            continue;
        }
        auto const coef = [&](int i, float& g, float& K) {
            double const R = s[i].radius, k = 1.0 / (2.0 * R);
            double const c[3] = { s[i].px - sh[0], s[i].py - sh[1], s[i].pz - sh[2] };
            g = static_cast<float>(-k * c[a]);
            K = static_cast<float>(k * ((c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) - R * R));
        };
        float g0, K0, g1, K1;
        coef(l[0], g0, K0);
        coef(l[1], g1, K1);
        if(s[l[0]].radius == s[l[1]].radius && g0 == -g1 && g0 != 0.0f && K0 == K1) {
            if(g0 > 0.0f) { // g = -k c: the +C sphere has the negative g
                std::swap(l[0], l[1]);
            }
            pair_mask |= 1 << a;
        }
    }
    for(auto const& l : big_near) {
        lists[2].insert(lists[2].end(), l.begin(), l.end());
    }
    out.counts.small_near = static_cast<int>(lists[0].size());
    out.counts.small_both = static_cast<int>(lists[1].size());
    out.counts.big_near = static_cast<int>(lists[2].size());
    out.counts.big_both = static_cast<int>(lists[3].size());
    out.counts.big_x = static_cast<int>(big_near[0].size());
    out.counts.big_y = static_cast<int>(big_near[1].size());
    out.counts.big_z = static_cast<int>(big_near[2].size());
    out.counts.uniform_k = any_big && uniform_k;
    out.counts.pair_mask = out.counts.uniform_k ? pair_mask : 0;
    // index-in-key truncates t by < 2^-19 relative: allowed while 2e-6 * (scene extent) << epsilon = 1e-4.
    // Extent = reach of the ordinary spheres and the cameras from the frame origin.
    double extent = 0.0;
    for(int i = 0; i < n; ++i) {
        if(s[i].radius <= kBigRadius) {
            double const x = s[i].px - sh[0], y = s[i].py - sh[1], z = s[i].pz - sh[2];
            extent = std::max(extent, std::sqrt(x * x + y * y + z * z) + s[i].radius);
        }
    }
    auto reach = [&](double const* q) {
        double const x = q[0] - sh[0], y = q[1] - sh[1], z = q[2] - sh[2];
        return std::sqrt(x * x + y * y + z * z);
    };
    if(ctx->have_camera) {
        extent = std::max(extent, reach(ctx->h_camera.pos));
    }
    if(ctx->have_sbcam) {
        extent = std::max(extent, reach(ctx->sb_cam8));
    }
    // A ray INSIDE a huge sphere (camera inside it, or a huge glass ball) can travel its whole diameter: the paths are then
    // not confined to `extent` at all, and the plain root formulas lose the next hit of a ray that starts on that sphere
    // (found by dev/fuzz_scenes.py: 1.5 % of the rays inside a R = 1000 ground escaped to the sky).  Such scenes take the
    // full-precision keys and the exact self-sphere roots, like the sandbox scene does.
    // The same goes, on a smaller scale, for an ordinary sphere that rays travel INSIDE of (glass, or the camera sits in
    // it): a ray that starts on it sees its own surface at c / 2 half_b with c = r^2 * 1e-7 of rounding noise, which passes
    // epsilon = 1e-4 for grazing rays once r is a few units -- and inside a mirror ball a grazing ray stays grazing (a
    // camera in a r = 5 mirror ball traced 22 % more rays than the oracle's arithmetic).  The reference's glass balls have
    // r <= 0.5; one unit is the limit for the plain formulas.
    double max_inside_radius = 0.0, min_inside_radius = 1.0;
    bool inside_a_mirror = false;
    for(int i : lists[1]) {
        max_inside_radius = std::max(max_inside_radius, s[i].radius);
        min_inside_radius = std::min(min_inside_radius, s[i].radius);
        inside_a_mirror = inside_a_mirror || s[i].reflection == 1;
    }
    // A camera inside a MIRROR ball: paths are chains of up to 99 reflections off the same concave surface, and the index
    // riding in the key's low mantissa bits truncates every hit distance the same way -- 2^-19 t, always short.  Glass keeps a
    // ray inside for a bounce or two and the error is noise; here it accumulates into a drift of the whole chord pattern
    // (dev/fuzz_scenes.py seed 51 scene 1: a r = 0.0016 lamp inside a r = 0.25 mirror ball lit 23 % of the pixels instead of
    // the 1 % of the FP64 oracle and of the full-precision keys).
    // At the other end, rays bouncing INSIDE a ball only some tens of epsilon across (a r = 0.006 mirror ball within the
    // lens' reach: 1.8 % more mirror hits than the FP64 kernel, whose counts the exact-self-root path reproduces to the
    // last ray) are decided by how a near-grazing chord of length ~epsilon is rounded: r >= 0.05 = 500 epsilon.
    // ... and a sphere much smaller than the rounding noise of |o - c|^2 (1e-7 of a few units squared) would be hit by
    // rays that pass it at several radii: the plain discriminant needs r >= 1e-3
    double min_radius = 1.0;
    for(int i = 0; i < n; ++i) {
        if(s[i].radius > 0.0) {
            min_radius = std::min(min_radius, s[i].radius);
        }
    }
    out.counts.embed_ok = extent <= 8.0 && out.counts.big_both == 0 && max_inside_radius <= 1.0 && min_inside_radius >= 0.05 &&
                          min_radius >= 1e-3 && !inside_a_mirror;
    if(!out.counts.embed_ok) {
        // the paired and the axis tests exist for the index-in-key kernels only: they have no exact self-sphere root
        out.counts.pair_mask = 0;
        out.counts.big_x = out.counts.big_y = out.counts.big_z = 0;
    }
    out.counts.fits_const = out.counts.small_near + out.counts.small_both <= kMaxConstSpheres &&
                            out.counts.big_near + out.counts.big_both <= kMaxConstSpheres;
    for(auto const& l : lists) {
        out.order.insert(out.order.end(), l.begin(), l.end());
    }

    size_t const N = static_cast<size_t>(std::max(n, 1));
    out.shade.resize(4 * N);
    for(int pos = 0; pos < n; ++pos) {
        int const i = out.order[static_cast<size_t>(pos)];
        double const R = s[i].radius;
        double const x = s[i].px - sh[0], y = s[i].py - sh[1], z = s[i].pz - sh[2];
        if(R > kBigRadius) {
            double const k = 1.0 / (2.0 * R);
            BigGeo b{};
            b.gx = static_cast<float>(-k * x);
            b.gy = static_cast<float>(-k * y);
            b.gz = static_cast<float>(-k * z);
            b.k = static_cast<float>(k);
            b.K = static_cast<float>(k * ((x * x + y * y + z * z) - R * R));
            b.two_r = static_cast<float>(2.0 * R);
            out.big_geo.push_back(b);
        }
        else {
            SmallGeo g{};
            g.cx = static_cast<float>(x);
            g.cy = static_cast<float>(y);
            g.cz = static_cast<float>(z);
            g.r2 = static_cast<float>(R * R);
            if(!(g.r2 > 0.0f)) {
                // a sphere of radius 0 (or one whose r^2 underflows binary32) is never hit in the reference either -- its
                // discriminant is -|perp|^2 -- but in binary32 the rounding noise of that difference is not: say "never"
                g.r2 = -1.0e30f;
            }
            out.small_geo.push_back(g);
        }
        double const inv_r = R > 1e-30 ? 1.0 / R : 1.0;
        double const p = std::max({ s[i].cr, s[i].cg, s[i].cb });
        double const inv_p = p > 0.0 ? 1.0 / p : 0.0;
        float4 a, b, c, d;
        a.x = static_cast<float>(-x * inv_r);
        a.y = static_cast<float>(-y * inv_r);
        a.z = static_cast<float>(-z * inv_r);
        a.w = static_cast<float>(inv_r);
        b.x = static_cast<float>(s[i].er);
        b.y = static_cast<float>(s[i].eg);
        b.z = static_cast<float>(s[i].eb);
        bool const emits = s[i].er != 0.0 || s[i].eg != 0.0 || s[i].eb != 0.0;
        int32_t const refl = s[i].reflection | (emits ? 0x100 : 0); // kEmissiveBit, ptb_mega_sorted.cuh
        std::memcpy(&b.w, &refl, sizeof(float));
        c.x = static_cast<float>(s[i].cr);
        c.y = static_cast<float>(s[i].cg);
        c.z = static_cast<float>(s[i].cb);
        c.w = static_cast<float>(p);
        d.x = static_cast<float>(s[i].cr * inv_p);
        d.y = static_cast<float>(s[i].cg * inv_p);
        d.z = static_cast<float>(s[i].cb * inv_p);
        d.w = static_cast<float>(p); // so that the roulette needs one plane: (color/p, p)
        size_t const P = static_cast<size_t>(pos);
        out.shade[P] = a;
        out.shade[N + P] = b;
        out.shade[2 * N + P] = c;
        out.shade[3 * N + P] = d;
    }

    // Device hierarchy: only when the small spheres do not fit the constant lists (otherwise the unrolled or the
    // constant-bank scan is faster than any traversal).  Leaves carry LIST POSITIONS.
    int const n_small = out.counts.small_near + out.counts.small_both;
    BvhTree small_tree = n_small > kMaxConstSpheres ? build_bvh(small_f32) : BvhTree{};
    if(n_small > kMaxConstSpheres && small_tree.max_depth <= kBvhStack) {
        std::vector<int> list_pos_of(static_cast<size_t>(n), -1); // original index -> list position
        for(int pos = 0; pos < n; ++pos) {
            list_pos_of[static_cast<size_t>(out.order[static_cast<size_t>(pos)])] = pos;
        }
        out.bvh = std::move(small_tree);
        for(int j : out.bvh.leaf_order) { // j = index into small_ids
            int const pos = list_pos_of[static_cast<size_t>(small_ids[static_cast<size_t>(j)])];
            out.bvh_geo.push_back(out.small_geo[static_cast<size_t>(pos)]);
            out.bvh_pos.push_back(pos | (pos >= out.counts.small_near ? static_cast<int>(0x80000000u) : 0));
        }
    }
    return out;
}

void pack_camera(ptb_context* ctx)
{
    RawCamera const& c = ctx->h_camera;
    double const* sh = ctx->shift;
    CameraF32& o = ctx->cams.cam;
    o.px = static_cast<float>(c.pos[0] - sh[0]);
    o.py = static_cast<float>(c.pos[1] - sh[1]);
    o.pz = static_cast<float>(c.pos[2] - sh[2]);
    o.rx = static_cast<float>(c.llc[0] - c.pos[0]);
    o.ry = static_cast<float>(c.llc[1] - c.pos[1]);
    o.rz = static_cast<float>(c.llc[2] - c.pos[2]);
    o.ax = static_cast<float>(c.ax[0]);
    o.ay = static_cast<float>(c.ax[1]);
    o.az = static_cast<float>(c.ax[2]);
    o.bx = static_cast<float>(c.ay[0]);
    o.by = static_cast<float>(c.ay[1]);
    o.bz = static_cast<float>(c.ay[2]);
    o.lens_radius = static_cast<float>(c.lens_radius);
    o.inv_w = ctx->width > 0 ? static_cast<float>(1.0 / ctx->width) : 0.0f;
    o.inv_h = ctx->height > 0 ? static_cast<float>(1.0 / ctx->height) : 0.0f;
    o.sub_len = ctx->ns > 0 ? static_cast<float>(1.0 / ctx->ns) : 0.0f;

    // sandbox/main.cpp:235-237: cam.d normalised, cx = (w * fov / h, 0, 0), cy = norm(cx x d) * fov
    SmallptCamF32& sbc = ctx->cams.sbcam;
    sbc = SmallptCamF32{};
    if(ctx->have_sbcam && ctx->width > 0) {
        double const* c8 = ctx->sb_cam8;
        double const len = std::sqrt(c8[3] * c8[3] + c8[4] * c8[4] + c8[5] * c8[5]);
        double const d[3] = { c8[3] / len, c8[4] / len, c8[5] / len };
        double const cxx = ctx->width * c8[6] / ctx->height;
        double cy[3] = { 0.0 * d[2] - 0.0 * d[1], 0.0 * d[0] - cxx * d[2], cxx * d[1] - 0.0 * d[0] };
        double const cyl = std::sqrt(cy[0] * cy[0] + cy[1] * cy[1] + cy[2] * cy[2]);
        sbc.ox = static_cast<float>(c8[0] - sh[0]);
        sbc.oy = static_cast<float>(c8[1] - sh[1]);
        sbc.oz = static_cast<float>(c8[2] - sh[2]);
        sbc.dx = static_cast<float>(d[0]);
        sbc.dy = static_cast<float>(d[1]);
        sbc.dz = static_cast<float>(d[2]);
        sbc.cxx = static_cast<float>(cxx);
        sbc.cyx = static_cast<float>(cy[0] / cyl * c8[6]);
        sbc.cyy = static_cast<float>(cy[1] / cyl * c8[6]);
        sbc.cyz = static_cast<float>(cy[2] / cyl * c8[6]);
        sbc.push = static_cast<float>(c8[7]);
        sbc.inv_w = static_cast<float>(1.0 / ctx->width);
        sbc.inv_h = static_cast<float>(1.0 / ctx->height);
    }
}

// (Re)build everything derived from (spheres, camera, image geometry) and push the
// global-memory parts to the device.
int rebuild_device_scene(ptb_context* ctx)
{
    pack_scene(ctx);
    PackedScene const ps = pack_geometry(ctx);
    pack_camera(ctx);
    int const n = ctx->n;
    ctx->counts = ps.counts;
    int const n_small = ps.counts.small_near + ps.counts.small_both;
    int const n_big = ps.counts.big_near + ps.counts.big_both;
    ctx->cs.n_small_near = ps.counts.small_near;
    ctx->cs.n_small = n_small;
    ctx->cs.n_big_near = ps.counts.big_near;
    ctx->cs.n_big = n_big;
    ctx->cs.n_total = n;
    if(ps.counts.fits_const) {
        for(int i = 0; i < n_small; ++i) {
            ctx->cs.small_geo[i] = ps.small_geo[static_cast<size_t>(i)];
        }
        int const n_axis = ps.counts.big_x + ps.counts.big_y + ps.counts.big_z;
        for(int i = 0; i < n_big; ++i) {
            BigGeo const& b = ps.big_geo[static_cast<size_t>(i)];
            ctx->cs.big_geo[i] = b;
            // axis spheres: the one non-zero component of g (x group first, then y, then z) and K
            int const axis = i < ps.counts.big_x ? 0 : (i < ps.counts.big_x + ps.counts.big_y ? 1 : 2);
            ctx->cs.axis_coef[2 * i] = i < n_axis ? (axis == 0 ? b.gx : (axis == 1 ? b.gy : b.gz)) : 0.0f;
            ctx->cs.axis_coef[2 * i + 1] = b.K;
        }
        for(int i = 0; i < n; ++i) {
            ctx->cs.order[i] = ps.order[static_cast<size_t>(i)];
        }
    }

    size_t const cap = static_cast<size_t>(std::max(n, 1));
    if(cap > ctx->geo_cap) {
        cudaFree(ctx->d_small);
        cudaFree(ctx->d_big);
        cudaFree(ctx->d_order);
        cudaFree(ctx->d_shade);
        ctx->d_small = nullptr;
        ctx->d_big = nullptr;
        ctx->d_order = nullptr;
        ctx->d_shade = nullptr;
        ctx->geo_cap = 0;
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_small, cap * sizeof(SmallGeo)));
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_big, cap * sizeof(BigGeo)));
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_order, cap * sizeof(int)));
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_shade, 4 * cap * sizeof(float4)));
        ctx->geo_cap = cap;
    }
    cudaStream_t const st = ctx->stream;
    if(n_small > 0) {
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_small, ps.small_geo.data(), ps.small_geo.size() * sizeof(SmallGeo),
                                      cudaMemcpyHostToDevice, st));
    }
    if(n_big > 0) {
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_big, ps.big_geo.data(), ps.big_geo.size() * sizeof(BigGeo),
                                      cudaMemcpyHostToDevice, st));
    }
    if(n > 0) {
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_order, ps.order.data(), ps.order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_shade, ps.shade.data(), 4 * static_cast<size_t>(n) * sizeof(float4),
                                      cudaMemcpyHostToDevice, st));
    }
    ctx->have_bvh = !ps.bvh_geo.empty();
    ctx->bvh_depth = ps.bvh.max_depth;
    ctx->bvh_node_count = static_cast<int>(ps.bvh.nodes.size());
    if(ctx->have_bvh) {
        size_t const nn = std::max<size_t>(ps.bvh.nodes.size(), 1), nl = ps.bvh_geo.size();
        if(nn > ctx->bvh_node_cap) {
            cudaFree(ctx->d_bvh_nodes);
            ctx->d_bvh_nodes = nullptr;
            ctx->bvh_node_cap = 0;
            PTB_CUDA(ctx, cudaMalloc(&ctx->d_bvh_nodes, nn * sizeof(BvhNode64)));
            ctx->bvh_node_cap = nn;
        }
        if(nl > ctx->bvh_leaf_cap) {
            cudaFree(ctx->d_bvh_geo);
            cudaFree(ctx->d_bvh_pos);
            ctx->d_bvh_geo = nullptr;
            ctx->d_bvh_pos = nullptr;
            ctx->bvh_leaf_cap = 0;
            PTB_CUDA(ctx, cudaMalloc(&ctx->d_bvh_geo, nl * sizeof(SmallGeo)));
            PTB_CUDA(ctx, cudaMalloc(&ctx->d_bvh_pos, nl * sizeof(int)));
            ctx->bvh_leaf_cap = nl;
        }
        if(!ps.bvh.nodes.empty()) {
            PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_bvh_nodes, ps.bvh.nodes.data(), ps.bvh.nodes.size() * sizeof(BvhNode64),
                                          cudaMemcpyHostToDevice, st));
        }
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_bvh_geo, ps.bvh_geo.data(), nl * sizeof(SmallGeo), cudaMemcpyHostToDevice, st));
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_bvh_pos, ps.bvh_pos.data(), nl * sizeof(int), cudaMemcpyHostToDevice, st));
        ctx->bvh_root = ps.bvh.root;
    }
    PTB_CUDA(ctx, cudaStreamSynchronize(st)); // ps goes out of scope
    return PTB_OK;
}

// rebuild_device_scene for callers that already HAVE a scene (camera / image changes): a failure invalidates it
int rebuild_or_invalidate(ptb_context* ctx)
{
    int const rc = rebuild_device_scene(ctx);
    if(rc != PTB_OK) {
        ctx->have_scene = false;
    }
    return rc;
}

ShadePlanes shade_planes(ptb_context* ctx)
{
    size_t const N = static_cast<size_t>(ctx->n);
    ShadePlanes sp;
    sp.a = ctx->d_shade;
    sp.b = ctx->d_shade + N;
    sp.c = ctx->d_shade + 2 * N;
    sp.d = ctx->d_shade + 3 * N;
    return sp;
}

GeoLists geo_lists(ptb_context* ctx, bool use_bvh = true)
{
    GeoLists g{};
    g.small_geo = ctx->d_small;
    g.big_geo = ctx->d_big;
    g.order = ctx->d_order;
    if(use_bvh && ctx->have_bvh) {
        g.bvh_nodes = ctx->d_bvh_nodes;
        g.bvh_geo = ctx->d_bvh_geo;
        g.bvh_pos = ctx->d_bvh_pos;
        g.bvh_root = ctx->bvh_root;
    }
    return g;
}

int require_ready(ptb_context* ctx, bool need_image, bool smallpt = false, bool any_camera = false)
{
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(!ctx->have_scene) {
        return fail(ctx, PTB_ERR_STATE, "no scene: call ptb_upload_scene first");
    }
    if(any_camera ? (!ctx->have_camera && !ctx->have_sbcam) : (smallpt ? !ctx->have_sbcam : !ctx->have_camera)) {
        return fail(ctx, PTB_ERR_STATE, smallpt ? "no camera: call ptb_set_smallpt_camera first" : "no camera: call ptb_set_camera first");
    }
    if(need_image && (ctx->width <= 0 || active_accum(ctx) == nullptr)) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    return PTB_OK;
}

} // namespace

namespace ptb {
float4* api_active_accum(ptb_context* ctx)
{
    return active_accum(ctx);
}
int api_fail(ptb_context* ctx, int code, char const* what)
{
    return fail(ctx, code, what);
}
int api_fail_cuda(ptb_context* ctx, cudaError_t e, char const* what)
{
    return fail_cuda(ctx, e, what);
}
int api_exception(ptb_context* ctx, char const* entry) noexcept
{
    // called from a catch(...) handler: rethrow to learn what it was
    char const* what = "unknown C++ exception";
    int code = PTB_ERR_INTERNAL;
    char buffer[256];
    try {
        throw;
    }
    catch(std::bad_alloc const&) {
        what = "out of host memory";
        code = PTB_ERR_MEMORY;
    }
    catch(std::exception const& e) {
        std::snprintf(buffer, sizeof buffer, "%s", e.what());
        what = buffer;
    }
    catch(...) {
    }
    try {
        std::string& err = ctx != nullptr ? ctx->err : g_create_error;
        err = std::string(entry) + ": " + what;
    }
    catch(...) { // no memory for the message either: the code still says what happened
    }
    return code;
}
} // namespace ptb

// A handle made by ptb_create_multi stands for several member contexts: the call is theirs (ptb_multi.cpp)
#define PTB_GROUP(ctx, call) \
    do { \
        if((ctx) != nullptr && (ctx)->group != nullptr) { \
            return (call); \
        } \
    } while(0)

extern "C" {

int ptb_device_count(void)
try {
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}
PTB_CATCH(nullptr, "ptb_device_count")

char const* ptb_last_error(ptb_context const* ctx)
{
    return ctx != nullptr ? ctx->err.c_str() : g_create_error.c_str();
}

int ptb_create(int device, ptb_context** out)
try {
    if(out == nullptr) {
        g_create_error = "ptb_create: out is null";
        return PTB_ERR_ARGUMENT;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        g_create_error = "ptb_create: no usable CUDA device (this library has no CPU path)";
        return PTB_ERR_NO_DEVICE;
    }
    if(device < 0 || device >= n) {
        g_create_error = "ptb_create: device index out of range";
        return PTB_ERR_NO_DEVICE;
    }
    auto* ctx = new ptb_context{};
    ctx->device = device;
    auto bail = [&](cudaError_t err, char const* what) {
        g_create_error = std::string("ptb_create: ") + what + ": " + cudaGetErrorString(err);
        ptb_destroy(ctx);
        return PTB_ERR_CUDA;
    };
    if((e = cudaSetDevice(device)) != cudaSuccess) {
        return bail(e, "cudaSetDevice");
    }
    if((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
        return bail(e, "cudaDeviceGetAttribute");
    }
    if((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        return bail(e, "cudaStreamCreate");
    }
    ctx->stream = ctx->own_stream;
    if((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        return bail(e, "cudaEventCreate");
    }
    if((e = cudaMalloc(&ctx->d_counters, sizeof(DeviceCounters))) != cudaSuccess) {
        return bail(e, "cudaMalloc(counters)");
    }
    if((e = cudaMemset(ctx->d_counters, 0, sizeof(DeviceCounters))) != cudaSuccess) {
        return bail(e, "cudaMemset(counters)");
    }
    if((e = cudaMalloc(&ctx->d_camera, sizeof(RawCamera))) != cudaSuccess) {
        return bail(e, "cudaMalloc(camera)");
    }
    *out = ctx;
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_create")

void ptb_destroy(ptb_context* ctx)
{
    if(ctx == nullptr) {
        return;
    }
    if(ctx->group != nullptr) {
        multi_destroy(ctx);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    comm_destroy(ctx); // collective when the context is a rank of a job: peers unmap before anything is freed
    if(ctx->own_stream != nullptr) {
        cudaStreamSynchronize(ctx->own_stream);
    }
    cudaFree(ctx->d_spheres);
    cudaFree(ctx->d_camera);
    cudaFree(ctx->d_sbcam8);
    cudaFree(ctx->d_small);
    cudaFree(ctx->d_big);
    cudaFree(ctx->d_order);
    cudaFree(ctx->d_shade);
    cudaFree(ctx->d_bvh_nodes);
    cudaFree(ctx->d_bvh_geo);
    cudaFree(ctx->d_bvh_pos);
    cudaFree(ctx->d_accum);
    cudaFree(ctx->d_accum64);
    cudaFree(ctx->d_rgb);
    cudaFree(ctx->d_rgb8);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->wf.planes);
    cudaFree(ctx->wf.words);
    cudaFree(ctx->wf.counters);
    if(ctx->ev0 != nullptr) {
        cudaEventDestroy(ctx->ev0);
    }
    if(ctx->ev1 != nullptr) {
        cudaEventDestroy(ctx->ev1);
    }
    if(ctx->own_stream != nullptr) {
        cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
}

int ptb_set_stream(ptb_context* ctx, void* cuda_stream)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_stream: a multi-GPU context runs on its members' private streams");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = static_cast<cudaStream_t>(cuda_stream); // nullptr == the legacy default stream
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_set_stream")

int ptb_reset_stream(ptb_context* ctx)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) {
        return PTB_OK;
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_reset_stream")

int ptb_synchronize(ptb_context* ctx)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_synchronize(ctx));
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_synchronize")

int ptb_upload_scene(ptb_context* ctx, void const* spheres, size_t count, size_t stride)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_upload_scene(ctx, spheres, count, stride));
    // count == 0 is a scene: every ray misses and sees the sky (main.cpp:114-120), spheres may then be null
    if((spheres == nullptr && count != 0) || stride < PTB_SPHERE_BYTES || count >= (1u << 24)) {
        // list positions travel in 24 bits next to the path's depth, and 0xFFFFFF means "starts on no sphere" (ptb_mega_sorted.cuh: kNoSphere)
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_upload_scene: need spheres of stride >= 88 bytes (fewer than 2^24 of them)");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    // from here until everything derived from the new list is on the device there is NO usable scene: a CUDA failure
    // half way (out of memory on a re-upload) must not leave the previous upload's flag over new counts and freed buffers
    ctx->have_scene = false;
    ctx->h_spheres.resize(count);
    auto const* src = static_cast<unsigned char const*>(spheres);
    for(size_t i = 0; i < count; ++i) {
        std::memcpy(static_cast<void*>(&ctx->h_spheres[i]), src + i * stride, PTB_SPHERE_BYTES);
        ctx->h_spheres[i].pad_ = 0;
        int const r = ctx->h_spheres[i].reflection;
        if(r < 0 || r > 2) {
            ctx->h_spheres.clear();
            ctx->have_scene = false;
            return fail(ctx, PTB_ERR_ARGUMENT, "ptb_upload_scene: reflection must be 0 (diffuse), 1 (specular) or 2 (dielectric)");
        }
        // radius, position, emission, colour: the reference would render NaN pixels from a non-finite one; here it would
        // also reach the host-side packing (frame choice, hierarchy build), so it is an argument error
        double v[10];
        std::memcpy(v, &ctx->h_spheres[i], sizeof(v));
        for(double x : v) {
            if(!std::isfinite(x)) {
                ctx->h_spheres.clear();
                ctx->have_scene = false;
                return fail(ctx, PTB_ERR_ARGUMENT, "ptb_upload_scene: a sphere holds a non-finite number");
            }
        }
    }
    ctx->n = static_cast<int>(count);
    if(count > ctx->d_spheres_cap) {
        cudaFree(ctx->d_spheres);
        ctx->d_spheres = nullptr;
        ctx->d_spheres_cap = 0;
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_spheres, count * sizeof(RawSphere)));
        ctx->d_spheres_cap = count;
    }
    if(count > 0) {
        PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_spheres, ctx->h_spheres.data(), count * sizeof(RawSphere),
                                      cudaMemcpyHostToDevice, ctx->stream));
    }
    {
        // first guess for the sorted megakernel: surface seen by a ray ~ radius^2, walls (huge spheres) capped
        double w[3] = { 0.0, 0.0, 0.0 };
        for(RawSphere const& sp : ctx->h_spheres) {
            double const r = std::min(std::fabs(sp.radius), 4.0);
            w[sp.reflection] += r * r;
        }
        ctx->inline_material = w[1] >= w[0] ? 1 : 0;
    }
    int const rc = rebuild_device_scene(ctx);
    ctx->have_scene = rc == PTB_OK;
    return rc;
}
PTB_CATCH(ctx, "ptb_upload_scene")

int ptb_set_camera(ptb_context* ctx, void const* camera, size_t bytes)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_set_camera(ctx, camera, bytes));
    if(camera == nullptr || bytes != PTB_CAMERA_BYTES) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_camera: need a 176-byte pt::camera");
    }
    {
        double v[PTB_CAMERA_BYTES / sizeof(double)];
        std::memcpy(v, camera, sizeof(v));
        for(double x : v) {
            if(!std::isfinite(x)) {
                return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_camera: the camera holds a non-finite number");
            }
        }
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::memcpy(static_cast<void*>(&ctx->h_camera), camera, sizeof(RawCamera));
    PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_camera, &ctx->h_camera, sizeof(RawCamera), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_camera = true;
    if(ctx->have_scene) {
        return rebuild_or_invalidate(ctx);
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_set_camera")

int ptb_set_smallpt_camera(ptb_context* ctx, double const* cam8)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_set_smallpt_camera(ctx, cam8));
    bool finite8 = cam8 != nullptr;
    for(int i = 0; finite8 && i < 8; ++i) {
        finite8 = std::isfinite(cam8[i]);
    }
    if(!finite8 || (cam8[3] == 0.0 && cam8[4] == 0.0 && cam8[5] == 0.0)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_smallpt_camera: need position(3), non-zero direction(3), fov factor, push");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::memcpy(ctx->sb_cam8, cam8, sizeof(ctx->sb_cam8));
    if(ctx->d_sbcam8 == nullptr) {
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_sbcam8, sizeof(ctx->sb_cam8)));
    }
    PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_sbcam8, ctx->sb_cam8, sizeof(ctx->sb_cam8), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_sbcam = true;
    if(ctx->have_scene) {
        return rebuild_or_invalidate(ctx);
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_set_smallpt_camera")

int ptb_set_image(ptb_context* ctx, int width, int height, int num_subpixels)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(width <= 0 || height <= 0 || num_subpixels <= 0 || num_subpixels > 8 ||
       static_cast<unsigned long long>(width) * static_cast<unsigned long long>(height) *
               static_cast<unsigned long long>(num_subpixels * num_subpixels) >=
           (1ull << 31)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_image: bad geometry (need 1 <= num_subpixels <= 8 and < 2^31 sub-pixels)");
    }
    PTB_GROUP(ctx, multi_set_image(ctx, width, height, num_subpixels));
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(ctx->comm != nullptr) {
        // one process per GPU: other ranks may hold mappings of the buffers about to be freed (collective)
        int const rc = comm_release_peers(ctx);
        if(rc != PTB_OK) {
            return rc;
        }
    }
    ctx->buffers_epoch++;
    // from here until every buffer of the new geometry exists there is NO image: an allocation that fails half way
    // (out of memory at 8K) must not leave the new size over missing buffers -- require_ready() then says "no image"
    ctx->width = 0;
    ctx->height = 0;
    ctx->nslots = static_cast<size_t>(width) * static_cast<size_t>(height) * static_cast<size_t>(num_subpixels * num_subpixels);
    size_t const bytes = ctx->nslots * sizeof(float4);
    if(bytes > ctx->d_accum_bytes) {
        cudaFree(ctx->d_accum);
        ctx->d_accum = nullptr;
        ctx->d_accum_bytes = 0;
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_accum, bytes));
        ctx->d_accum_bytes = bytes;
    }
    cudaFree(ctx->d_accum64);
    ctx->d_accum64 = nullptr;
    ctx->accum64_used = false;
    cudaFree(ctx->d_rgb);
    cudaFree(ctx->d_rgb8);
    ctx->d_rgb = nullptr;
    ctx->d_rgb8 = nullptr;
    size_t const npix = static_cast<size_t>(width) * static_cast<size_t>(height);
    PTB_CUDA(ctx, cudaMalloc(&ctx->d_rgb, npix * 3 * sizeof(double)));
    PTB_CUDA(ctx, cudaMalloc(&ctx->d_rgb8, npix * 3));
    if(ctx->ext_accum != nullptr && ctx->ext_accum_bytes < bytes) {
        ctx->ext_accum = nullptr;
        ctx->ext_accum_bytes = 0;
    }
    ctx->width = width;
    ctx->height = height;
    ctx->ns = num_subpixels;
    if(ctx->have_scene && ctx->have_sbcam) {
        int const rc = rebuild_or_invalidate(ctx); // near-root-only classification depends on the aspect ratio
        if(rc != PTB_OK) {
            return rc;
        }
    }
    pack_camera(ctx);
    return ptb_clear(ctx);
}
PTB_CATCH(ctx, "ptb_set_image")

int ptb_clear(ptb_context* ctx)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_clear(ctx));
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    if(active_accum(ctx) != nullptr && ctx->nslots > 0) {
        PTB_CUDA(ctx, cudaMemsetAsync(active_accum(ctx), 0, ctx->nslots * sizeof(float4), ctx->stream));
    }
    if(ctx->d_accum64 != nullptr) {
        PTB_CUDA(ctx, cudaMemsetAsync(ctx->d_accum64, 0, ctx->nslots * 4 * sizeof(double), ctx->stream));
    }
    ctx->accum64_used = false;
    PTB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(DeviceCounters), ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->seen_diffuse = ctx->seen_specular = 0;
    uint64_t const launches = ctx->stats.kernel_launches;
    ctx->stats = ptb_stats{};
    ctx->stats.kernel_launches = launches;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_clear")

int ptb_render(ptb_context* ctx, uint64_t seed, uint32_t first_sample, uint32_t samples_per_subpixel, uint32_t flags)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    uint32_t const variant = flags & PTB_VARIANT_MASK;
    uint32_t const precision = flags & PTB_PRECISION_MASK;
    uint32_t const integrator = flags & PTB_INTEGRATOR_MASK;
    uint32_t const accel = flags & PTB_ACCEL_MASK;
    uint32_t const codegen = flags & PTB_CODEGEN_MASK;
    if((flags & ~(PTB_VARIANT_MASK | PTB_PRECISION_MASK | PTB_INTEGRATOR_MASK | PTB_ACCEL_MASK | PTB_CODEGEN_MASK)) != 0 ||
       variant > PTB_VARIANT_MEGAKERNEL_SORTED || (accel != PTB_ACCEL_AUTO && accel != PTB_ACCEL_SCAN) ||
       (codegen != PTB_CODEGEN_AUTO && codegen != PTB_CODEGEN_PRECOMPILED) ||
       (precision != PTB_PRECISION_FP32 && precision != PTB_PRECISION_FP64) ||
       (integrator != PTB_INTEGRATOR_PT && integrator != PTB_INTEGRATOR_SMALLPT)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_render: unknown flags");
    }
    bool const smallpt = integrator == PTB_INTEGRATOR_SMALLPT;
    PTB_GROUP(ctx, multi_render(ctx, seed, first_sample, samples_per_subpixel, flags));
    int rc = require_ready(ctx, true, smallpt);
    if(rc != PTB_OK) {
        return rc;
    }
    if(smallpt && (variant != PTB_VARIANT_MEGAKERNEL || ctx->ns != 2)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_render: the smallpt integrator is megakernel-only and uses 2x2 sub-pixels (sandbox/main.cpp:248-249)");
    }
    if(variant == PTB_VARIANT_WAVEFRONT && precision == PTB_PRECISION_FP64) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_render: the wavefront variant exists in FP32 only");
    }
    if(static_cast<uint64_t>(first_sample) + samples_per_subpixel > 0xFFFFFFFFull) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_render: sample range overflows 32 bits");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    if(ctx->comm != nullptr) {
        // one process per GPU: this rank traces its share of the range (ptb_sample_share); the rest is the same call
        uint32_t first = 0, count = 0;
        int32_t info[6];
        ptb_comm_info(ctx, info);
        ptb_sample_share(samples_per_subpixel, info[0], info[1], &first, &count);
        first_sample += first;
        samples_per_subpixel = count;
    }
    if(precision == PTB_PRECISION_FP64 && ctx->d_accum64 == nullptr) {
        // (before the early return: with several GPUs a member whose share is empty still takes part in the sum)
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_accum64, ctx->nslots * 4 * sizeof(double)));
        PTB_CUDA(ctx, cudaMemsetAsync(ctx->d_accum64, 0, ctx->nslots * 4 * sizeof(double), ctx->stream));
        ctx->buffers_epoch++;
    }
    if(precision == PTB_PRECISION_FP64) {
        ctx->accum64_used = true;
    }
    if(samples_per_subpixel == 0) {
        ctx->stats.last_render_ms = 0.0;
        return PTB_OK; // the reference renders a black image for spp < 4 (main.cpp:206)
    }
    ctx->last_launch_jit = false; // the FP64 and the wavefront kernels are never compiled at run time
    cudaStream_t st = ctx->stream;
    if(variant == PTB_VARIANT_WAVEFRONT && st == nullptr) {
        // CUDA graphs cannot be captured on the legacy default stream: drain it and use the private one;
        // ptb_render blocks until the work is done, so later work on the default stream is still ordered
        PTB_CUDA(ctx, cudaStreamSynchronize(nullptr));
        st = ctx->own_stream;
    }
    uint64_t const key = seed_key(seed);
    int launches = 0;

    // Scene-specialised code: compiled (once per scene, ~0.4 s) BEFORE the timed region starts.
    int sorted_inline = ctx->inline_material;
    if(ctx->have_bvh && accel != PTB_ACCEL_SCAN) {
        // Scenes behind the hierarchy scatter NOTHING in place: then every lane takes a new ray after every bounce, all 32
        // from the top of READY -- the batch the camera or the scatter stage pushed last, rays of one kind that descend
        // the tree alike (measured on config 5: -9.3 % time against diffuse in place, -7.4 % against mirror in place)
        sorted_inline = -1;
    }
    if(char const* force = std::getenv("PTB_INLINE_MATERIAL")) { // experiments (dev/)
        int const f = std::atoi(force);
        sorted_inline = f < 0 ? -1 : f == 0 ? 0 : 1;
    }
    JitKernel const* jit_kernel = nullptr;
    // Any layout that fits the unrolled scan qualifies, not only the ones the library ships precompiled: a scene of up to
    // 16 spheres (32 when the index cannot ride in the key) that has no precompiled specialisation would otherwise take
    // the run-time-count kernel.
    {
        SceneCounts const& c = ctx->counts;
        int const total = c.small_near + c.small_both + c.big_near + c.big_both;
        bool const unrollable = c.fits_const && total >= 1 && total <= (c.embed_ok ? 16 : 32);
        if((variant == PTB_VARIANT_MEGAKERNEL_SORTED || variant == PTB_VARIANT_MEGAKERNEL) && precision == PTB_PRECISION_FP32 &&
           codegen == PTB_CODEGEN_AUTO && ctx->n <= kSmemShadeSpheres && unrollable) {
            // Compiling costs 0.2-0.3 s and buys ~5 %: it pays for itself after ~5 s of rendering.  A single call that large
            // (>= 2^35 paths) compiles at once; otherwise the kernel is compiled when the same scene is rendered a second
            // time (progressive rendering, benchmarks, animations with a static scene).
            bool const eager = static_cast<uint64_t>(ctx->nslots) * samples_per_subpixel >= (1ull << 35);
            JitCache::Kind const kind = variant == PTB_VARIANT_MEGAKERNEL_SORTED ? JitCache::kSorted
                                        : (smallpt ? JitCache::kInPlaceSmallpt : JitCache::kInPlacePt);
            jit_kernel = ctx->jit.get(ctx->cs, ctx->counts, kind, sorted_inline, eager);
        }
    }

    // only a launch that reads the per-device constant symbol has to take turns with other contexts on this GPU
    std::unique_lock<std::mutex> turn(device_render_mutex(ctx->device), std::defer_lock);
    if(precision == PTB_PRECISION_FP32 && jit_kernel == nullptr) {
        turn.lock();
    }
    PTB_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    if(precision == PTB_PRECISION_FP64) {
        if(smallpt) {
            PTB_CUDA(ctx, launch_smallpt_render_f64(key, first_sample, samples_per_subpixel, static_cast<uint32_t>(ctx->width),
                                                    static_cast<uint32_t>(ctx->height), ctx->d_spheres, ctx->n, ctx->d_sbcam8,
                                                    ctx->d_accum64, ctx->d_counters, st));
        }
        else {
            PTB_CUDA(ctx, launch_render_f64(key, first_sample, samples_per_subpixel, static_cast<uint32_t>(ctx->width),
                                            static_cast<uint32_t>(ctx->height), static_cast<uint32_t>(ctx->ns),
                                            ctx->d_spheres, ctx->n, ctx->d_camera, ctx->d_accum64, ctx->d_counters, st));
        }
        launches = 1;
    }
    else {
        if(jit_kernel == nullptr) {
            PTB_CUDA(ctx, upload_const_scene(ctx->cs, st));
        }
        RenderParamsF32 p{};
        p.cams = ctx->cams;
        p.key = key;
        p.first_sample = first_sample;
        p.samples = samples_per_subpixel;
        p.width = static_cast<uint32_t>(ctx->width);
        p.height = static_cast<uint32_t>(ctx->height);
        p.ns = static_cast<uint32_t>(ctx->ns);
        p.nslots = static_cast<uint32_t>(ctx->nslots);
        p.ngroups = (p.nslots + 31u) / 32u;
        // tile = 32 slots x chunk samples; aim for >= 16 tiles per resident warp so the
        // tail of the launch is short, but keep tiles >= 8 samples deep
        uint64_t const resident_warps = static_cast<uint64_t>(ctx->sm_count) * 64ull;
        uint32_t chunk = samples_per_subpixel;
        while(chunk > 8u && static_cast<uint64_t>(p.ngroups) * ((samples_per_subpixel + chunk - 1u) / chunk) < 16ull * resident_warps) {
            chunk = (chunk + 1u) / 2u;
        }
        p.chunk = chunk;
        p.nchunks = (samples_per_subpixel + chunk - 1u) / chunk;
        p.ntiles = p.ngroups * p.nchunks;
        p.accum = active_accum(ctx);
        p.counters = ctx->d_counters;
        p.shade = shade_planes(ctx);
        p.geo = geo_lists(ctx, accel != PTB_ACCEL_SCAN);
        p.n_total = ctx->n;
        p.key_mask = ~((1u << 4) - 1u); // kIdBits of ptb_path_f32.cuh
        if(variant == PTB_VARIANT_WAVEFRONT) {
            uint64_t const items = static_cast<uint64_t>(p.nslots) * p.samples;
            uint32_t const pool = static_cast<uint32_t>(std::min<uint64_t>(1ull << 22, std::max<uint64_t>(items, 1024)));
            if(ctx->wf.pool < pool) {
                cudaFree(ctx->wf.planes);
                cudaFree(ctx->wf.words);
                ctx->wf.planes = nullptr;
                ctx->wf.words = nullptr;
                ctx->wf.pool = 0;
                PTB_CUDA(ctx, cudaMalloc(&ctx->wf.planes, kWfPlanesPerPool * static_cast<size_t>(pool) * sizeof(float4)));
                PTB_CUDA(ctx, cudaMalloc(&ctx->wf.words, kWfWordsPerPool * static_cast<size_t>(pool) * sizeof(uint32_t)));
                if(ctx->wf.counters == nullptr) {
                    PTB_CUDA(ctx, cudaMalloc(&ctx->wf.counters, sizeof(WavefrontCounters)));
                }
                ctx->wf.pool = pool;
            }
            WavefrontBuffers buf = ctx->wf;
            buf.pool = pool;
            PTB_CUDA(ctx, launch_wavefront(buf, p, ctx->counts, ctx->sm_count, st, &launches));
        }
        else if(variant == PTB_VARIANT_MEGAKERNEL_SORTED) {
            ctx->last_launch_jit = jit_kernel != nullptr;
            if(jit_kernel != nullptr) {
                PTB_CUDA(ctx, ctx->jit.launch(*jit_kernel, p, ctx->sm_count, st, &launches));
            }
            else {
                PTB_CUDA(ctx, launch_megakernel_sorted(p, ctx->counts, ctx->sm_count, st, &launches, sorted_inline));
            }
        }
        else {
            ctx->last_launch_jit = jit_kernel != nullptr;
            if(jit_kernel != nullptr) {
                PTB_CUDA(ctx, ctx->jit.launch(*jit_kernel, p, ctx->sm_count, st, &launches));
            }
            else {
                PTB_CUDA(ctx, launch_megakernel(p, ctx->counts, ctx->sm_count, st, &launches, smallpt));
            }
        }
    }
    PTB_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    PTB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.0f;
    PTB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.last_render_ms = ms;
    ctx->stats.total_render_ms += ms;
    ctx->stats.kernel_launches += static_cast<uint64_t>(launches);
    ctx->stats.paths += static_cast<uint64_t>(ctx->nslots) * samples_per_subpixel;
    if(variant == PTB_VARIANT_MEGAKERNEL_SORTED && precision == PTB_PRECISION_FP32) {
        // feedback for the next launch: which of diffuse / specular did this one hit more often (32 bytes, stream is idle)
        DeviceCounters c{};
        PTB_CUDA(ctx, cudaMemcpy(&c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        unsigned long long const nd = c.diffuse - std::min(c.diffuse, ctx->seen_diffuse);
        unsigned long long const nsp = c.specular - std::min(c.specular, ctx->seen_specular);
        if(nd + nsp > 0) {
            ctx->inline_material = nsp >= nd ? 1 : 0;
        }
        ctx->seen_diffuse = c.diffuse;
        ctx->seen_specular = c.specular;
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_render")

static int resolve_common(ptb_context* ctx, double* rgb_out, uint8_t* rgb8_out)
{
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_resolve(ctx, rgb_out, rgb8_out, nullptr));
    if(ctx->comm != nullptr) {
        return comm_resolve(ctx, rgb_out, rgb8_out, nullptr);
    }
    int rc = require_ready(ctx, true, false, true);
    if(rc != PTB_OK) {
        return rc;
    }
    if(rgb_out == nullptr && rgb8_out == nullptr) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_resolve: output pointer is null");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t const st = ctx->stream;
    size_t const npix = static_cast<size_t>(ctx->width) * static_cast<size_t>(ctx->height);
    PTB_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    PTB_CUDA(ctx, launch_resolve(active_accum(ctx), ctx->accum64_used ? ctx->d_accum64 : nullptr,
                                 static_cast<uint32_t>(ctx->width), static_cast<uint32_t>(ctx->height),
                                 static_cast<uint32_t>(ctx->ns), rgb_out != nullptr ? ctx->d_rgb : nullptr,
                                 rgb8_out != nullptr ? ctx->d_rgb8 : nullptr, st));
    PTB_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->stats.kernel_launches += 1;
    if(rgb_out != nullptr) {
        PTB_CUDA(ctx, cudaMemcpyAsync(rgb_out, ctx->d_rgb, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if(rgb8_out != nullptr) {
        PTB_CUDA(ctx, cudaMemcpyAsync(rgb8_out, ctx->d_rgb8, npix * 3, cudaMemcpyDeviceToHost, st));
    }
    PTB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.0f;
    PTB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.last_resolve_ms = ms;
    return PTB_OK;
}

int ptb_resolve(ptb_context* ctx, double* rgb_out)
try {
    return resolve_common(ctx, rgb_out, nullptr);
}
PTB_CATCH(ctx, "ptb_resolve")

int ptb_resolve_rgb8(ptb_context* ctx, uint8_t* rgb8_out)
try {
    return resolve_common(ctx, nullptr, rgb8_out);
}
PTB_CATCH(ctx, "ptb_resolve_rgb8")

int ptb_resolve_device(ptb_context* ctx, void** device_rgb)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr || ctx->comm != nullptr) {
        if(device_rgb == nullptr) {
            return fail(ctx, PTB_ERR_ARGUMENT, "ptb_resolve_device: output pointer is null");
        }
        return ctx->group != nullptr ? multi_resolve(ctx, nullptr, nullptr, device_rgb) : comm_resolve(ctx, nullptr, nullptr, device_rgb);
    }
    int rc = require_ready(ctx, true, false, true);
    if(rc != PTB_OK) {
        return rc;
    }
    if(device_rgb == nullptr) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_resolve_device: output pointer is null");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, launch_resolve(active_accum(ctx), ctx->accum64_used ? ctx->d_accum64 : nullptr,
                                 static_cast<uint32_t>(ctx->width), static_cast<uint32_t>(ctx->height),
                                 static_cast<uint32_t>(ctx->ns), ctx->d_rgb, nullptr, ctx->stream));
    ctx->stats.kernel_launches += 1;
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *device_rgb = ctx->d_rgb;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_resolve_device")

int ptb_measure_fp32_peak(ptb_context* ctx, double* tflops_out)
try {
    if(ctx == nullptr || tflops_out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, ptb_measure_fp32_peak(multi_root(ctx), tflops_out));
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    float* scratch = nullptr;
    PTB_CUDA(ctx, cudaMalloc(&scratch, sizeof(float)));
    cudaStream_t const st = ctx->stream;
    double flop = 0.0;
    double best_ms = 1e30;
    cudaError_t e = launch_fp32_peak(ctx->sm_count, 256, scratch, st, &flop); // warm-up
    for(int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {
        e = cudaEventRecord(ctx->ev0, st);
        if(e == cudaSuccess) {
            e = launch_fp32_peak(ctx->sm_count, 4096, scratch, st, &flop);
        }
        if(e == cudaSuccess) {
            e = cudaEventRecord(ctx->ev1, st);
        }
        if(e == cudaSuccess) {
            e = cudaStreamSynchronize(st);
        }
        float ms = 0.0f;
        if(e == cudaSuccess) {
            e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        }
        if(e == cudaSuccess && ms < best_ms) {
            best_ms = ms;
        }
    }
    cudaFree(scratch);
    ctx->stats.kernel_launches += 6;
    if(e != cudaSuccess) {
        return fail_cuda(ctx, e, "ptb_measure_fp32_peak");
    }
    *tflops_out = flop / (best_ms * 1e-3) * 1e-12;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_measure_fp32_peak")

int ptb_accum_buffer(ptb_context* ctx, void** device_ptr, size_t* bytes)
try {
    if(ctx == nullptr || device_ptr == nullptr || bytes == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_accum_buffer: a multi-GPU context holds one buffer per GPU and sums them itself");
    }
    if(active_accum(ctx) == nullptr) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    *device_ptr = active_accum(ctx);
    *bytes = ctx->nslots * sizeof(float4);
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_accum_buffer")

int ptb_set_accum_buffer(ptb_context* ctx, void* device_ptr, size_t bytes)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_accum_buffer: a multi-GPU context holds one buffer per GPU and sums them itself");
    }
    if(device_ptr == nullptr) {
        ctx->ext_accum = nullptr;
        ctx->ext_accum_bytes = 0;
        return PTB_OK;
    }
    if(ctx->width <= 0) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(bytes < ctx->nslots * sizeof(float4) || (reinterpret_cast<uintptr_t>(device_ptr) & 15u) != 0) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_set_accum_buffer: need a 16-byte aligned buffer of width*height*ns*ns*16 bytes");
    }
    ctx->ext_accum = static_cast<float4*>(device_ptr);
    ctx->ext_accum_bytes = bytes;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_set_accum_buffer")

int ptb_download_accum(ptb_context* ctx, float* out, size_t floats)
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_download_accum(ctx, out, floats));
    if(active_accum(ctx) == nullptr) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(floats < ctx->nslots * 4) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_download_accum: output too small");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaMemcpyAsync(out, active_accum(ctx), ctx->nslots * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_download_accum")

// ---- checkpoint / resume (include/ptb200.h; the reference's TODO, README.md:9) -------------------------------------
int ptb_upload_accum(ptb_context* ctx, float const* in, size_t floats)
try {
    if(ctx == nullptr || in == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_upload_accum(ctx, in, floats));
    if(active_accum(ctx) == nullptr) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(floats != ctx->nslots * 4) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_upload_accum: need exactly width*height*ns*ns*4 floats (a checkpoint of the same image geometry)");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaMemcpyAsync(active_accum(ctx), in, ctx->nslots * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_upload_accum")

int ptb_download_accum64(ptb_context* ctx, double* out, size_t doubles)
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_download_accum64(ctx, out, doubles));
    if(ctx->width <= 0) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(doubles < ctx->nslots * 4) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_download_accum64: output too small");
    }
    if(ctx->d_accum64 == nullptr) { // no FP64 render since ptb_set_image: the buffer is all zeros
        std::memset(out, 0, ctx->nslots * 4 * sizeof(double));
        return PTB_OK;
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    PTB_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_accum64, ctx->nslots * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_download_accum64")

int ptb_upload_accum64(ptb_context* ctx, double const* in, size_t doubles)
try {
    if(ctx == nullptr || in == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_upload_accum64(ctx, in, doubles));
    if(ctx->width <= 0) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(doubles != ctx->nslots * 4) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_upload_accum64: need exactly width*height*ns*ns*4 doubles");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    if(ctx->d_accum64 == nullptr) {
        PTB_CUDA(ctx, cudaMalloc(&ctx->d_accum64, ctx->nslots * 4 * sizeof(double)));
        ctx->buffers_epoch++;
    }
    PTB_CUDA(ctx, cudaMemcpyAsync(ctx->d_accum64, in, ctx->nslots * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->accum64_used = true;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_upload_accum64")

int ptb_get_stats(ptb_context* ctx, ptb_stats* out)
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, multi_get_stats(ctx, out));
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    DeviceCounters c{};
    PTB_CUDA(ctx, cudaMemcpyAsync(&c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    PTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.rays = c.rays;
    ctx->stats.hits_diffuse = c.diffuse;
    ctx->stats.hits_specular = c.specular;
    ctx->stats.hits_dielectric = c.dielectric;
    *out = ctx->stats;
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_get_stats")

int ptb_scene_layout(ptb_context* ctx, int32_t out[10])
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, ptb_scene_layout(multi_root(ctx), out));
    if(!ctx->have_scene) {
        return fail(ctx, PTB_ERR_STATE, "no scene: call ptb_upload_scene first");
    }
    SceneCounts const& c = ctx->counts;
    int32_t const v[10] = { c.small_near, c.small_both, c.big_near, c.big_both, c.big_x, c.big_y, c.big_z,
                            (c.uniform_k ? 1 : 0) | (c.embed_ok ? 2 : 0), c.fits_const ? 1 : 0,
                            megakernel_has_specialisation(c) ? 1 : 0 };
    std::memcpy(out, v, sizeof(v));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_scene_layout")

int ptb_jit_info(ptb_context* ctx, int32_t out[5])
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, ptb_jit_info(multi_root(ctx), out));
    out[0] = ctx->jit.available() ? 1 : 0;
    out[1] = ctx->jit.compiled();
    out[2] = ctx->jit.failures();
    out[3] = ctx->last_launch_jit ? 1 : 0;
    out[4] = static_cast<int32_t>(ctx->jit.compile_ms());
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_jit_info")

char const* ptb_jit_last_error(ptb_context* ctx)
{
    if(ctx != nullptr && ctx->group != nullptr) {
        return ptb_jit_last_error(multi_root(ctx));
    }
    return ctx != nullptr ? ctx->jit.last_error().c_str() : "";
}

int ptb_trace_samples(ptb_context* ctx, uint64_t seed, uint32_t const* x, uint32_t const* y, uint32_t const* sx,
                      uint32_t const* sy, uint32_t const* sample, size_t count, uint32_t flags, int32_t* primary_hit_out,
                      double* radiance_out, double* ray_out, uint32_t* draws_out)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) { // the parity probe needs one GPU: the root's
        int const rc_ = ptb_trace_samples(multi_root(ctx), seed, x, y, sx, sy, sample, count, flags, primary_hit_out, radiance_out,
                                          ray_out, draws_out);
        if(rc_ != PTB_OK) {
            ctx->err = multi_root(ctx)->err;
        }
        return rc_;
    }
    bool const smallpt = (flags & PTB_INTEGRATOR_MASK) == PTB_INTEGRATOR_SMALLPT;
    int rc = require_ready(ctx, false, smallpt);
    if(rc != PTB_OK) {
        return rc;
    }
    if(ctx->width <= 0) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(smallpt && ctx->ns != 2) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_samples: the smallpt integrator uses 2x2 sub-pixels");
    }
    if(x == nullptr || y == nullptr || sx == nullptr || sy == nullptr || sample == nullptr || primary_hit_out == nullptr ||
       radiance_out == nullptr || count > (1u << 30)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_samples: null pointer");
    }
    uint32_t const precision = flags & PTB_PRECISION_MASK;
    if(precision != PTB_PRECISION_FP32 && precision != PTB_PRECISION_FP64) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_samples: unknown flags");
    }
    std::unique_lock<std::mutex> turn(device_render_mutex(ctx->device), std::defer_lock);
    if(precision == PTB_PRECISION_FP32) {
        turn.lock(); // the FP32 probe is a precompiled kernel: it reads the per-device constant symbol
    }
    for(size_t i = 0; i < count; ++i) {
        if(x[i] >= static_cast<uint32_t>(ctx->width) || y[i] >= static_cast<uint32_t>(ctx->height) ||
           sx[i] >= static_cast<uint32_t>(ctx->ns) || sy[i] >= static_cast<uint32_t>(ctx->ns)) {
            return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_samples: coordinate outside the image");
        }
    }
    if(count == 0) {
        return PTB_OK;
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t const st = ctx->stream;
    uint32_t* d_in = nullptr;
    int32_t* d_hit = nullptr;
    double* d_rad = nullptr;
    double* d_ray = nullptr;
    uint32_t* d_draws = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_in);
        cudaFree(d_hit);
        cudaFree(d_rad);
        cudaFree(d_ray);
        cudaFree(d_draws);
    };
#define PTB_CUDA_T(call) \
    do { \
        cudaError_t const e_ = (call); \
        if(e_ != cudaSuccess) { \
            cleanup(); \
            return fail_cuda(ctx, e_, #call); \
        } \
    } while(0)
    PTB_CUDA_T(cudaMalloc(&d_in, 5 * count * sizeof(uint32_t)));
    PTB_CUDA_T(cudaMalloc(&d_hit, count * sizeof(int32_t)));
    PTB_CUDA_T(cudaMalloc(&d_rad, 3 * count * sizeof(double)));
    PTB_CUDA_T(cudaMalloc(&d_ray, 6 * count * sizeof(double)));
    PTB_CUDA_T(cudaMalloc(&d_draws, count * sizeof(uint32_t)));
    uint32_t const* srcs[5] = { x, y, sx, sy, sample };
    for(int k = 0; k < 5; ++k) {
        PTB_CUDA_T(cudaMemcpyAsync(d_in + static_cast<size_t>(k) * count, srcs[k], count * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, st));
    }
    ProbeParams q{};
    q.key = seed_key(seed);
    q.width = static_cast<uint32_t>(ctx->width);
    q.height = static_cast<uint32_t>(ctx->height);
    q.ns = static_cast<uint32_t>(ctx->ns);
    q.x = d_in;
    q.y = d_in + count;
    q.sx = d_in + 2 * count;
    q.sy = d_in + 3 * count;
    q.sample = d_in + 4 * count;
    q.count = static_cast<uint32_t>(count);
    q.primary_hit = d_hit;
    q.radiance = d_rad;
    q.ray = d_ray;
    q.draws = d_draws;
    q.cams = ctx->cams;
    if(precision == PTB_PRECISION_FP64) {
        if(smallpt) {
            PTB_CUDA_T(launch_smallpt_probe_f64(q, ctx->d_spheres, ctx->n, ctx->d_sbcam8, st));
        }
        else {
            PTB_CUDA_T(launch_probe_f64(q, ctx->d_spheres, ctx->n, ctx->d_camera, st));
        }
    }
    else {
        PTB_CUDA_T(upload_const_scene(ctx->cs, st));
        PTB_CUDA_T(launch_probe_f32(q, ctx->counts, shade_planes(ctx), geo_lists(ctx, (flags & PTB_ACCEL_MASK) != PTB_ACCEL_SCAN), st, smallpt));
    }
    ctx->stats.kernel_launches += 1;
    PTB_CUDA_T(cudaMemcpyAsync(primary_hit_out, d_hit, count * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PTB_CUDA_T(cudaMemcpyAsync(radiance_out, d_rad, 3 * count * sizeof(double), cudaMemcpyDeviceToHost, st));
    if(ray_out != nullptr) {
        PTB_CUDA_T(cudaMemcpyAsync(ray_out, d_ray, 6 * count * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if(draws_out != nullptr) {
        PTB_CUDA_T(cudaMemcpyAsync(draws_out, d_draws, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    PTB_CUDA_T(cudaStreamSynchronize(st));
    if(ray_out != nullptr && precision == PTB_PRECISION_FP32) {
        // FP32 rays live in the shifted frame: report them in world coordinates
        for(size_t i = 0; i < count; ++i) {
            ray_out[6 * i + 0] += ctx->shift[0];
            ray_out[6 * i + 1] += ctx->shift[1];
            ray_out[6 * i + 2] += ctx->shift[2];
        }
    }
    cleanup();
#undef PTB_CUDA_T
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_trace_samples")

int ptb_trace_paths(ptb_context* ctx, uint64_t seed, uint32_t const* x, uint32_t const* y, uint32_t const* sx,
                    uint32_t const* sy, uint32_t const* sample, size_t count, int trail_len, int32_t* trail_out)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, ptb_trace_paths(multi_root(ctx), seed, x, y, sx, sy, sample, count, trail_len, trail_out));
    int rc = require_ready(ctx, false, false);
    if(rc != PTB_OK) {
        return rc;
    }
    if(ctx->width <= 0) {
        return fail(ctx, PTB_ERR_STATE, "no image: call ptb_set_image first");
    }
    if(x == nullptr || y == nullptr || sx == nullptr || sy == nullptr || sample == nullptr || trail_out == nullptr ||
       trail_len < 1 || trail_len > 100 || count > (1u << 26)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_paths: null pointer, or trail_len outside 1..100");
    }
    for(size_t i = 0; i < count; ++i) {
        if(x[i] >= static_cast<uint32_t>(ctx->width) || y[i] >= static_cast<uint32_t>(ctx->height) ||
           sx[i] >= static_cast<uint32_t>(ctx->ns) || sy[i] >= static_cast<uint32_t>(ctx->ns)) {
            return fail(ctx, PTB_ERR_ARGUMENT, "ptb_trace_paths: coordinate outside the image");
        }
    }
    if(count == 0) {
        return PTB_OK;
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t const st = ctx->stream;
    uint32_t* d_in = nullptr;
    int32_t* d_trail = nullptr;
    size_t const cells = count * static_cast<size_t>(trail_len);
    cudaError_t e = cudaMalloc(&d_in, 5 * count * sizeof(uint32_t));
    e = e == cudaSuccess ? cudaMalloc(&d_trail, cells * sizeof(int32_t)) : e;
    uint32_t const* srcs[5] = { x, y, sx, sy, sample };
    for(int k = 0; k < 5 && e == cudaSuccess; ++k) {
        e = cudaMemcpyAsync(d_in + static_cast<size_t>(k) * count, srcs[k], count * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    }
    ProbeParams q{};
    q.key = seed_key(seed);
    q.width = static_cast<uint32_t>(ctx->width);
    q.height = static_cast<uint32_t>(ctx->height);
    q.ns = static_cast<uint32_t>(ctx->ns);
    q.x = d_in;
    q.y = d_in + count;
    q.sx = d_in + 2 * count;
    q.sy = d_in + 3 * count;
    q.sample = d_in + 4 * count;
    q.count = static_cast<uint32_t>(count);
    e = e == cudaSuccess ? launch_trail_f64(q, ctx->d_spheres, ctx->n, ctx->d_camera, d_trail, trail_len, st) : e;
    e = e == cudaSuccess ? cudaMemcpyAsync(trail_out, d_trail, cells * sizeof(int32_t), cudaMemcpyDeviceToHost, st) : e;
    e = e == cudaSuccess ? cudaStreamSynchronize(st) : e;
    cudaFree(d_in);
    cudaFree(d_trail);
    ctx->stats.kernel_launches += 1;
    if(e != cudaSuccess) {
        return fail_cuda(ctx, e, "ptb_trace_paths");
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_trace_paths")

int ptb_rng_draws(ptb_context* ctx, uint64_t seed, uint32_t const* slot, uint32_t const* sample, size_t count,
                  int n_draws, double* draws_out)
try {
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    PTB_GROUP(ctx, ptb_rng_draws(multi_root(ctx), seed, slot, sample, count, n_draws, draws_out));
    if(slot == nullptr || sample == nullptr || draws_out == nullptr || n_draws <= 0 || count == 0 || count > (1u << 28)) {
        return fail(ctx, PTB_ERR_ARGUMENT, "ptb_rng_draws: bad arguments");
    }
    PTB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t const st = ctx->stream;
    uint32_t* d_in = nullptr;
    double* d_out = nullptr;
    size_t const total = count * static_cast<size_t>(n_draws);
    cudaError_t e = cudaMalloc(&d_in, 2 * count * sizeof(uint32_t));
    if(e == cudaSuccess) {
        e = cudaMalloc(&d_out, total * sizeof(double));
    }
    if(e == cudaSuccess) {
        e = cudaMemcpyAsync(d_in, slot, count * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    }
    if(e == cudaSuccess) {
        e = cudaMemcpyAsync(d_in + count, sample, count * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    }
    if(e == cudaSuccess) {
        e = launch_rng_draws(seed_key(seed), d_in, d_in + count, static_cast<uint32_t>(count), n_draws, d_out, st);
    }
    if(e == cudaSuccess) {
        e = cudaMemcpyAsync(draws_out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, st);
    }
    if(e == cudaSuccess) {
        e = cudaStreamSynchronize(st);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    ctx->stats.kernel_launches += 1;
    if(e != cudaSuccess) {
        return fail_cuda(ctx, e, "ptb_rng_draws");
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_rng_draws")

} // extern "C"
