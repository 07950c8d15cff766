// ptb_f32.cu -- FP32 throughput kernels: the persistent megakernel and the FP32 probe.
//
// Replaces the row-parallel CPU loop of /root/reference/src/main.cpp:217-236 (768
// taskflow row tasks, each walking x, sy, sx, s sequentially with one mt19937).
//
// Megakernel design (B200: 148 SMs, 4 schedulers/SM, FP32 issue-bound, no tensor work):
//   * persistent grid: SM count x resident blocks; warps pull WORK TILES
//     (32 consecutive sub-pixel slots x `chunk` samples) from one global cursor, so
//     the tail of the launch is one tile, not one wave;
//   * inside a tile the 32 lanes share a pool of (slot, sample) items.  A lane whose
//     path ended takes the next ready camera sample at the top of the loop ("path
//     regeneration"), so the closest-hit scan -- the bulk of the instructions -- always
//     runs with every lane busy, whatever the spread of path lengths (geometric,
//     mean 12.3 bounces, limit 100 in the box scenes).  Camera samples are generated
//     32 at a time, by all lanes together, into a per-warp shared-memory ring;
//   * the random stream is keyed by (seed, slot, sample), so which lane traces which
//     item does not change the image;
//   * sphere geometry is read as constant-bank operands of the FFMA/FADD
//     instructions (kernels are specialised on the two list lengths and fully
//     unrolled); shading planes sit in shared memory as float4 SoA;
//   * a finished path adds {r,g,b,1} to its slot with ONE 16-byte vector reduction
//     (red.global.add.v4.f32, sm_90+): no read-modify-write, no per-thread image
//     state, progressive by construction.
#include <cstdlib>

#include "ptb_kernels.h"
#include "ptb_path_f32.cuh"
#include "ptb_smallpt_f32.cuh"

namespace ptb {

__constant__ ConstSceneF32 c_scene;

cudaError_t upload_const_scene(ConstSceneF32 const& cs, cudaStream_t stream)
{
    return cudaMemcpyToSymbolAsync(c_scene, &cs, sizeof(cs), 0, cudaMemcpyHostToDevice, stream);
}

} // namespace ptb

#include "ptb_mega_inplace.cuh"

namespace ptb {

// ---- specialisation table ------------------------------------------------------------------
// (small near-only, small both-roots, big near-only, big both-roots) list lengths with a fully
// unrolled kernel.  Everything else runs the generic run-time-count variant.
#define PTB_MEGA_SPECIALISATIONS(X) \
    X(2, 1, 5, 0, 2, 2, 1, true, true, 3)   /* box_scene.hpp / box_mirror_scene.hpp: light + mirror ball | glass ball | 5 R=1e6 walls on the frame axes, left/right and top/bottom as mirror-image pairs */ \
    X(2, 1, 5, 0, 2, 2, 1, true, true, 0)   /* the same without the pairing                                                                          */ \
    X(3, 1, 1, 0, 0, 1, 0, true, true, 0)   /* simple_scene.hpp: mirror, centre, light | glass | ground (R=100, on the y axis)                       */ \
    X(2, 2, 1, 0, 0, 1, 0, true, true, 0)   /* depth-of-field scene (BASELINE config 4): two glass spheres                                             */ \
    X(2, 1, 5, 0, 0, 0, 0, false, true, 0)  /* box scenes in a frame where the walls are not axis spheres                                             */ \
    X(0, 3, 0, 5, 0, 0, 0, false, false, 0) /* box scenes with the camera inside every sphere's reach: all both-roots (full-precision keys)           */ \
    X(3, 1, 0, 6, 0, 0, 0, true, false, 0)  /* sandbox/main.cpp: 2 mirrors + light | glass | six R=1e5 walls seen from INSIDE; ~300 units across     */ \
    X(1, 0, 0, 0, 0, 0, 0, false, true, 0) \
    X(0, 1, 0, 0, 0, 0, 0, false, true, 0) \
    X(8, 0, 0, 0, 0, 0, 0, false, true, 0)

#define PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm) \
    ((c).small_near == (a) && (c).small_both == (b) && (c).big_near == (cc) && (c).big_both == (d) && (c).big_x == (bx) && \
     (c).big_y == (by) && (c).big_z == (bz) && (c).uniform_k == (uk) && (c).embed_ok == (em) && (c).pair_mask == (pm) && \
     (c).fits_const)

bool megakernel_has_specialisation(SceneCounts const& c)
{
#define X(a, b, cc, d, bx, by, bz, uk, em, pm) \
    if(PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm)) { \
        return true; \
    }
    PTB_MEGA_SPECIALISATIONS(X)
#undef X
    return false;
}

template<class Shape, bool kSmem, class Integ>
static cudaError_t launch_one(RenderParamsF32 const& p, int sm_count, cudaStream_t stream)
{
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_kernel<Shape, kSmem, Integ>, kMegaThreads, 0);
    if(e != cudaSuccess) {
        return e;
    }
    if(per_sm < 1) {
        per_sm = 1;
    }
    if(char const* cap = std::getenv("PTB_BLOCKS_PER_SM")) { // occupancy experiments (dev/)
        int const v = std::atoi(cap);
        per_sm = v >= 1 && v < per_sm ? v : per_sm;
    }
    unsigned long long blocks = static_cast<unsigned long long>(sm_count) * static_cast<unsigned long long>(per_sm);
    unsigned long long const blocks_needed =
        (static_cast<unsigned long long>(p.ntiles) + (kMegaThreads / 32) - 1) / (kMegaThreads / 32);
    if(blocks > blocks_needed) {
        blocks = blocks_needed;
    }
    if(blocks < 1) {
        blocks = 1;
    }
    mega_kernel<Shape, kSmem, Integ><<<static_cast<unsigned>(blocks), kMegaThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

template<class Integ>
static cudaError_t launch_mega_integrator(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream)
{
    bool const smem = p.n_total <= kSmemShadeSpheres;
#define X(a, b, cc, d, bx, by, bz, uk, em, pm) \
    if(PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm) && smem) { \
        return launch_one<SceneShape<(a), (b), (cc), (d), (bx), (by), (bz), (uk), (em), (pm)>, true, Integ>(p, sm_count, stream); \
    }
    PTB_MEGA_SPECIALISATIONS(X)
#undef X
    if(smem) {
        return launch_one<GenericShape, true, Integ>(p, sm_count, stream);
    }
    return launch_one<GenericShape, false, Integ>(p, sm_count, stream);
}

cudaError_t launch_megakernel(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream,
                              int* launches, bool smallpt)
{
    cudaError_t e = cudaMemsetAsync(&p.counters->tile_cursor, 0, sizeof(unsigned long long), stream);
    if(e != cudaSuccess) {
        return e;
    }
    if(launches != nullptr) {
        *launches += 1;
    }
    return smallpt ? launch_mega_integrator<IntegratorSmallpt>(p, c, sm_count, stream)
                   : launch_mega_integrator<IntegratorPt>(p, c, sm_count, stream);
}

} // namespace ptb

#include "ptb_mega_sorted.cuh"
#include "ptb_wavefront.cuh"

namespace ptb {

// ---- FP32 probe ---------------------------------------------------------------------------------
// One thread per requested sample; same device functions AND the same (n_small, n_big)
// specialisation as the megakernel, so the probe traces what the renderer traces.
template<class Shape, class Integ>
__global__ void __launch_bounds__(kMegaThreads) probe_f32_kernel(ProbeParams const q, ShadePlanes const sp, GeoLists const geo)
{
    __shared__ float s_split[Integ::kSplit ? 2 * kSplitFields * kMegaThreads : 1];
    SplitStack split{ s_split, kMegaThreads, 0 };
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= q.count) {
        return;
    }
    uint32_t const x = q.x[i], y = q.y[i], sx = q.sx[i], sy = q.sy[i];
    uint32_t const slot = ((y * q.width + x) * q.ns + sy) * q.ns + sx;
    PathF32 p;
    p.rng = rng_open(q.key, slot, q.sample[i]);
    Integ::generate(p, q.cams, x, y, sx, sy);
    if(q.ray != nullptr) {
        q.ray[6 * i + 0] = p.ox;
        q.ray[6 * i + 1] = p.oy;
        q.ray[6 * i + 2] = p.oz;
        q.ray[6 * i + 3] = p.dx * p.len; // report the reference's un-normalised direction
        q.ray[6 * i + 4] = p.dy * p.len;
        q.ray[6 * i + 5] = p.dz * p.len;
    }
    {
        RayTerms const r0 = ray_terms(p, Shape::uniform_k ? c_scene.big_geo[0].k : 0.0f);
        float t0;
        int pos;
        bool const hit0 = closest_hit<Shape>(c_scene, geo, p, r0, t0, pos);
        q.primary_hit[i] = hit0 ? geo.order[pos] : -1; // list position -> the caller's sphere index
    }

    BounceCounters cnt{ 0, 0, 0, 0 };
    bool alive = true;
#ifdef PTB_DEBUG_NAN
    int bounce_index = 0; // development build (dev/build_variant.sh): report the first bounce that leaves a non-finite state
#endif
    while(alive) {
        RayTerms const r = ray_terms(p, Shape::uniform_k ? c_scene.big_geo[0].k : 0.0f);
        float t;
        int id;
        bool const hit = closest_hit<Shape>(c_scene, geo, p, r, t, id);
#ifdef PTB_DEBUG_NAN
        PathF32 const before = p;
#endif
        alive = Integ::bounce(p, hit, t, id, sp, cnt, split);
#ifdef PTB_DEBUG_NAN
        float const chk = p.ox + p.oy + p.oz + p.dx + p.dy + p.dz + p.tr + p.tg + p.tb + p.er + p.eg + p.eb;
        float const d2 = p.dx * p.dx + p.dy * p.dy + p.dz * p.dz;
        if(!isfinite(chk) || (alive && fabsf(d2 - 1.0f) > 1e-3f)) {
            q.radiance[3 * i + 0] = bounce_index;
            q.radiance[3 * i + 1] = hit ? id : -1;
            q.radiance[3 * i + 2] = t;
            if(q.ray != nullptr) {
                q.ray[6 * i + 0] = before.ox;
                q.ray[6 * i + 1] = before.oy;
                q.ray[6 * i + 2] = before.oz;
                q.ray[6 * i + 3] = before.dx;
                q.ray[6 * i + 4] = before.dy;
                q.ray[6 * i + 5] = before.dz;
                q.primary_hit[i] = __float_as_int(d2); // |d|^2 after the bounce
            }
            if(q.draws != nullptr) {
                q.draws[i] = 1000u + static_cast<uint32_t>(before.last + 1);
            }
            return;
        }
        bounce_index++;
#endif
    }
    q.radiance[3 * i + 0] = p.er;
    q.radiance[3 * i + 1] = p.eg;
    q.radiance[3 * i + 2] = p.eb;
    if(q.draws != nullptr) {
        q.draws[i] = 0;
    }
}

template<class Integ>
static cudaError_t launch_probe_integrator(ProbeParams const& p, SceneCounts const& c, ShadePlanes const& shade,
                                           GeoLists const& geo, cudaStream_t stream)
{
    unsigned const threads = kMegaThreads;
    unsigned const blocks = (p.count + threads - 1) / threads;
#define X(a, b, cc, d, bx, by, bz, uk, em, pm) \
    if(PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm)) { \
        probe_f32_kernel<SceneShape<(a), (b), (cc), (d), (bx), (by), (bz), (uk), (em), (pm)>, Integ><<<blocks, threads, 0, stream>>>(p, shade, geo); \
        return cudaGetLastError(); \
    }
    PTB_MEGA_SPECIALISATIONS(X)
#undef X
    probe_f32_kernel<GenericShape, Integ><<<blocks, threads, 0, stream>>>(p, shade, geo);
    return cudaGetLastError();
}

cudaError_t launch_probe_f32(ProbeParams const& p, SceneCounts const& c, ShadePlanes const& shade, GeoLists const& geo,
                             cudaStream_t stream, bool smallpt)
{
    if(p.count == 0) {
        return cudaSuccess;
    }
    return smallpt ? launch_probe_integrator<IntegratorSmallpt>(p, c, shade, geo, stream)
                   : launch_probe_integrator<IntegratorPt>(p, c, shade, geo, stream);
}

// ---- raw draws of the stream (device side), for tests/test_rng.py ----------------------------------
__global__ void rng_draws_kernel(uint64_t key, uint32_t const* slot, uint32_t const* sample, uint32_t count, int n_draws,
                                 double* out)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= count) {
        return;
    }
    Rng g = rng_open(key, slot[i], sample[i]);
    for(int k = 0; k < n_draws; ++k) {
        // odd draws go through the FP32 path, even through the FP64 path: both must
        // give the same 23-bit value
        out[static_cast<size_t>(i) * n_draws + k] = (k & 1) ? static_cast<double>(rng_uniform_f32(g)) : rng_uniform_f64(g);
    }
}

cudaError_t launch_rng_draws(uint64_t key, uint32_t const* slot, uint32_t const* sample, uint32_t count, int n_draws,
                             double* out, cudaStream_t stream)
{
    if(count == 0) {
        return cudaSuccess;
    }
    unsigned const threads = 128;
    rng_draws_kernel<<<(count + threads - 1) / threads, threads, 0, stream>>>(key, slot, sample, count, n_draws, out);
    return cudaGetLastError();
}

} // namespace ptb
