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

constexpr int kMegaThreads = 128;

cudaError_t upload_const_scene(ConstSceneF32 const& cs, cudaStream_t stream)
{
    return cudaMemcpyToSymbolAsync(c_scene, &cs, sizeof(cs), 0, cudaMemcpyHostToDevice, stream);
}

// Per-warp ring of pre-generated camera samples (shared memory).  Generating a primary ray
// costs ~130 instructions; done in place by the one or two lanes whose path just ended it
// would issue at <10 % lane utilisation on almost every iteration (measured: 30 % of all
// issue slots, profiles/r1_mega_v1_*).  Instead ALL 32 lanes generate one sample each when
// the ring runs low (full lane utilisation, once per ~12 iterations) and a lane whose
// path ended just pops a ready ray: two 16-byte shared loads.
constexpr int kRingSize = 64; // entries per warp, power of two, >= 2 * 32
struct WarpRing
{
    float4 a[kRingSize];      // ox, oy, dx, dy        (oz is the camera's z: the lens offset has no z)
    float4 b[kRingSize];      // dz, len, rng.state, rng.inc
    uint32_t slot[kRingSize]; // sub-pixel slot, kVoidSlot = nothing to trace
};
constexpr uint32_t kVoidSlot = 0xFFFFFFFFu;

// ---- integrator policies: what differs between the two programs of the reference ---------------------------
// src/main.cpp: thin-lens camera, iterative radiance (ptb_path_f32.cuh).  Ring word b.y carries `len`
// (the origin's z is the camera's: the lens offset has no z component).
struct IntegratorPt
{
    static constexpr bool kSplit = false;
    __device__ static __forceinline__ void generate(PathF32& g, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy)
    {
        gen_primary(g, c_scene.cam, x, y, sx, sy);
    }
    __device__ static __forceinline__ float ring_word(PathF32 const& g)
    {
        return g.len;
    }
    __device__ static __forceinline__ void unpack(PathF32& p, float word)
    {
        p.len = word;
        p.oz = c_scene.cam.pz;
    }
    __device__ static __forceinline__ bool bounce(PathF32& p, bool hit, float t, int id, ShadePlanes const& sp,
                                                  BounceCounters& cnt, SplitStack&)
    {
        return shade_bounce<true>(p, hit, t, id, sp, cnt);
    }
};
// sandbox/main.cpp (stand-alone smallpt): pinhole tent-filter camera, splitting glass (ptb_smallpt_f32.cuh).
// Directions are always unit (len == 1), so ring word b.y carries the origin's z instead.
struct IntegratorSmallpt
{
    static constexpr bool kSplit = true;
    __device__ static __forceinline__ void generate(PathF32& g, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy)
    {
        gen_smallpt(g, c_scene.sbcam, x, y, sx, sy);
    }
    __device__ static __forceinline__ float ring_word(PathF32 const& g)
    {
        return g.oz;
    }
    __device__ static __forceinline__ void unpack(PathF32& p, float word)
    {
        p.len = 1.0f;
        p.oz = word;
    }
    __device__ static __forceinline__ bool bounce(PathF32& p, bool hit, float t, int id, ShadePlanes const& sp,
                                                  BounceCounters& cnt, SplitStack& st)
    {
        return bounce_smallpt<true>(p, hit, t, id, sp, cnt, st);
    }
};

template<class Shape, bool kSmemShade, class Integ>
__global__ void __launch_bounds__(kMegaThreads) mega_kernel(RenderParamsF32 const prm)
{
    __shared__ float4 s_shade[kSmemShade ? 4 * kSmemShadeSpheres : 1];
    __shared__ WarpRing s_ring[kMegaThreads / 32];
    __shared__ float s_split[Integ::kSplit ? 2 * kSplitFields * kMegaThreads : 1];
    SplitStack split{ s_split, kMegaThreads, 0 };
    ShadePlanes sp = prm.shade;
    if constexpr(kSmemShade) {
        for(int i = threadIdx.x; i < prm.n_total; i += kMegaThreads) {
            s_shade[i] = prm.shade.a[i];
            s_shade[kSmemShadeSpheres + i] = prm.shade.b[i];
            s_shade[2 * kSmemShadeSpheres + i] = prm.shade.c[i];
            s_shade[3 * kSmemShadeSpheres + i] = prm.shade.d[i];
        }
        __syncthreads();
        sp.a = s_shade;
        sp.b = s_shade + kSmemShadeSpheres;
        sp.c = s_shade + 2 * kSmemShadeSpheres;
        sp.d = s_shade + 3 * kSmemShadeSpheres;
    }

    uint32_t const lane = threadIdx.x & 31u;
    uint32_t const lt_mask = (1u << lane) - 1u;
    WarpRing& ring = s_ring[threadIdx.x >> 5];

    // warp-uniform state: current work tile and ring occupancy
    uint32_t tile_sample0 = 0, tile_samples = 0, next_sample = 0;
    uint32_t ring_head = 0, ring_tail = 0, ring_count = 0;
    bool exhausted = false;
    // per-lane: the sub-pixel this lane GENERATES for in the current tile
    uint32_t gen_slot = kVoidSlot, gen_x = 0, gen_y = 0, gen_sx = 0, gen_sy = 0;

    bool alive = false;
    uint32_t slot = 0;
    PathF32 p;
    BounceCounters cnt{ 0, 0, 0, 0 };

    for(;;) {
        // ---- refill: every lane generates one camera sample --------------------------------
        if(ring_count <= kRingSize - 32 && !exhausted) {
            if(next_sample >= tile_samples) {
                unsigned long long t = 0;
                if(lane == 0) {
                    t = atomicAdd(&prm.counters->tile_cursor, 1ull);
                }
                t = __shfl_sync(0xffffffffu, t, 0);
                if(t >= prm.ntiles) {
                    exhausted = true;
                }
                else {
                    uint32_t const tile = static_cast<uint32_t>(t);
                    uint32_t const group = tile / prm.nchunks;
                    uint32_t const chunk = tile - group * prm.nchunks;
                    tile_sample0 = chunk * prm.chunk;
                    tile_samples = min(prm.chunk, prm.samples - tile_sample0);
                    next_sample = 0;
                    gen_slot = group * 32u + lane;
                    if(gen_slot < prm.nslots) {
                        slot_coords(gen_slot, prm.width, prm.ns, gen_x, gen_y, gen_sx, gen_sy);
                    }
                    else {
                        gen_slot = kVoidSlot;
                    }
                }
            }
            if(!exhausted) {
                PathF32 g;
                g.dx = g.dy = g.dz = g.ox = g.oy = g.oz = g.len = 0.0f;
                g.rng.state = g.rng.inc = 0u;
                if(gen_slot != kVoidSlot) {
                    g.rng = rng_open(prm.key, gen_slot, prm.first_sample + tile_sample0 + next_sample);
                    Integ::generate(g, gen_x, gen_y, gen_sx, gen_sy);
                }
                uint32_t const w = (ring_head + lane) & (kRingSize - 1);
                ring.a[w] = make_float4(g.ox, g.oy, g.dx, g.dy);
                ring.b[w] = make_float4(g.dz, Integ::ring_word(g), __uint_as_float(g.rng.state), __uint_as_float(g.rng.inc));
                ring.slot[w] = gen_slot;
                ring_head = (ring_head + 32u) & (kRingSize - 1);
                ring_count += 32u;
                next_sample += 1u;
                __syncwarp();
            }
        }

        // ---- lanes whose path ended pop a ready sample ----------------------------------------
        uint32_t const need = __ballot_sync(0xffffffffu, !alive);
        if(need != 0u && ring_count != 0u) {
            uint32_t const rank = __popc(need & lt_mask);
            if(!alive && rank < ring_count) {
                uint32_t const rd = (ring_tail + rank) & (kRingSize - 1);
                float4 const ea = ring.a[rd];
                float4 const eb = ring.b[rd];
                slot = ring.slot[rd];
                if(slot != kVoidSlot) {
                    p.ox = ea.x;
                    p.oy = ea.y;
                    p.dx = ea.z;
                    p.dy = ea.w;
                    p.dz = eb.x;
                    Integ::unpack(p, eb.y);
                    p.rng.state = __float_as_uint(eb.z);
                    p.rng.inc = __float_as_uint(eb.w);
                    p.tr = p.tg = p.tb = 1.0f;
                    p.er = p.eg = p.eb = 0.0f;
                    p.depth = 0;
                    p.last = -1;
                    alive = true;
                }
            }
            uint32_t const taken = min(static_cast<uint32_t>(__popc(need)), ring_count);
            ring_tail = (ring_tail + taken) & (kRingSize - 1);
            ring_count -= taken;
            __syncwarp();
        }

        if(!__any_sync(0xffffffffu, alive)) {
            if(exhausted && ring_count == 0u) {
                break;
            }
            continue;
        }

        // ---- one bounce ---------------------------------------------------------------------------
        if(alive) {
            RayTerms const r = ray_terms(p, Shape::uniform_k ? c_scene.big_geo[0].k : 0.0f);
            float t;
            int id;
            bool const hit = closest_hit<Shape>(c_scene, prm.geo, p, r, t, id);
            cnt.rays++;
            alive = Integ::bounce(p, hit, t, id, sp, cnt, split);
            if(!alive) {
                red_add_v4(prm.accum + slot, p.er, p.eg, p.eb, 1.0f);
            }
        }
    }

    uint32_t const rays = warp_sum(cnt.rays);
    uint32_t const nd = warp_sum(cnt.diffuse);
    uint32_t const nsp = warp_sum(cnt.specular);
    uint32_t const ndi = warp_sum(cnt.dielectric);
    if(lane == 0) {
        atomicAdd(&prm.counters->rays, static_cast<unsigned long long>(rays));
        atomicAdd(&prm.counters->diffuse, static_cast<unsigned long long>(nd));
        atomicAdd(&prm.counters->specular, static_cast<unsigned long long>(nsp));
        atomicAdd(&prm.counters->dielectric, static_cast<unsigned long long>(ndi));
    }
}

// ---- specialisation table ------------------------------------------------------------------
// (small near-only, small both-roots, big near-only, big both-roots) list lengths with a fully
// unrolled kernel.  Everything else runs the generic run-time-count variant.
#define PTB_MEGA_SPECIALISATIONS(X) \
    X(2, 1, 5, 0, 2, 2, 1, true, true, 3)   /* box_scene.hpp / box_mirror_scene.hpp: light + mirror ball | glass ball | 5 R=1e6 walls on the frame axes, left/right and top/bottom as mirror-image pairs */ \
    X(2, 1, 5, 0, 2, 2, 1, true, true, 0)   /* the same without the pairing                                                                          */ \
    X(3, 1, 1, 0, 0, 1, 0, true, true, 0)   /* simple_scene.hpp: mirror, centre, light | glass | ground (R=100, on the y axis)                       */ \
    X(2, 2, 1, 0, 0, 1, 0, true, true, 0)   /* depth-of-field scene (BASELINE config 4): two glass spheres                                             */ \
    X(2, 1, 5, 0, 0, 0, 0, false, true, 0)  /* box scenes in a frame where the walls are not axis spheres                                             */ \
    X(0, 3, 0, 5, 0, 0, 0, false, true, 0)  /* box scenes with the camera inside every sphere's reach: all both-roots                                 */ \
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
    Integ::generate(p, x, y, sx, sy);
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
