// ptb_mega_inplace.cuh -- the in-place ("path regeneration") megakernel and the two integrator policies
// (included by ptb_f32.cu and, at run time, by the translation unit ptb_jit.cpp hands to NVRTC).
//
// Replaces the row-parallel CPU loop of /root/reference/src/main.cpp:217-236 and, with IntegratorSmallpt, the
// OpenMP loop of /root/reference/sandbox/main.cpp:241-269.  Design notes: ptb_f32.cu.
#pragma once

#include "ptb_kernels.h"
#include "ptb_path_f32.cuh"
#include "ptb_smallpt_f32.cuh"

namespace ptb {

constexpr int kMegaThreads = 128;

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
    __device__ static __forceinline__ void generate(PathF32& g, CameraPair const& cams, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy)
    {
        gen_primary(g, cams.cam, x, y, sx, sy);
    }
    __device__ static __forceinline__ float ring_word(PathF32 const& g)
    {
        return g.len;
    }
    __device__ static __forceinline__ void unpack(PathF32& p, CameraPair const& cams, float word)
    {
        p.len = word;
        p.oz = cams.cam.pz;
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
    __device__ static __forceinline__ void generate(PathF32& g, CameraPair const& cams, uint32_t x, uint32_t y, uint32_t sx, uint32_t sy)
    {
        gen_smallpt(g, cams.sbcam, x, y, sx, sy);
    }
    __device__ static __forceinline__ float ring_word(PathF32 const& g)
    {
        return g.oz;
    }
    __device__ static __forceinline__ void unpack(PathF32& p, CameraPair const&, float word)
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
#ifdef PTB_JIT_SCENE_INIT
    JitSceneT<Shape> const scene = PTB_JIT_SCENE_INIT; // run-time compiled build: coefficients as literals (ptb_jit.cpp)
#else
    ConstSceneF32 const& scene = c_scene;
#endif
    float const k_uniform = Shape::uniform_k ? scene.big_geo[0].k : 0.0f;

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
                    Integ::generate(g, prm.cams, gen_x, gen_y, gen_sx, gen_sy);
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
                    Integ::unpack(p, prm.cams, eb.y);
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
            RayTerms const r = ray_terms(p, k_uniform);
            float t;
            int id;
            bool const hit = closest_hit<Shape>(scene, prm.geo, p, r, t, id);
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

} // namespace ptb
