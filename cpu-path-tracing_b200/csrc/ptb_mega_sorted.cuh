// ptb_mega_sorted.cuh -- the material-sorted megakernel (included by ptb_f32.cu).
//
// Same job as mega_kernel (ptb_f32.cu): the row tasks of /root/reference/src/main.cpp:217-236,
// one persistent launch, warps pulling tiles from a global cursor.  What changes is WHERE the
// three-way material branch of main.cpp:141-154 runs.
//
// Measured on mega_kernel (profiles/r1_mega_v3_box_mirror_ncu.md): the closest-hit scan issues
// at 31.8 of 32 lanes, but everything after it diverges -- diffuse_ray (58 instructions) runs
// on 63 % of the loop iterations at 2.2 lanes, dielectric_ray at 3.3 lanes, together ~22 % of
// all issue slots for ~12 % of the bounces.  Here a warp keeps two rings of complete path
// states in shared memory:
//
//   READY  rays waiting for a closest-hit query (fresh camera samples and scattered paths)
//   PARK   paths whose hit needs work that only a few lanes would do
//
// Per loop iteration every lane holds one ray in registers: scan (all lanes), the part of the
// bounce that is the same for every hit (hit point, emission, Russian roulette, throughput), then
//   - a hit on the scene's DOMINANT material (template parameter kInline: mirror in box_mirror --
//     80 % of the bounces -- diffuse in the plain box; chosen by the host from the previous
//     launch's counters) scatters in place,
//   - any other surviving hit -- another material, or the last allowed depth -- is PARKED
//     (four 16-byte shared stores),
//   - lanes whose path ended or was parked take the next READY ray (four 16-byte loads),
//   - when 32 paths are parked the whole warp scatters them at once -- one entry per lane, all
//     lanes busy: diffuse_ray / specular_ray / dielectric_ray -- and appends the survivors to READY.
// Camera samples are generated 32 at a time into READY as before.  The random stream is keyed
// by (seed, slot, sample) and travels with the path, so the image does not depend on any of this.
//
// Emission leaves the path state: it is added to the slot where it is picked up
// (red.global.add.v4.f32 with weight 0), and the sample COUNT of a slot is added once per tile when
// the tile is fetched, so a path's end is silent unless it sees the sky -- a parked
// path is 14 words, four planes.
//
// Ring accounting (all warp-uniform): tokens = lanes + READY + PARK.  Camera samples are only
// generated when READY + PARK <= 32, so tokens <= 96.  At the top of the loop either every lane
// holds a ray -- then PARK >= 32 implies READY <= 32, room for the scatter stage's 32 results --
// or READY is empty.  The scatter stage leaves PARK < 32, so the bounce may park all 32 lanes.
#pragma once

namespace ptb {

constexpr int kSortedThreads = 128;
#ifdef PTB_SORTED_BLOCKS
constexpr int kSortedBlocksPerSm = PTB_SORTED_BLOCKS; // dev/build_variant.sh experiments
#else
constexpr int kSortedBlocksPerSm = 6;
#endif
constexpr int kSortedPool = 64;     // entries per pool
constexpr int kEmissiveBit = 0x100; // in ShadePlanes::b.w next to the reflection tag (ptb_api.cpp: pack_geometry)
constexpr uint32_t kNoSphere = 0x00FFFFFFu; // low 24 bits of the depth | sphere word of a camera ray (sign-extends to -1)

// Byte layout of one warp's pools: two stacks x four planes of 16-byte cells [kSortedPool]
//   plane a = origin xyz, len           plane b = direction xyz, slot
//   plane c = throughput rgb, depth << 24 | sphere the ray starts on        plane d = rng.state, rng.inc (8 of the 16 bytes:
//   one address register + immediate plane offsets serve all four accesses of an entry)
constexpr uint32_t kPlaneBytes = kSortedPool * 16u;
constexpr uint32_t kStackBytes = 4u * kPlaneBytes;
constexpr uint32_t kPoolBytes = 2u * kStackBytes; // READY at +0, PARK at +kStackBytes

// Explicit shared-window accesses: a 32-bit address register + immediate plane offset per access
// (through generic pointers the compiler re-derives the window base in every loop iteration).
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" : : "r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" : : "r"(addr), "r"(x), "r"(y) : "memory");
}

template<class Shape, bool kSmemShade, int kInline>
__global__ void __launch_bounds__(kSortedThreads, kSortedBlocksPerSm) mega_sorted_kernel(RenderParamsF32 const prm)
{
    __shared__ __align__(16) float4 s_shade[kSmemShade ? 4 * kSmemShadeSpheres : 1];
    __shared__ __align__(16) unsigned char s_pool[(kSortedThreads / 32) * kPoolBytes];
    if constexpr(kSmemShade) {
        for(int i = threadIdx.x; i < prm.n_total; i += kSortedThreads) {
            s_shade[i] = prm.shade.a[i];
            s_shade[kSmemShadeSpheres + i] = prm.shade.b[i];
            s_shade[2 * kSmemShadeSpheres + i] = prm.shade.c[i];
            s_shade[3 * kSmemShadeSpheres + i] = prm.shade.d[i];
        }
        __syncthreads();
    }
    // shading plane k of sphere i: shared window byte address shade_base + (k * stride + i) * 16, or the global array
    // (through a shuffle: a value the compiler can re-derive from special registers gets re-derived in the loop --
    //  S2R + LEA at the head of the shading chain -- instead of living in a register)
    uint32_t const shade_base = __shfl_sync(0xffffffffu, static_cast<uint32_t>(__cvta_generic_to_shared(s_shade)), 0);
    constexpr uint32_t kShadePlane = kSmemShadeSpheres * 16u;
    float4 const* const gshade = prm.shade.a; // the global planes are contiguous (ptb_api.cpp: shade_planes)
    int const gstride = prm.n_total;
    auto const shade = [&](int plane, int id) -> float4 {
        if constexpr(kSmemShade) {
            return lds128(shade_base + static_cast<uint32_t>(plane) * kShadePlane + static_cast<uint32_t>(id) * 16u);
        }
        else {
            return gshade[plane * gstride + id];
        }
    };

    // colour plane of a bounce: plane 3 (colour / p, p) under Russian roulette, plane 2 (colour, p) before -- two predicated
    // loads into the same registers instead of selecting an address
    auto const shade_pair = [&](int id, bool second) -> float4 {
        if constexpr(kSmemShade) {
            float4 v;
            uint32_t const at = shade_base + static_cast<uint32_t>(id) * 16u;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                         "@q ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%6];\n\t"
                         "@!q ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%7];\n\t}"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(at), "r"(static_cast<uint32_t>(second)), "n"(3u * kShadePlane), "n"(2u * kShadePlane));
            return v;
        }
        else {
            return gshade[(second ? 3 : 2) * gstride + id];
        }
    };
    // Scenes enclosed by walls (three or more huge spheres) have long paths: most bounces happen past depth 4, nearly every
    // warp has a lane that needs the roulette draw, and the branch around it (BSSY / BRA / BSYNC) costs more than letting
    // the other lanes compute it too: -3.3 % time on box_mirror, -2.7 % on box, but +1.5 % on the two-bounce dof_glass.
    constexpr bool kFlatRoulette = !Shape::generic && Shape::big_near + Shape::big_both >= 3;
    constexpr uint32_t kFull = 0xffffffffu;
    uint32_t const lane = threadIdx.x & 31u;
    uint32_t const lane_bit = 1u << lane;
    uint32_t const lt_mask = lane_bit - 1u;
    uint32_t const ready_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_pool)) + (threadIdx.x >> 5) * kPoolBytes;
    uint32_t const park_base = ready_base + kStackBytes;
#ifdef PTB_JIT_SCENE_INIT
    // Run-time compiled build (ptb_jit.cpp): the scene's coefficients arrive as LITERALS and fold into immediates of
    // the FFMA / FADD instructions -- no uniform loads in the scan, no register-file bandwidth for them either.
    JitSceneT<Shape> const scene = PTB_JIT_SCENE_INIT;
#else
    ConstSceneF32 const& scene = c_scene;
#endif
    float const k_uniform = Shape::uniform_k ? scene.big_geo[0].k : 0.0f;
    uint32_t const keep_reg = prm.key_mask; // see closest_hit: a run-time value so that it stays in a register

    // warp-uniform state.  Both pools are STACKS: one byte offset each (16 x the number of entries), no wrap-around and
    // no tail pointer.  Which ray a lane takes next does not matter -- the stream travels with the path -- and what is
    // left at the bottom of READY is traced when the tiles run out.
    uint32_t tile_sample0 = 0, tile_samples = 0, next_sample = 0;
    uint32_t ready_top = 0, park_top = 0;
    int refill_level = 32 * 16; // camera samples are generated while READY + PARK hold <= 32 entries; -1 once the tiles are gone
    bool exhausted = false;
    // per lane: the sub-pixel this lane generates camera samples for in the current tile
    uint32_t gen_slot = 0, gen_valid = 0, gen_x = 0, gen_y = 0, gen_sx = 0, gen_sy = 0;

    // Which lanes hold a ray is kept as a warp-uniform mask: "who needs a ray" costs no vote, and the
    // common case -- every lane busy -- is one compare.
    uint32_t am = 0u;
    uint32_t slot = 0;
    // depth and the sphere the ray starts on share one word, depth << 24 | sphere (kNoSphere for a camera ray): the path
    // state is 14 words -- three 16-byte planes and one 8-byte plane -- and the bounce never unpacks it: "depth > 4" and
    // "depth < 99" are unsigned compares of the whole word, ++depth adds 1 << 24, a hit replaces the low 24 bits
    uint32_t dl = kNoSphere;
    PathF32 p;
    p.er = p.eg = p.eb = 0.0f; // unused here: emission is flushed where it is picked up
    p.ox = p.oy = p.oz = p.dx = p.dy = p.dz = p.len = p.tr = p.tg = p.tb = 0.0f;
    p.rng.state = p.rng.inc = 0u;
    p.depth = 0;
    p.last = -1;
    BounceCounters cnt{ 0, 0, 0, 0 };

    // lanes of `need` take the top READY entries, lowest lane first; returns the new "holds a ray" mask
    auto const pop = [&](uint32_t need) -> uint32_t {
        uint32_t const rank16 = __popc(need & lt_mask) * 16u;
        bool const take = (need & lane_bit) != 0u && rank16 < ready_top;
        if(take) {
            uint32_t const at = ready_top - 16u - rank16;
            uint32_t const rd = ready_base + at;
            float4 const ea = lds128(rd);
            float4 const eb = lds128(rd + kPlaneBytes);
            float4 const ec = lds128(rd + 2u * kPlaneBytes);
            uint2 const ed = lds64(rd + 3u * kPlaneBytes);
            p.ox = ea.x;
            p.oy = ea.y;
            p.oz = ea.z;
            p.len = ea.w;
            p.dx = eb.x;
            p.dy = eb.y;
            p.dz = eb.z;
            slot = __float_as_uint(eb.w);
            p.tr = ec.x;
            p.tg = ec.y;
            p.tb = ec.z;
            dl = __float_as_uint(ec.w);
            p.rng.state = ed.x;
            p.rng.inc = ed.y;
        }
        uint32_t const wanted = static_cast<uint32_t>(__popc(need)) * 16u;
        uint32_t const taken = min(wanted, ready_top);
        ready_top -= taken;
        return taken == wanted ? kFull : (~need | __ballot_sync(kFull, take));
    };
    // one complete path state into a pool at byte offset `at` (16 x entry index)
    auto const push = [&](uint32_t base, uint32_t at, PathF32 const& q, uint32_t q_slot, float tr, float tg, float tb, uint32_t q_dl) {
        sts128(base + at, q.ox, q.oy, q.oz, q.len);
        sts128(base + at + kPlaneBytes, q.dx, q.dy, q.dz, __uint_as_float(q_slot));
        sts128(base + at + 2u * kPlaneBytes, tr, tg, tb, __uint_as_float(q_dl));
        sts64(base + at + 3u * kPlaneBytes, q.rng.state, q.rng.inc);
    };

    for(;;) {
        // ================= slow path: pool maintenance, a few percent of the iterations ===========================
        bool const park_full = park_top >= 32u * 16u;
        if(am != kFull || park_full || static_cast<int>(ready_top + park_top) <= refill_level) {
            // ---- scatter stage: up to 32 parked paths, one per lane -------------------------------------------
            if(park_full || (am != kFull && exhausted && ready_top == 0u && park_top != 0u)) {
                __syncwarp(); // the parked entries were written by other lanes
                uint32_t const k16 = min(32u * 16u, park_top);
                bool out = false;
                PathF32 q;
                q.ox = q.oy = q.oz = q.dx = q.dy = q.dz = q.len = 0.0f;
                q.rng.state = q.rng.inc = 0u;
                float4 eb = make_float4(0.0f, 0.0f, 0.0f, 0.0f), ec = eb;
                uint32_t dl1 = 0u;
                if(lane * 16u < k16) {
                    uint32_t const at = park_top - k16 + lane * 16u;
                    uint32_t const rd = park_base + at;
                    float4 const ea = lds128(rd);
                    eb = lds128(rd + kPlaneBytes);
                    ec = lds128(rd + 2u * kPlaneBytes);
                    uint2 const ed = lds64(rd + 3u * kPlaneBytes);
                    q.ox = ea.x;
                    q.oy = ea.y;
                    q.oz = ea.z;
                    q.len = ea.w;
                    q.dx = eb.x;
                    q.dy = eb.y;
                    q.dz = eb.z;
                    q.rng.state = ed.x;
                    q.rng.inc = ed.y;
                    uint32_t const qdl = __float_as_uint(ec.w);
                    int const last = static_cast<int>(qdl & kNoSphere);
                    dl1 = qdl + (1u << 24); // main.cpp:111 ++depth
                    // outward normal at the hit point, as in the bounce that parked the path (hit_record.cpp:6); the
                    // material is the sphere's: read back from its shading plane rather than carried through the pool
                    float4 const sa = shade(0, last);
                    int const mat = __float_as_int(shade(1, last).w) & 0xff;
                    if(mat == 1) {
                        cnt.specular++;
                        mirror_at_hit(q, sa);
                    }
                    else {
                        float nx, ny, nz;
                        unit_normal(q.ox, q.oy, q.oz, sa, nx, ny, nz);
                        float const dn = fmaf(nx, q.dx, fmaf(ny, q.dy, nz * q.dz));
                        if(mat == 0) {
                            cnt.diffuse++;
                            bool const front = dn < 0.0f; // hit_record.cpp:7
                            scatter_diffuse(q, front ? nx : -nx, front ? ny : -ny, front ? nz : -nz);
                        }
                        else {
                            cnt.dielectric++;
                            scatter_dielectric(q, nx, ny, nz, dn);
                        }
                    }
                    // the ray scattered by the last iteration is never traced (main.cpp:111): the path ends here
                    out = dl1 < (static_cast<uint32_t>(kDepthLimit) << 24);
                }
                uint32_t const om = __ballot_sync(kFull, out);
                if(out) {
                    push(ready_base, ready_top + __popc(om & lt_mask) * 16u, q, __float_as_uint(eb.w), ec.x, ec.y, ec.z, dl1);
                }
                park_top -= k16;
                ready_top += static_cast<uint32_t>(__popc(om)) * 16u;
                __syncwarp();
            }

            // ---- refill: one camera sample per lane (main.cpp:186-190, camera.cpp:19-38) -----------------
            if(static_cast<int>(ready_top + park_top) <= refill_level) {
                __syncwarp(); // entries read by earlier pops are about to be overwritten
                if(next_sample >= tile_samples) {
                    unsigned long long t = 0;
                    if(lane == 0) {
                        t = atomicAdd(&prm.counters->tile_cursor, 1ull);
                    }
                    t = __shfl_sync(kFull, t, 0);
                    if(t >= prm.ntiles) {
                        exhausted = true;
                        refill_level = -1;
                    }
                    else {
                        uint32_t const tile = static_cast<uint32_t>(t);
                        uint32_t const group = tile / prm.nchunks;
                        uint32_t const chunk = tile - group * prm.nchunks;
                        tile_sample0 = chunk * prm.chunk;
                        tile_samples = min(prm.chunk, prm.samples - tile_sample0);
                        next_sample = 0;
                        gen_slot = group * 32u + lane;
                        gen_valid = min(32u, prm.nslots - group * 32u); // the slots of a tile are a prefix of its lanes
                        if(lane < gen_valid) {
                            slot_coords(gen_slot, prm.width, prm.ns, gen_x, gen_y, gen_sx, gen_sy);
                            // the slot's sample count (main.cpp:192 divides by it) is credited HERE, once per tile, for all the
                            // samples this lane will start: a path that ends by roulette or at the depth limit then ends
                            // silently -- no divergent "weight 1" add at 3 of 32 lanes in nine of ten iterations
                            red_add_v4(prm.accum + gen_slot, 0.0f, 0.0f, 0.0f, static_cast<float>(tile_samples));
                        }
                    }
                }
                if(!exhausted) {
                    if(lane < gen_valid) {
                        PathF32 g;
                        g.rng = rng_open(prm.key, gen_slot, prm.first_sample + tile_sample0 + next_sample);
                        gen_primary(g, prm.cams.cam, gen_x, gen_y, gen_sx, gen_sy);
                        push(ready_base, ready_top + lane * 16u, g, gen_slot, 1.0f, 1.0f, 1.0f, kNoSphere);
                    }
                    ready_top += gen_valid * 16u;
                    next_sample += 1u;
                    __syncwarp();
                }
            }

            // ---- lanes that found READY empty at the end of their last bounce try again ---------------------
            if(am != kFull && ready_top != 0u) {
                am = pop(~am);
            }
            if(am == 0u) {
                if(exhausted && ready_top == 0u && park_top == 0u) {
                    break;
                }
                continue;
            }
        }

        // ================= one bounce: main.cpp:111-155 =========================================================
        // what the lane needs from the turnover: 0 = keeps its ray, 1 = parks it and takes a new one, 2 = takes a new one
        // (no ray, or its path ended).  One value instead of two flags: three predicate moves fewer per bounce (-1.8 % time).
        int state = 2;
        bool ended = false;
        bool const alive = (am & lane_bit) != 0u;
        if(alive) {
            if constexpr(Shape::generic || !Shape::embed) {
                p.last = static_cast<int>(dl << 8) >> 8; // full-precision keys: the self-sphere root needs the list position
            }
            RayTerms const r = ray_terms(p, k_uniform);
            float t;
            int id;
            bool const hit = closest_hit<Shape, true>(scene, prm.geo, p, r, t, id, keep_reg);
            cnt.rays++;
            if(!hit) {
                // main.cpp:116-119 sky gradient on the unit direction; the path ends
                float const tt = 0.5f * (p.dy + 1.0f);
                float const omt = 1.0f - tt;
                red_add_v4(prm.accum + slot, p.tr * fmaf(0.5f, tt, omt), p.tg * fmaf(0.7f, tt, omt), p.tb * (omt + tt), 0.0f);
                ended = true;
            }
            else {
                // hit_record.cpp:5-6
                p.ox = fmaf(p.dx, t, p.ox);
                p.oy = fmaf(p.dy, t, p.oy);
                p.oz = fmaf(p.dz, t, p.oz);
                // main.cpp:128-139 Russian roulette once depth > 4: p = max(color), survivor weight color / p
                bool const roulette = dl >= (static_cast<uint32_t>(kRouletteThreshold + 1) << 24);
                dl = (dl & ~kNoSphere) | static_cast<uint32_t>(id);
                float4 const sb = shade(1, id);
                int const tag = __float_as_int(sb.w);
                if((tag & kEmissiveBit) != 0) { // main.cpp:126
                    red_add_v4(prm.accum + slot, p.tr * sb.x, p.tg * sb.y, p.tb * sb.z, 0.0f);
                }
                float4 const col = shade_pair(id, roulette);
                if constexpr(kFlatRoulette) {
                    // the draw without a branch: every lane computes a uniform from a COPY of its stream, only lanes past
                    // depth 4 keep the advanced state
                    Rng g2 = p.rng;
                    float const u = rng_uniform_f32(g2);
                    p.rng.state = roulette ? g2.state : p.rng.state;
                    ended = roulette && !(u < col.w);
                }
                else if(roulette) {
                    ended = !(rng_uniform_f32(p.rng) < col.w);
                }
                p.tr *= col.x;
                p.tg *= col.y;
                p.tb *= col.z;
                if(ended) {
                    // main.cpp:131-133: the path dies; its sample was counted when its tile was fetched
                }
                else if(((tag ^ kInline) & 0xff) == 0 && dl < (static_cast<uint32_t>(kDepthLimit - 1) << 24)) {
                    float4 const sa = shade(0, id);
                    if constexpr(kInline == 1) {
                        cnt.specular++;
                        mirror_at_hit(p, sa); // specular_ray, main.cpp:60-67
                    }
                    else {
                        float nx, ny, nz; // outward unit normal, hit_record.cpp:6
                        unit_normal(p.ox, p.oy, p.oz, sa, nx, ny, nz);
                        cnt.diffuse++;
                        bool const front = fmaf(nx, p.dx, fmaf(ny, p.dy, nz * p.dz)) < 0.0f; // hit_record.cpp:7
                        scatter_diffuse(p, front ? nx : -nx, front ? ny : -ny, front ? nz : -nz);
                    }
                    dl += 1u << 24;
                    state = 0;
                }
                else {
                    state = 1; // another material, or the last allowed depth (main.cpp:111)
                }
            }
        }

        // ================= turnover: park, then hand a READY ray to every lane without one ========================
        bool const park = state == 1;
        uint32_t const pm = __ballot_sync(kFull, park);
        uint32_t const need = __ballot_sync(kFull, state != 0);
        if(need != 0u) {
            if(park) {
                push(park_base, park_top + __popc(pm & lt_mask) * 16u, p, slot, p.tr, p.tg, p.tb, dl);
            }
            park_top += static_cast<uint32_t>(__popc(pm)) * 16u;
            am = pop(need);
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

#ifndef __CUDACC_RTC__ // host side: launchers of the precompiled instantiations

template<class Shape, bool kSmem, int kInline>
static cudaError_t launch_sorted_one(RenderParamsF32 const& p, int sm_count, cudaStream_t stream)
{
    int per_sm = 0;
    cudaError_t e =
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_sorted_kernel<Shape, kSmem, kInline>, kSortedThreads, 0);
    if(e != cudaSuccess) {
        return e;
    }
    per_sm = per_sm < 1 ? 1 : per_sm;
    if(char const* cap = std::getenv("PTB_BLOCKS_PER_SM")) { // occupancy experiments (dev/)
        int const v = std::atoi(cap);
        per_sm = v >= 1 && v < per_sm ? v : per_sm;
    }
    unsigned long long blocks = static_cast<unsigned long long>(sm_count) * static_cast<unsigned long long>(per_sm);
    unsigned long long const blocks_needed =
        (static_cast<unsigned long long>(p.ntiles) + (kSortedThreads / 32) - 1) / (kSortedThreads / 32);
    blocks = blocks > blocks_needed ? blocks_needed : blocks;
    blocks = blocks < 1 ? 1 : blocks;
    mega_sorted_kernel<Shape, kSmem, kInline><<<static_cast<unsigned>(blocks), kSortedThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

template<int kInline>
static cudaError_t launch_sorted_inline(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream)
{
    bool const smem = p.n_total <= kSmemShadeSpheres;
#define X(a, b, cc, d, bx, by, bz, uk, em, pm) \
    if(PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm) && smem) { \
        return launch_sorted_one<SceneShape<(a), (b), (cc), (d), (bx), (by), (bz), (uk), (em), (pm)>, true, kInline>(p, sm_count, stream); \
    }
    PTB_MEGA_SPECIALISATIONS(X)
#undef X
    if(smem) {
        return launch_sorted_one<GenericShape, true, kInline>(p, sm_count, stream);
    }
    return launch_sorted_one<GenericShape, false, kInline>(p, sm_count, stream);
}

// inline_material: 0 = diffuse_ray runs in place, 1 = specular_ray runs in place; the other two are sorted
// through the PARK ring.  Any value gives the same image; the right one is the scene's commonest material.
cudaError_t launch_megakernel_sorted(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream,
                                     int* launches, int inline_material)
{
    cudaError_t e = cudaMemsetAsync(&p.counters->tile_cursor, 0, sizeof(unsigned long long), stream);
    if(e != cudaSuccess) {
        return e;
    }
    if(launches != nullptr) {
        *launches += 1;
    }
    if(inline_material < 0) {
        // no material in place: every surviving hit is parked, so after a bounce ALL lanes take new rays and take them from
        // the top of READY -- a batch the scatter stage or the camera just pushed, i.e. rays of one kind.  Precompiled for the
        // run-time-count kernel only (scenes behind the hierarchy, where lanes of one kind descend alike)
        return p.n_total <= kSmemShadeSpheres ? launch_sorted_one<GenericShape, true, -1>(p, sm_count, stream)
                                              : launch_sorted_one<GenericShape, false, -1>(p, sm_count, stream);
    }
    return inline_material == 0 ? launch_sorted_inline<0>(p, c, sm_count, stream) : launch_sorted_inline<1>(p, c, sm_count, stream);
}

#endif // __CUDACC_RTC__

} // namespace ptb
