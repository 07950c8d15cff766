// ptb_wavefront.cuh -- the wavefront / material-sorted variant (included by ptb_f32.cu so
// that it shares the constant-bank scene and the path arithmetic with the megakernel).
//
// Same paths, same random stream, same accumulation as the megakernel; what differs is
// WHERE a path lives between bounces.  The megakernel keeps a path in the registers of one
// lane for its whole life, so the material branches of main.cpp:141-154 and the camera-sample
// hand-over diverge inside a warp.  Here paths live in HBM as dense STREAMS of records (SoA
// float4 planes) and every stage is a kernel that reads one stream front to back and appends
// whole records to its output streams:
//
//   plan + regen : top the active stream up to the pool size with new camera samples
//                  (positions and (sub-pixel, sample) items are assigned arithmetically:
//                  no atomics, no free list)
//   extend       : closest hit + the material-independent half of the bounce (sky, emission,
//                  Russian roulette, throughput).  A mirror hit continues in place -- the
//                  mirror bounce is 7 instructions, not worth a trip through HBM -- and goes
//                  straight to the next active stream; diffuse and dielectric hits are
//                  appended to the stream of THEIR material; finished paths add to the
//                  accumulation buffer and vanish
//   scatter<diffuse|dielectric> : one kernel per material stream -> every warp runs one
//                  scatter function on consecutive records; survivors are appended to the
//                  next active stream
//
// A first version kept the records in a fixed pool and moved 4-byte indices through the
// queues: after a few bounces the index order is random, every float4 is its own 32-byte
// sector and the kernels ran at 0.7 TB/s (extend 722 us per 4 M rays).  Moving the records
// keeps every load and store coalesced.  Appends are block-aggregated: one global atomic per
// block, round and destination.
//
// One iteration = plan, regen, extend, 2 scatters, rotate; an even and an odd iteration are
// captured once in a CUDA graph and replayed; kernels read their work counts from device
// memory, so there is no host round trip per iteration.  Bound: HBM bandwidth (136 B per
// mirror bounce, 304 B per diffuse/dielectric bounce), where the megakernel is FP32-issue bound.
#pragma once

namespace ptb {

constexpr int kWfThreads = 256;

// One stream of path records, SoA.  plane n is only filled for material streams.
struct WfStream
{
    float4* a;      // ox, oy, oz, len
    float4* b;      // dx, dy, dz, depth (int bits)
    float4* c;      // tr, tg, tb, rng.state (bits)
    float4* d;      // er, eg, eb, rng.inc (bits)
    float4* n;      // outward normal at the hit, n.d
    uint32_t* slot; // sub-pixel slot
};

struct WavefrontState
{
    WfStream active[2];
    WfStream mat[2]; // 0 = diffuse, 1 = dielectric
    WavefrontCounters* ctr;
    uint32_t pool;
};

__device__ __forceinline__ void wf_load(WfStream const& s, uint32_t i, PathF32& p, uint32_t& slot)
{
    float4 const a = s.a[i];
    float4 const b = s.b[i];
    float4 const c = s.c[i];
    float4 const d = s.d[i];
    p.ox = a.x;
    p.oy = a.y;
    p.oz = a.z;
    p.len = a.w;
    p.dx = b.x;
    p.dy = b.y;
    p.dz = b.z;
    uint32_t const packed = __float_as_uint(b.w); // depth (< 2^8) | (last + 1) << 8, unsigned: list positions reach 2^24
    p.depth = static_cast<int>(packed & 0xFFu);
    p.last = static_cast<int>(packed >> 8) - 1;
    p.tr = c.x;
    p.tg = c.y;
    p.tb = c.z;
    p.rng.state = __float_as_uint(c.w);
    p.er = d.x;
    p.eg = d.y;
    p.eb = d.z;
    p.rng.inc = __float_as_uint(d.w);
    slot = s.slot[i];
}

__device__ __forceinline__ void wf_store(WfStream const& s, uint32_t i, PathF32 const& p, uint32_t slot)
{
    s.a[i] = make_float4(p.ox, p.oy, p.oz, p.len);
    s.b[i] = make_float4(p.dx, p.dy, p.dz, __uint_as_float(static_cast<uint32_t>(p.depth) | (static_cast<uint32_t>(p.last + 1) << 8)));
    s.c[i] = make_float4(p.tr, p.tg, p.tb, __uint_as_float(p.rng.state));
    s.d[i] = make_float4(p.er, p.eg, p.eb, __uint_as_float(p.rng.inc));
    s.slot[i] = slot;
}

// Block-aggregated stream append: every thread of the block calls it once per round with the
// destination it wants (-1 = none) and gets back the record position in that stream.
template<int NQ>
struct BlockAppend
{
    uint32_t count[NQ];
    uint32_t base[NQ];
};

template<int NQ>
__device__ __forceinline__ uint32_t block_append(BlockAppend<NQ>& ba, int dest, uint32_t* const (&counters)[NQ])
{
    uint32_t const lane = threadIdx.x & 31u;
    if(threadIdx.x < NQ) {
        ba.count[threadIdx.x] = 0;
    }
    __syncthreads();
    uint32_t rank = 0;
#pragma unroll
    for(int q = 0; q < NQ; ++q) {
        uint32_t const mask = __ballot_sync(0xffffffffu, dest == q);
        if(mask != 0u) {
            uint32_t const leader = static_cast<uint32_t>(__ffs(static_cast<int>(mask))) - 1u;
            uint32_t at = 0;
            if(lane == leader) {
                at = atomicAdd(&ba.count[q], static_cast<uint32_t>(__popc(mask)));
            }
            at = __shfl_sync(0xffffffffu, at, static_cast<int>(leader));
            if(dest == q) {
                rank = at + static_cast<uint32_t>(__popc(mask & ((1u << lane) - 1u)));
            }
        }
    }
    __syncthreads();
    if(threadIdx.x < NQ && ba.count[threadIdx.x] != 0u) {
        ba.base[threadIdx.x] = atomicAdd(counters[threadIdx.x], ba.count[threadIdx.x]);
    }
    __syncthreads();
    uint32_t pos = 0;
#pragma unroll
    for(int q = 0; q < NQ; ++q) {
        if(dest == q) {
            pos = ba.base[q] + rank;
        }
    }
    return pos;
}

__global__ void wf_init_kernel(WavefrontCounters* c)
{
    *c = WavefrontCounters{};
}

// decide how many camera samples join the active stream this iteration (one thread)
__global__ void wf_plan_kernel(WavefrontCounters* c, int cur, uint32_t pool, unsigned long long total_items)
{
    uint32_t const have = c->n_active[cur];
    unsigned long long const left = total_items > c->cursor ? total_items - c->cursor : 0ull;
    uint32_t const room = pool > have ? pool - have : 0u;
    uint32_t const n = static_cast<uint32_t>(left < room ? left : room);
    c->regen_base = have;
    c->regen_item0 = c->cursor;
    c->regen_count = n;
    c->n_active[cur] = have + n;
    c->cursor += n;
}

__global__ void wf_rotate_kernel(WavefrontCounters* c, int cur)
{
    c->n_active[cur] = 0;
    c->n_mat[0] = c->n_mat[1] = 0;
    c->iterations += 1;
}

__global__ void __launch_bounds__(kWfThreads) wf_regen_kernel(WavefrontState const w, RenderParamsF32 const prm, int cur)
{
    uint32_t const count = w.ctr->regen_count;
    uint32_t const base = w.ctr->regen_base;
    unsigned long long const item0 = w.ctr->regen_item0; // item = sample * nslots + slot
    for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        unsigned long long const item = item0 + j;
        uint32_t const sample = static_cast<uint32_t>(item / prm.nslots);
        uint32_t const slot = static_cast<uint32_t>(item - static_cast<unsigned long long>(sample) * prm.nslots);
        PathF32 p;
        p.rng = rng_open(prm.key, slot, prm.first_sample + sample);
        uint32_t x, y, sx, sy;
        slot_coords(slot, prm.width, prm.ns, x, y, sx, sy);
        gen_primary(p, prm.cams.cam, x, y, sx, sy);
        wf_store(w.active[cur], base + j, p, slot);
    }
}

template<class Shape, bool kSmemShade>
__global__ void __launch_bounds__(kWfThreads) wf_extend_kernel(WavefrontState const w, RenderParamsF32 const prm, int cur)
{
    __shared__ float4 s_shade[kSmemShade ? 4 * kSmemShadeSpheres : 1];
    __shared__ BlockAppend<3> ba; // 0 = next active (mirror bounce done here), 1 = diffuse, 2 = dielectric
    ShadePlanes sp = prm.shade;
    if constexpr(kSmemShade) {
        for(int i = threadIdx.x; i < prm.n_total; i += kWfThreads) {
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
    uint32_t* const counters[3] = { &w.ctr->n_active[cur ^ 1], &w.ctr->n_mat[0], &w.ctr->n_mat[1] };
    WfStream const& in = w.active[cur];
    uint32_t const nactive = w.ctr->n_active[cur];
    uint32_t const stride = gridDim.x * blockDim.x;
    uint32_t const rounds = (nactive + stride - 1) / stride;
    uint32_t rays = 0, nd = 0, nsp = 0, ndi = 0;
    for(uint32_t k = 0; k < rounds; ++k) {
        uint32_t const i = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
        int dest = -1;
        PathF32 p;
        uint32_t slot = 0;
        float nx = 0.f, ny = 0.f, nz = 0.f, dn = 0.f;
        if(i < nactive) {
            wf_load(in, i, p, slot);
            RayTerms const r = ray_terms(p, Shape::uniform_k ? c_scene.big_geo[0].k : 0.0f);
            float t;
            int id;
            bool const hit = closest_hit<Shape>(c_scene, prm.geo, p, r, t, id);
            rays++;
            int refl = 0;
            bool alive = shade_common(p, hit, t, id, sp, nx, ny, nz, refl);
            if(alive) {
                if(refl == 1) {
                    nsp++;
                    mirror_at_hit(p, sp.a[id]); // specular_ray, main.cpp:60-67
                    p.depth++;
                    alive = p.depth < kDepthLimit; // main.cpp:111
                    dest = 0;
                }
                else {
                    dn = fmaf(nx, p.dx, fmaf(ny, p.dy, nz * p.dz));
                    nd += refl == 0;
                    ndi += refl == 2;
                    dest = refl == 0 ? 1 : 2;
                }
            }
            if(!alive) {
                red_add_v4(prm.accum + slot, p.er, p.eg, p.eb, 1.0f);
                dest = -1;
            }
        }
        uint32_t const pos = block_append(ba, dest, counters);
        if(dest == 0) {
            wf_store(w.active[cur ^ 1], pos, p, slot);
        }
        else if(dest > 0) {
            WfStream const& out = w.mat[dest - 1];
            wf_store(out, pos, p, slot);
            out.n[pos] = make_float4(nx, ny, nz, dn);
        }
    }
    rays = warp_sum(rays);
    nd = warp_sum(nd);
    nsp = warp_sum(nsp);
    ndi = warp_sum(ndi);
    if((threadIdx.x & 31u) == 0u && rays != 0u) {
        atomicAdd(&prm.counters->rays, static_cast<unsigned long long>(rays));
        atomicAdd(&prm.counters->diffuse, static_cast<unsigned long long>(nd));
        atomicAdd(&prm.counters->specular, static_cast<unsigned long long>(nsp));
        atomicAdd(&prm.counters->dielectric, static_cast<unsigned long long>(ndi));
    }
}

// One kernel per material stream: every lane of every warp runs the same scatter function.
// kMat: 0 = diffuse (main.cpp:44-58), 1 = dielectric (main.cpp:69-97)
template<int kMat>
__global__ void __launch_bounds__(kWfThreads) wf_scatter_kernel(WavefrontState const w, RenderParamsF32 const prm, int next)
{
    __shared__ BlockAppend<1> ba;
    uint32_t* const counters[1] = { &w.ctr->n_active[next] };
    WfStream const& in = w.mat[kMat];
    uint32_t const count = w.ctr->n_mat[kMat];
    uint32_t const stride = gridDim.x * blockDim.x;
    uint32_t const rounds = (count + stride - 1) / stride;
    for(uint32_t k = 0; k < rounds; ++k) {
        uint32_t const i = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
        int dest = -1;
        PathF32 p;
        uint32_t slot = 0;
        if(i < count) {
            wf_load(in, i, p, slot);
            float4 const n = in.n[i];
            if constexpr(kMat == 0) {
                bool const front = n.w < 0.0f; // hit_record.cpp:7
                scatter_diffuse(p, front ? n.x : -n.x, front ? n.y : -n.y, front ? n.z : -n.z);
            }
            else {
                scatter_dielectric(p, n.x, n.y, n.z, n.w);
            }
            p.depth++;
            if(p.depth < kDepthLimit) { // main.cpp:111
                dest = 0;
            }
            else {
                red_add_v4(prm.accum + slot, p.er, p.eg, p.eb, 1.0f);
            }
        }
        uint32_t const pos = block_append(ba, dest, counters);
        if(dest == 0) {
            wf_store(w.active[next], pos, p, slot);
        }
    }
}

constexpr int kWfKernelsPerIteration = 6;

template<class Shape, bool kSmem>
static cudaError_t wf_enqueue_iteration(WavefrontState const& w, RenderParamsF32 const& p, int cur, int grid,
                                        cudaStream_t st)
{
    unsigned long long const total_items = static_cast<unsigned long long>(p.nslots) * p.samples;
    wf_plan_kernel<<<1, 1, 0, st>>>(w.ctr, cur, w.pool, total_items);
    wf_regen_kernel<<<grid, kWfThreads, 0, st>>>(w, p, cur);
    wf_extend_kernel<Shape, kSmem><<<grid, kWfThreads, 0, st>>>(w, p, cur);
    wf_scatter_kernel<0><<<grid, kWfThreads, 0, st>>>(w, p, cur ^ 1);
    wf_scatter_kernel<1><<<grid, kWfThreads, 0, st>>>(w, p, cur ^ 1);
    wf_rotate_kernel<<<1, 1, 0, st>>>(w.ctr, cur);
    return cudaGetLastError();
}

template<class Shape, bool kSmem>
static cudaError_t wf_run(WavefrontState const& w, RenderParamsF32 const& p, int sm_count, cudaStream_t st, int* launches)
{
    int const grid = sm_count * 8;
    wf_init_kernel<<<1, 1, 0, st>>>(w.ctr);
    cudaError_t e = cudaGetLastError();
    if(e != cudaSuccess) {
        return e;
    }
    *launches += 1;

    // capture one even + one odd iteration, replay
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if(e != cudaSuccess) {
        return e;
    }
    cudaError_t const e1 = wf_enqueue_iteration<Shape, kSmem>(w, p, 0, grid, st);
    cudaError_t const e2 = wf_enqueue_iteration<Shape, kSmem>(w, p, 1, grid, st);
    e = cudaStreamEndCapture(st, &graph);
    if(e1 != cudaSuccess || e2 != cudaSuccess || e != cudaSuccess) {
        if(graph != nullptr) {
            cudaGraphDestroy(graph);
        }
        return e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e);
    }
    e = cudaGraphInstantiate(&exec, graph, 0);
    if(e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return e;
    }

    unsigned long long const total_items = static_cast<unsigned long long>(p.nslots) * p.samples;
    WavefrontCounters h{};
    int const batch = 16; // graph replays (= 32 iterations) between host checks
    for(;;) {
        for(int k = 0; k < batch && e == cudaSuccess; ++k) {
            e = cudaGraphLaunch(exec, st);
            *launches += 2 * kWfKernelsPerIteration;
        }
        if(e == cudaSuccess) {
            e = cudaMemcpyAsync(&h, w.ctr, sizeof(h), cudaMemcpyDeviceToHost, st);
        }
        if(e == cudaSuccess) {
            e = cudaStreamSynchronize(st);
        }
        if(e != cudaSuccess) {
            break;
        }
        // after a whole replay the odd stream is empty and the even one holds the survivors
        if(h.cursor >= total_items && h.n_active[0] == 0 && h.n_active[1] == 0) {
            break;
        }
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    return e;
}

// planes per stream: a b c d (+ n for material streams); float4 planes are laid out stream after stream
static WfStream wf_carve(float4*& planes, uint32_t*& words, uint32_t pool, bool with_normal)
{
    WfStream s{};
    s.a = planes;
    s.b = planes + pool;
    s.c = planes + 2 * static_cast<size_t>(pool);
    s.d = planes + 3 * static_cast<size_t>(pool);
    planes += 4 * static_cast<size_t>(pool);
    if(with_normal) {
        s.n = planes;
        planes += pool;
    }
    s.slot = words;
    words += pool;
    return s;
}

cudaError_t launch_wavefront(WavefrontBuffers const& buf, RenderParamsF32 const& p, SceneCounts const& c, int sm_count,
                             cudaStream_t stream, int* launches)
{
    WavefrontState w{};
    float4* planes = buf.planes;
    uint32_t* words = buf.words;
    w.active[0] = wf_carve(planes, words, buf.pool, false);
    w.active[1] = wf_carve(planes, words, buf.pool, false);
    w.mat[0] = wf_carve(planes, words, buf.pool, true);
    w.mat[1] = wf_carve(planes, words, buf.pool, true);
    w.ctr = buf.counters;
    w.pool = buf.pool;
    bool const smem = p.n_total <= kSmemShadeSpheres;
#define X(a, b, cc, d, bx, by, bz, uk, em, pm) \
    if(PTB_COUNTS_MATCH(c, a, b, cc, d, bx, by, bz, uk, em, pm) && smem) { \
        return wf_run<SceneShape<(a), (b), (cc), (d), (bx), (by), (bz), (uk), (em), (pm)>, true>(w, p, sm_count, stream, launches); \
    }
    PTB_MEGA_SPECIALISATIONS(X)
#undef X
    if(smem) {
        return wf_run<GenericShape, true>(w, p, sm_count, stream, launches);
    }
    return wf_run<GenericShape, false>(w, p, sm_count, stream, launches);
}

} // namespace ptb
