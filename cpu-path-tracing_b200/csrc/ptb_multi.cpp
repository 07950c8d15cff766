// ptb_multi.cpp -- several GPUs behind the entry points of include/ptb200.h.
//
// The reference fills its image with one blocking call (src/main.cpp:214-236: a taskflow task per row on the host's
// cores).  Here the samples of every sub-pixel are split over the GPUs (ptb_sample_share), every GPU accumulates its
// share in its own float4 buffer, and the buffers are summed before the non-linear resolve (main.cpp:192-196).
//
//   Group     ptb_create_multi: one process, n member contexts; the handle broadcasts scene / camera / image calls,
//             runs ptb_render on one host thread per GPU, and sums + resolves in ptb_resolve*.
//   RankComm  ptb_comm_init_rank: one process per GPU (torchrun); the same calls, ptb_resolve* collective.
//
// Two transports for the sum (PTB_TRANSPORT_*):
//   PEER  resolve_kernel (ptb_resolve.cu) launched on EVERY GPU over its share of the rows with all GPUs' buffers as
//         peer mappings: reduce-scatter + resolve + gather in one pass, pixels stored straight into the root's image.
//         Peer mappings come from cudaDeviceEnablePeerAccess (one process) or CUDA IPC handles exchanged through the
//         communicator (one process per GPU).
//   NCCL  ncclReduce(sum) into a scratch buffer on the root, then the single-GPU resolve.  libnccl.so.2 is loaded at
//         run time (dlopen), the communicator belongs to the handle.
#include "ptb_context.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <thread>

namespace ptb {

// ---- libnccl.so.2 at run time ---------------------------------------------------------------------------------------------
// Only what this file calls, declared by hand (the library is not a build dependency): nccl.h 2.27 signatures.
namespace {

struct NcclUniqueId
{
    char internal[PTB_COMM_ID_BYTES]; // NCCL_UNIQUE_ID_BYTES
};
using NcclComm = void*;
constexpr int kNcclSum = 0, kNcclChar = 0, kNcclInt = 2, kNcclFloat = 7, kNcclDouble = 8;

struct NcclApi
{
    void* lib = nullptr;
    int version = 0;
    std::string error;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, int const*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(void const*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(void const*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(void const*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    char const* (*GetErrorString)(int) = nullptr;
    bool ok() const { return lib != nullptr; }
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // Order: PTB_NCCL_LIB (the Python binding points it at the libnccl its torch would load, so that one process never
        // holds two NCCL builds under one soname), a libnccl.so.2 the process has loaded already, the system's.
        char const* env = std::getenv("PTB_NCCL_LIB");
        if(env != nullptr && *env != '\0') {
            api.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        }
        if(api.lib == nullptr) {
            api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        }
        for(char const* n : { "libnccl.so.2", "libnccl.so" }) {
            if(api.lib == nullptr) {
                api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            }
        }
        if(api.lib == nullptr) {
            api.error = "libnccl.so.2 not found (set PTB_NCCL_LIB)";
            return;
        }
        bool all = true;
        auto sym = [&](auto& fn, char const* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.lib, name));
            all = all && fn != nullptr;
        };
        sym(api.GetVersion, "ncclGetVersion");
        sym(api.GetUniqueId, "ncclGetUniqueId");
        sym(api.CommInitRank, "ncclCommInitRank");
        sym(api.CommInitAll, "ncclCommInitAll");
        sym(api.CommDestroy, "ncclCommDestroy");
        sym(api.Reduce, "ncclReduce");
        sym(api.AllReduce, "ncclAllReduce");
        sym(api.AllGather, "ncclAllGather");
        sym(api.GroupStart, "ncclGroupStart");
        sym(api.GroupEnd, "ncclGroupEnd");
        sym(api.GetErrorString, "ncclGetErrorString");
        if(!all) {
            api.error = "libnccl.so.2 lacks an entry point this library needs";
            api.lib = nullptr;
            return;
        }
        api.GetVersion(&api.version);
    });
    return api;
}

int fail_nccl(ptb_context* ctx, int rc, char const* what)
{
    ctx->err = std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + nccl().GetErrorString(rc) + ")";
    return PTB_ERR_CUDA;
}

#define PTB_NCCL(ctx, call) \
    do { \
        int const r_ = (call); \
        if(r_ != 0) { \
            return fail_nccl((ctx), r_, #call); \
        } \
    } while(0)
#define PTB_CU(ctx, call) \
    do { \
        cudaError_t const e_ = (call); \
        if(e_ != cudaSuccess) { \
            return api_fail_cuda((ctx), e_, #call); \
        } \
    } while(0)

// cuMemGetAddressRange: an IPC handle names the ALLOCATION a pointer lives in -- small cudaMalloc blocks are carved out
// of larger ones -- so the importer needs the pointer's offset inside it.  The runtime API has no call for that.
using CuGetRange = int (*)(unsigned long long*, size_t*, unsigned long long);
CuGetRange cu_get_range()
{
    static CuGetRange fn = [] {
        void* lib = nullptr;
        for(char const* n : { "libcuda.so.1", "libcuda.so" }) {
            if((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) {
                break;
            }
        }
        return lib != nullptr ? reinterpret_cast<CuGetRange>(dlsym(lib, "cuMemGetAddressRange_v2")) : nullptr;
    }();
    return fn;
}

void share_of(uint32_t total, int n, int rank, uint32_t& first, uint32_t& count)
{
    uint32_t const base = total / static_cast<uint32_t>(n), extra = total % static_cast<uint32_t>(n);
    uint32_t const r = static_cast<uint32_t>(rank);
    first = r * base + std::min(r, extra);
    count = base + (r < extra ? 1u : 0u);
}

void rows_of(int height, int n, int rank, uint32_t& y0, uint32_t& y1)
{
    uint32_t first, count;
    share_of(static_cast<uint32_t>(height), n, rank, first, count);
    y0 = first;
    y1 = first + count;
}

} // namespace

// =====================================================================================================================
// One process, several GPUs
// =====================================================================================================================
struct Group
{
    std::vector<ptb_context*> members; // members[0] owns the image
    bool peer_ok = false;
    int transport_pref = PTB_TRANSPORT_AUTO;
    int last_transport = 0;
    std::vector<NcclComm> comms; // made on first use of the NCCL transport
    float4* d_sum32 = nullptr;   // NCCL transport: the reduced buffers on the root
    double* d_sum64 = nullptr;
    size_t sum_slots = 0;
    double last_resolve_ms = 0.0;
};

ptb_context* multi_root(ptb_context* ctx)
{
    return ctx->group->members[0];
}

namespace {

int adopt_error(ptb_context* handle, ptb_context* member, int rc)
{
    if(rc != PTB_OK) {
        handle->err = "GPU " + std::to_string(member->device) + ": " + member->err;
    }
    return rc;
}

template<class F>
int for_each_member(ptb_context* handle, F&& f)
{
    for(ptb_context* m : handle->group->members) {
        int const rc = f(m);
        if(rc != PTB_OK) {
            return adopt_error(handle, m, rc);
        }
    }
    return PTB_OK;
}

// run f(member) on one host thread per GPU; first failure wins
template<class F>
int in_parallel(ptb_context* handle, F&& f)
{
    std::vector<ptb_context*> const& ms = handle->group->members;
    std::vector<int> rc(ms.size(), PTB_OK);
    std::vector<std::thread> th;
    th.reserve(ms.size());
    bool started = true;
    try {
        for(size_t g = 1; g < ms.size(); ++g) {
            th.emplace_back([&, g] { rc[g] = f(ms[g], static_cast<int>(g)); }); // f is an entry point: it does not throw
        }
    }
    catch(...) { // std::system_error: no thread to be had -- the ones that did start must still be joined
        started = false;
    }
    if(started) {
        rc[0] = f(ms[0], 0);
    }
    for(std::thread& t : th) {
        t.join();
    }
    if(!started) {
        return api_fail(handle, PTB_ERR_INTERNAL, "cannot start one host thread per GPU");
    }
    for(size_t g = 0; g < ms.size(); ++g) {
        if(rc[g] != PTB_OK) {
            return adopt_error(handle, ms[g], rc[g]);
        }
    }
    return PTB_OK;
}

int group_nccl_comms(ptb_context* handle)
{
    Group* G = handle->group;
    if(!G->comms.empty()) {
        return PTB_OK;
    }
    if(!nccl().ok()) {
        return api_fail(handle, PTB_ERR_STATE, nccl().error.c_str());
    }
    std::vector<int> devs;
    for(ptb_context* m : G->members) {
        devs.push_back(m->device);
    }
    G->comms.assign(devs.size(), nullptr);
    int const rc = nccl().CommInitAll(G->comms.data(), static_cast<int>(devs.size()), devs.data());
    if(rc != 0) {
        G->comms.clear();
        return fail_nccl(handle, rc, "ncclCommInitAll");
    }
    return PTB_OK;
}

} // namespace

void multi_destroy(ptb_context* ctx)
{
    Group* G = ctx->group;
    for(ptb_context* m : G->members) {
        cudaSetDevice(m->device);
        cudaStreamSynchronize(m->stream);
    }
    for(NcclComm c : G->comms) {
        if(c != nullptr) {
            nccl().CommDestroy(c);
        }
    }
    if(!G->members.empty()) {
        cudaSetDevice(G->members[0]->device);
        cudaFree(G->d_sum32);
        cudaFree(G->d_sum64);
    }
    for(ptb_context* m : G->members) {
        ptb_destroy(m);
    }
    delete G;
    ctx->group = nullptr;
}

int multi_synchronize(ptb_context* ctx)
{
    return for_each_member(ctx, [](ptb_context* m) { return ptb_synchronize(m); });
}

int multi_upload_scene(ptb_context* ctx, void const* spheres, size_t count, size_t stride)
{
    // the packing (frame choice, hierarchy build for large scenes) is host work per member: do it side by side
    return in_parallel(ctx, [&](ptb_context* m, int) { return ptb_upload_scene(m, spheres, count, stride); });
}

int multi_set_camera(ptb_context* ctx, void const* camera, size_t bytes)
{
    return in_parallel(ctx, [&](ptb_context* m, int) { return ptb_set_camera(m, camera, bytes); });
}

int multi_set_smallpt_camera(ptb_context* ctx, double const* cam8)
{
    return in_parallel(ctx, [&](ptb_context* m, int) { return ptb_set_smallpt_camera(m, cam8); });
}

int multi_set_image(ptb_context* ctx, int width, int height, int ns)
{
    int const rc = for_each_member(ctx, [&](ptb_context* m) { return ptb_set_image(m, width, height, ns); });
    if(rc == PTB_OK) {
        ctx->width = width;
        ctx->height = height;
        ctx->ns = ns;
        ctx->nslots = multi_root(ctx)->nslots;
    }
    return rc;
}

int multi_clear(ptb_context* ctx)
{
    return for_each_member(ctx, [](ptb_context* m) { return ptb_clear(m); });
}

int multi_render(ptb_context* ctx, uint64_t seed, uint32_t first_sample, uint32_t samples, uint32_t flags)
{
    int const n = static_cast<int>(ctx->group->members.size());
    return in_parallel(ctx, [&](ptb_context* m, int g) {
        uint32_t first, count;
        share_of(samples, n, g, first, count);
        return ptb_render(m, seed, first_sample + first, count, flags);
    });
}

int multi_resolve(ptb_context* ctx, double* rgb_out, uint8_t* rgb8_out, void** device_rgb)
{
    Group* G = ctx->group;
    std::vector<ptb_context*> const& ms = G->members;
    ptb_context* root = ms[0];
    int const n = static_cast<int>(ms.size());
    if(root->width <= 0 || !root->have_scene) {
        return api_fail(ctx, PTB_ERR_STATE, "ptb_resolve: scene / image not set");
    }
    if(rgb_out == nullptr && rgb8_out == nullptr && device_rgb == nullptr) {
        return api_fail(ctx, PTB_ERR_ARGUMENT, "ptb_resolve: output pointer is null");
    }
    bool any64 = false;
    for(ptb_context* m : ms) {
        any64 = any64 || m->accum64_used;
    }
    bool const want_f64 = rgb_out != nullptr || device_rgb != nullptr;
    uint32_t const W = static_cast<uint32_t>(root->width), H = static_cast<uint32_t>(root->height), NS = static_cast<uint32_t>(root->ns);
    bool const peer = G->transport_pref != PTB_TRANSPORT_NCCL && G->peer_ok;
    if(!peer && G->transport_pref == PTB_TRANSPORT_PEER) {
        // asked for, impossible on this machine: say so once in the handle's message, then go on over NCCL
        ctx->err = "peer mapping unavailable between the GPUs of this context: NCCL transport used";
    }

    if(peer) {
        ResolveSources src{};
        src.n = n;
        for(int g = 0; g < n; ++g) {
            src.accum32[g] = api_active_accum(ms[static_cast<size_t>(g)]);
            src.accum64[g] = ms[static_cast<size_t>(g)]->accum64_used ? ms[static_cast<size_t>(g)]->d_accum64 : nullptr;
        }
        for(int g = 0; g < n; ++g) {
            ptb_context* m = ms[static_cast<size_t>(g)];
            uint32_t y0, y1;
            rows_of(root->height, n, g, y0, y1);
            PTB_CU(ctx, cudaSetDevice(m->device));
            PTB_CU(ctx, cudaEventRecord(m->ev0, m->stream));
            PTB_CU(ctx, launch_resolve_rows(src, W, H, NS, y0, y1, want_f64 ? root->d_rgb : nullptr,
                                            rgb8_out != nullptr ? root->d_rgb8 : nullptr, m->stream));
            PTB_CU(ctx, cudaEventRecord(m->ev1, m->stream));
            m->stats.kernel_launches += 1;
        }
        double worst = 0.0;
        for(ptb_context* m : ms) {
            PTB_CU(ctx, cudaSetDevice(m->device));
            PTB_CU(ctx, cudaStreamSynchronize(m->stream));
            float ms_ = 0.0f;
            PTB_CU(ctx, cudaEventElapsedTime(&ms_, m->ev0, m->ev1));
            worst = std::max(worst, static_cast<double>(ms_));
        }
        G->last_resolve_ms = worst;
        G->last_transport = PTB_TRANSPORT_PEER;
    }
    else {
        int rc = group_nccl_comms(ctx);
        if(rc != PTB_OK) {
            return rc;
        }
        PTB_CU(ctx, cudaSetDevice(root->device));
        if(G->sum_slots < root->nslots) {
            cudaFree(G->d_sum32);
            cudaFree(G->d_sum64);
            G->d_sum32 = nullptr;
            G->d_sum64 = nullptr;
            G->sum_slots = 0;
            PTB_CU(ctx, cudaMalloc(&G->d_sum32, root->nslots * sizeof(float4)));
            G->sum_slots = root->nslots;
        }
        if(any64 && G->d_sum64 == nullptr) {
            PTB_CU(ctx, cudaMalloc(&G->d_sum64, G->sum_slots * 4 * sizeof(double)));
        }
        if(any64) {
            for(ptb_context* m : ms) { // every rank of the collective needs a buffer, used or not
                if(m->d_accum64 == nullptr) {
                    PTB_CU(ctx, cudaSetDevice(m->device));
                    PTB_CU(ctx, cudaMalloc(&m->d_accum64, m->nslots * 4 * sizeof(double)));
                    PTB_CU(ctx, cudaMemsetAsync(m->d_accum64, 0, m->nslots * 4 * sizeof(double), m->stream));
                    m->buffers_epoch++;
                }
            }
        }
        PTB_CU(ctx, cudaSetDevice(root->device));
        PTB_CU(ctx, cudaEventRecord(root->ev0, root->stream));
        PTB_NCCL(ctx, nccl().GroupStart());
        for(int g = 0; g < n; ++g) {
            ptb_context* m = ms[static_cast<size_t>(g)];
            PTB_NCCL(ctx, nccl().Reduce(api_active_accum(m), G->d_sum32, root->nslots * 4, kNcclFloat, kNcclSum, 0,
                                        G->comms[static_cast<size_t>(g)], m->stream));
        }
        PTB_NCCL(ctx, nccl().GroupEnd());
        if(any64) {
            PTB_NCCL(ctx, nccl().GroupStart());
            for(int g = 0; g < n; ++g) {
                ptb_context* m = ms[static_cast<size_t>(g)];
                PTB_NCCL(ctx, nccl().Reduce(m->d_accum64, G->d_sum64, root->nslots * 4, kNcclDouble, kNcclSum, 0,
                                            G->comms[static_cast<size_t>(g)], m->stream));
            }
            PTB_NCCL(ctx, nccl().GroupEnd());
        }
        PTB_CU(ctx, cudaSetDevice(root->device));
        PTB_CU(ctx, launch_resolve(G->d_sum32, any64 ? G->d_sum64 : nullptr, W, H, NS, want_f64 ? root->d_rgb : nullptr,
                                   rgb8_out != nullptr ? root->d_rgb8 : nullptr, root->stream));
        PTB_CU(ctx, cudaEventRecord(root->ev1, root->stream));
        root->stats.kernel_launches += 1;
        for(ptb_context* m : ms) {
            PTB_CU(ctx, cudaSetDevice(m->device));
            PTB_CU(ctx, cudaStreamSynchronize(m->stream));
        }
        float ms_ = 0.0f;
        PTB_CU(ctx, cudaEventElapsedTime(&ms_, root->ev0, root->ev1));
        G->last_resolve_ms = ms_;
        G->last_transport = PTB_TRANSPORT_NCCL;
    }

    PTB_CU(ctx, cudaSetDevice(root->device));
    size_t const npix = static_cast<size_t>(W) * H;
    if(rgb_out != nullptr) {
        PTB_CU(ctx, cudaMemcpyAsync(rgb_out, root->d_rgb, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, root->stream));
    }
    if(rgb8_out != nullptr) {
        PTB_CU(ctx, cudaMemcpyAsync(rgb8_out, root->d_rgb8, npix * 3, cudaMemcpyDeviceToHost, root->stream));
    }
    PTB_CU(ctx, cudaStreamSynchronize(root->stream));
    if(device_rgb != nullptr) {
        *device_rgb = root->d_rgb;
    }
    return PTB_OK;
}

// Checkpoint of a group = the SUM of the members' buffers (what a single GPU would hold); restoring puts it on the
// root and zeroes the others, so that a checkpoint moves freely between GPU counts.
int multi_download_accum(ptb_context* ctx, float* out, size_t floats)
{
    std::vector<ptb_context*> const& ms = ctx->group->members;
    size_t const need = multi_root(ctx)->nslots * 4;
    if(out == nullptr || floats < need) {
        return api_fail(ctx, PTB_ERR_ARGUMENT, "ptb_download_accum: output too small");
    }
    int rc = adopt_error(ctx, ms[0], ptb_download_accum(ms[0], out, floats));
    std::vector<float> tmp(ms.size() > 1 ? need : 0);
    for(size_t g = 1; g < ms.size() && rc == PTB_OK; ++g) {
        rc = adopt_error(ctx, ms[g], ptb_download_accum(ms[g], tmp.data(), tmp.size()));
        for(size_t i = 0; i < need && rc == PTB_OK; ++i) {
            out[i] += tmp[i];
        }
    }
    return rc;
}

int multi_download_accum64(ptb_context* ctx, double* out, size_t doubles)
{
    std::vector<ptb_context*> const& ms = ctx->group->members;
    size_t const need = multi_root(ctx)->nslots * 4;
    if(out == nullptr || doubles < need) {
        return api_fail(ctx, PTB_ERR_ARGUMENT, "ptb_download_accum64: output too small");
    }
    int rc = adopt_error(ctx, ms[0], ptb_download_accum64(ms[0], out, doubles));
    std::vector<double> tmp(ms.size() > 1 ? need : 0);
    for(size_t g = 1; g < ms.size() && rc == PTB_OK; ++g) {
        rc = adopt_error(ctx, ms[g], ptb_download_accum64(ms[g], tmp.data(), tmp.size()));
        for(size_t i = 0; i < need && rc == PTB_OK; ++i) {
            out[i] += tmp[i];
        }
    }
    return rc;
}

int multi_upload_accum(ptb_context* ctx, float const* in, size_t floats)
{
    std::vector<ptb_context*> const& ms = ctx->group->members;
    int rc = adopt_error(ctx, ms[0], ptb_upload_accum(ms[0], in, floats));
    for(size_t g = 1; g < ms.size() && rc == PTB_OK; ++g) {
        PTB_CU(ctx, cudaSetDevice(ms[g]->device));
        PTB_CU(ctx, cudaMemsetAsync(api_active_accum(ms[g]), 0, ms[g]->nslots * sizeof(float4), ms[g]->stream));
        PTB_CU(ctx, cudaStreamSynchronize(ms[g]->stream));
    }
    return rc;
}

int multi_upload_accum64(ptb_context* ctx, double const* in, size_t doubles)
{
    std::vector<ptb_context*> const& ms = ctx->group->members;
    int rc = adopt_error(ctx, ms[0], ptb_upload_accum64(ms[0], in, doubles));
    for(size_t g = 1; g < ms.size() && rc == PTB_OK; ++g) {
        if(ms[g]->d_accum64 != nullptr) {
            PTB_CU(ctx, cudaSetDevice(ms[g]->device));
            PTB_CU(ctx, cudaMemsetAsync(ms[g]->d_accum64, 0, ms[g]->nslots * 4 * sizeof(double), ms[g]->stream));
            PTB_CU(ctx, cudaStreamSynchronize(ms[g]->stream));
        }
    }
    return rc;
}

int multi_get_stats(ptb_context* ctx, ptb_stats* out)
{
    ptb_stats sum{};
    int const rc = for_each_member(ctx, [&](ptb_context* m) {
        ptb_stats s{};
        int const r = ptb_get_stats(m, &s);
        sum.paths += s.paths;
        sum.rays += s.rays;
        sum.kernel_launches += s.kernel_launches;
        sum.hits_diffuse += s.hits_diffuse;
        sum.hits_specular += s.hits_specular;
        sum.hits_dielectric += s.hits_dielectric;
        sum.last_render_ms = std::max(sum.last_render_ms, s.last_render_ms);   // the members run side by side:
        sum.total_render_ms = std::max(sum.total_render_ms, s.total_render_ms); // the job takes as long as the slowest
        return r;
    });
    sum.last_resolve_ms = ctx->group->last_resolve_ms;
    *out = sum;
    return rc;
}

// =====================================================================================================================
// One process per GPU
// =====================================================================================================================
namespace {

// what a rank tells the others about its buffers before a resolve
struct PeerRecord
{
    uint64_t epoch;
    uint64_t nslots;
    int32_t can_export; // internal (cudaMalloc'ed) accumulation buffer in use, offsets known
    int32_t has64;
    int32_t want_peer;  // PTB_TRANSPORT_NCCL not requested
    int32_t want;       // rank 0 only, per call: bit 0 = FP64 image wanted, bit 1 = 8-bit image wanted
    uint64_t off32, off64, off_rgb, off_rgb8;
    cudaIpcMemHandle_t h32, h64, h_rgb, h_rgb8;
};

struct Imported
{
    cudaIpcMemHandle_t handle;
    void* base;
};

} // namespace

struct RankComm
{
    NcclComm comm = nullptr;
    int n = 1, rank = 0;
    int transport_pref = PTB_TRANSPORT_AUTO;
    int last_transport = 0;
    bool peer_failed = false; // a mapping attempt failed somewhere: stay on NCCL
    PeerRecord mine{};
    uint64_t mine_epoch = ~0ull;
    bool mine_external = false;
    std::vector<PeerRecord> seen;    // records the current mappings were made from
    std::vector<Imported> imported;  // opened IPC allocations
    ResolveSources src{};
    double* root_rgb = nullptr;
    uint8_t* root_rgb8 = nullptr;
    bool mapped = false;
    unsigned char* d_xchg = nullptr; // (n + 1) records: [0] mine, [1..n] everyone's
    PeerRecord* h_xchg = nullptr;    // pinned, n records
    int* d_flag = nullptr;           // barrier / agreement scratch (2 ints)
    int* h_flag = nullptr;           // pinned
    float4* d_sum32 = nullptr;       // root, NCCL transport
    double* d_sum64 = nullptr;
    size_t sum_slots = 0;
};

namespace {

void close_imports(RankComm* C)
{
    for(Imported const& im : C->imported) {
        cudaIpcCloseMemHandle(im.base);
    }
    C->imported.clear();
    C->mapped = false;
    C->seen.clear();
}

void* import_handle(RankComm* C, cudaIpcMemHandle_t const& h)
{
    for(Imported const& im : C->imported) {
        if(std::memcmp(&im.handle, &h, sizeof(h)) == 0) {
            return im.base; // one allocation may hold several of a peer's (small) buffers: open it once
        }
    }
    void* base = nullptr;
    if(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    C->imported.push_back(Imported{ h, base });
    return base;
}

bool export_buffer(void* ptr, cudaIpcMemHandle_t& h, uint64_t& off)
{
    std::memset(&h, 0, sizeof(h));
    off = 0;
    if(ptr == nullptr) {
        return true;
    }
    CuGetRange const range = cu_get_range();
    unsigned long long base = 0;
    size_t size = 0;
    if(range == nullptr || range(&base, &size, reinterpret_cast<unsigned long long>(ptr)) != 0) {
        return false;
    }
    if(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    off = reinterpret_cast<unsigned long long>(ptr) - base;
    return true;
}

// same buffers behind two records? (everything but the per-call `want`)
bool same_mapping(PeerRecord const& a, PeerRecord const& b)
{
    return a.epoch == b.epoch && a.nslots == b.nslots && a.can_export == b.can_export && a.has64 == b.has64 &&
           a.off32 == b.off32 && a.off64 == b.off64 && a.off_rgb == b.off_rgb && a.off_rgb8 == b.off_rgb8 &&
           std::memcmp(&a.h32, &b.h32, 4 * sizeof(cudaIpcMemHandle_t)) == 0;
}

void fill_record(ptb_context* ctx, int want)
{
    RankComm* C = ctx->comm;
    bool const external = ctx->ext_accum != nullptr; // a caller-owned buffer may come from any allocator: not exported
    if(C->mine_epoch != ctx->buffers_epoch || C->mine_external != external) {
        PeerRecord r{};
        r.epoch = ctx->buffers_epoch;
        r.nslots = ctx->nslots;
        bool ok = !external;
        ok = ok && export_buffer(ctx->d_accum, r.h32, r.off32);
        ok = ok && export_buffer(ctx->d_accum64, r.h64, r.off64);
        ok = ok && export_buffer(ctx->d_rgb, r.h_rgb, r.off_rgb);
        ok = ok && export_buffer(ctx->d_rgb8, r.h_rgb8, r.off_rgb8);
        r.can_export = ok ? 1 : 0;
        C->mine = r;
        C->mine_epoch = ctx->buffers_epoch;
        C->mine_external = external;
    }
    C->mine.has64 = ctx->accum64_used ? 1 : 0;
    C->mine.want_peer = C->transport_pref != PTB_TRANSPORT_NCCL ? 1 : 0;
    C->mine.want = want;
}

// stream-ordered barrier that also sums a flag over the ranks (min via negative sum is not needed: flags are 0 / 1)
int all_sum(ptb_context* ctx, int value, int* result)
{
    RankComm* C = ctx->comm;
    *C->h_flag = value;
    PTB_CU(ctx, cudaMemcpyAsync(C->d_flag, C->h_flag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PTB_NCCL(ctx, nccl().AllReduce(C->d_flag, C->d_flag + 1, 1, kNcclInt, kNcclSum, C->comm, ctx->stream));
    if(result != nullptr) {
        PTB_CU(ctx, cudaMemcpyAsync(C->h_flag, C->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PTB_CU(ctx, cudaStreamSynchronize(ctx->stream));
        *result = *C->h_flag;
    }
    return PTB_OK;
}

} // namespace

// Called (by every rank) before buffers that peers may have mapped are freed: ptb_set_image, ptb_destroy.
int comm_release_peers(ptb_context* ctx)
{
    RankComm* C = ctx->comm;
    if(C == nullptr) {
        return PTB_OK;
    }
    PTB_CU(ctx, cudaSetDevice(ctx->device));
    close_imports(C);
    int dummy = 0;
    return all_sum(ctx, 0, &dummy); // nobody frees before everybody has unmapped
}

void comm_destroy(ptb_context* ctx)
{
    RankComm* C = ctx->comm;
    if(C == nullptr) {
        return;
    }
    cudaSetDevice(ctx->device);
    if(C->comm != nullptr) {
        int dummy = 0;
        close_imports(C);
        (void)all_sum(ctx, 0, &dummy);
        nccl().CommDestroy(C->comm);
    }
    cudaFree(C->d_xchg);
    cudaFree(C->d_flag);
    cudaFree(C->d_sum32);
    cudaFree(C->d_sum64);
    cudaFreeHost(C->h_xchg);
    cudaFreeHost(C->h_flag);
    delete C;
    ctx->comm = nullptr;
}

int comm_resolve(ptb_context* ctx, double* rgb_out, uint8_t* rgb8_out, void** device_rgb)
{
    RankComm* C = ctx->comm;
    bool const root = C->rank == 0;
    if(ctx->width <= 0 || !ctx->have_scene || api_active_accum(ctx) == nullptr) {
        return api_fail(ctx, PTB_ERR_STATE, "ptb_resolve: scene / image not set");
    }
    if(root && rgb_out == nullptr && rgb8_out == nullptr && device_rgb == nullptr) {
        return api_fail(ctx, PTB_ERR_ARGUMENT, "ptb_resolve: output pointer is null on rank 0");
    }
    PTB_CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t const st = ctx->stream;
    int const n = C->n;

    // 1. everybody's record (doubles as the barrier "all ranks have finished rendering": ptb_render blocks).  Which
    //    image rank 0 wants travels in its record: it decides which of its buffers the others store pixels into.
    fill_record(ctx, root ? ((rgb_out != nullptr || device_rgb != nullptr ? 1 : 0) | (rgb8_out != nullptr ? 2 : 0)) : 0);
    PTB_CU(ctx, cudaMemcpyAsync(C->d_xchg, &C->mine, sizeof(PeerRecord), cudaMemcpyHostToDevice, st));
    PTB_NCCL(ctx, nccl().AllGather(C->d_xchg, C->d_xchg + sizeof(PeerRecord), sizeof(PeerRecord), kNcclChar, C->comm, st));
    PTB_CU(ctx, cudaMemcpyAsync(C->h_xchg, C->d_xchg + sizeof(PeerRecord), static_cast<size_t>(n) * sizeof(PeerRecord),
                                cudaMemcpyDeviceToHost, st));
    PTB_CU(ctx, cudaStreamSynchronize(st));
    bool peer = !C->peer_failed, any64 = false;
    for(int g = 0; g < n; ++g) {
        PeerRecord const& r = C->h_xchg[g];
        peer = peer && r.can_export != 0 && r.want_peer != 0;
        any64 = any64 || r.has64 != 0;
        if(r.nslots != ctx->nslots) {
            return api_fail(ctx, PTB_ERR_STATE, "ptb_resolve: the ranks of this job hold images of different geometry");
        }
    }
    int const want_all = C->h_xchg[0].want;

    // 2. (re)map when a record changed; every rank must succeed or nobody uses the mappings
    if(peer) {
        bool stale = !C->mapped || C->seen.size() != static_cast<size_t>(n);
        for(int g = 0; g < n && !stale; ++g) {
            stale = !same_mapping(C->seen[static_cast<size_t>(g)], C->h_xchg[g]);
        }
        if(stale) {
            close_imports(C);
            bool ok = true;
            ResolveSources src{};
            src.n = n;
            for(int g = 0; g < n && ok; ++g) {
                PeerRecord const& r = C->h_xchg[g];
                if(g == C->rank) {
                    src.accum32[g] = ctx->d_accum;
                    src.accum64[g] = r.has64 != 0 ? ctx->d_accum64 : nullptr;
                    continue;
                }
                auto* b32 = static_cast<unsigned char*>(import_handle(C, r.h32));
                ok = ok && b32 != nullptr;
                src.accum32[g] = ok ? reinterpret_cast<float4 const*>(b32 + r.off32) : nullptr;
                if(ok && r.has64 != 0) {
                    auto* b64 = static_cast<unsigned char*>(import_handle(C, r.h64));
                    ok = ok && b64 != nullptr;
                    src.accum64[g] = ok ? reinterpret_cast<double const*>(b64 + r.off64) : nullptr;
                }
            }
            if(ok && !root) {
                PeerRecord const& r0 = C->h_xchg[0];
                auto* brgb = static_cast<unsigned char*>(import_handle(C, r0.h_rgb));
                auto* brgb8 = static_cast<unsigned char*>(import_handle(C, r0.h_rgb8));
                ok = brgb != nullptr && brgb8 != nullptr;
                C->root_rgb = ok ? reinterpret_cast<double*>(brgb + r0.off_rgb) : nullptr;
                C->root_rgb8 = ok ? reinterpret_cast<uint8_t*>(brgb8 + r0.off_rgb8) : nullptr;
            }
            else if(ok) {
                C->root_rgb = ctx->d_rgb;
                C->root_rgb8 = ctx->d_rgb8;
            }
            int failures = 0;
            int const rc = all_sum(ctx, ok ? 0 : 1, &failures);
            if(rc != PTB_OK) {
                return rc;
            }
            if(failures != 0) {
                close_imports(C);
                C->peer_failed = true; // CUDA IPC is not usable between these processes: NCCL from now on
                peer = false;
            }
            else {
                C->src = src;
                C->seen.assign(C->h_xchg, C->h_xchg + n);
                C->mapped = true;
            }
        }
    }

    bool const want_f64 = (want_all & 1) != 0, want_u8 = (want_all & 2) != 0;
    uint32_t const W = static_cast<uint32_t>(ctx->width), H = static_cast<uint32_t>(ctx->height), NS = static_cast<uint32_t>(ctx->ns);
    PTB_CU(ctx, cudaEventRecord(ctx->ev0, st));
    if(peer) {
        uint32_t y0, y1;
        rows_of(ctx->height, n, C->rank, y0, y1);
        PTB_CU(ctx, launch_resolve_rows(C->src, W, H, NS, y0, y1, want_f64 ? C->root_rgb : nullptr, want_u8 ? C->root_rgb8 : nullptr, st));
        ctx->stats.kernel_launches += 1;
        // nobody returns (and clears or frees a buffer) before every GPU has read what it needs and stored its pixels
        int const rc = all_sum(ctx, 0, nullptr);
        if(rc != PTB_OK) {
            return rc;
        }
        C->last_transport = PTB_TRANSPORT_PEER;
    }
    else {
        if(root && C->sum_slots < ctx->nslots) {
            cudaFree(C->d_sum32);
            cudaFree(C->d_sum64);
            C->d_sum32 = nullptr;
            C->d_sum64 = nullptr;
            C->sum_slots = 0;
            PTB_CU(ctx, cudaMalloc(&C->d_sum32, ctx->nslots * sizeof(float4)));
            C->sum_slots = ctx->nslots;
        }
        if(root && any64 && C->d_sum64 == nullptr) {
            PTB_CU(ctx, cudaMalloc(&C->d_sum64, C->sum_slots * 4 * sizeof(double)));
        }
        if(any64 && ctx->d_accum64 == nullptr) {
            PTB_CU(ctx, cudaMalloc(&ctx->d_accum64, ctx->nslots * 4 * sizeof(double)));
            PTB_CU(ctx, cudaMemsetAsync(ctx->d_accum64, 0, ctx->nslots * 4 * sizeof(double), st));
            ctx->buffers_epoch++;
        }
        PTB_NCCL(ctx, nccl().Reduce(api_active_accum(ctx), C->d_sum32, ctx->nslots * 4, kNcclFloat, kNcclSum, 0, C->comm, st));
        if(any64) {
            PTB_NCCL(ctx, nccl().Reduce(ctx->d_accum64, C->d_sum64, ctx->nslots * 4, kNcclDouble, kNcclSum, 0, C->comm, st));
        }
        if(root) {
            PTB_CU(ctx, launch_resolve(C->d_sum32, any64 ? C->d_sum64 : nullptr, W, H, NS, want_f64 ? ctx->d_rgb : nullptr,
                                       want_u8 ? ctx->d_rgb8 : nullptr, st));
            ctx->stats.kernel_launches += 1;
        }
        C->last_transport = PTB_TRANSPORT_NCCL;
    }
    PTB_CU(ctx, cudaEventRecord(ctx->ev1, st));
    if(root) {
        size_t const npix = static_cast<size_t>(W) * H;
        if(rgb_out != nullptr) {
            PTB_CU(ctx, cudaMemcpyAsync(rgb_out, ctx->d_rgb, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        if(rgb8_out != nullptr) {
            PTB_CU(ctx, cudaMemcpyAsync(rgb8_out, ctx->d_rgb8, npix * 3, cudaMemcpyDeviceToHost, st));
        }
    }
    PTB_CU(ctx, cudaStreamSynchronize(st));
    float ms = 0.0f;
    PTB_CU(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.last_resolve_ms = ms;
    if(device_rgb != nullptr) {
        *device_rgb = root ? ctx->d_rgb : nullptr;
    }
    return PTB_OK;
}

} // namespace ptb

// =====================================================================================================================
extern "C" {

int ptb_sample_share(uint32_t total, int n_ranks, int rank, uint32_t* first_out, uint32_t* count_out)
try {
    if(n_ranks < 1 || rank < 0 || rank >= n_ranks || first_out == nullptr || count_out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    ptb::share_of(total, n_ranks, rank, *first_out, *count_out);
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_sample_share")

int ptb_create_multi(int const* devices, int n_devices, ptb_context** out)
try {
    if(out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    *out = nullptr;
    if(devices == nullptr || n_devices < 1 || n_devices > ptb::kMaxGpus) {
        return PTB_ERR_ARGUMENT;
    }
    for(int a = 0; a < n_devices; ++a) {
        for(int b = a + 1; b < n_devices; ++b) {
            if(devices[a] == devices[b]) {
                return PTB_ERR_ARGUMENT; // a GPU listed twice
            }
        }
    }
    auto* G = new ptb::Group{};
    auto* handle = new ptb_context{};
    handle->group = G;
    handle->device = devices[0];
    for(int g = 0; g < n_devices; ++g) {
        ptb_context* m = nullptr;
        int const rc = ptb_create(devices[g], &m);
        if(rc != PTB_OK) {
            ptb_destroy(handle); // ptb_last_error(NULL) still holds ptb_create's message
            return rc;
        }
        G->members.push_back(m);
    }
    // peer mappings both ways between every pair, or none at all
    bool peer = true;
    for(int a = 0; a < n_devices && peer; ++a) {
        for(int b = 0; b < n_devices && peer; ++b) {
            int can = 0;
            if(a != b && (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || can == 0)) {
                peer = false;
            }
        }
    }
    for(int a = 0; a < n_devices && peer; ++a) {
        cudaSetDevice(devices[a]);
        for(int b = 0; b < n_devices && peer; ++b) {
            if(a == b) {
                continue;
            }
            cudaError_t const e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                peer = false;
            }
            (void)cudaGetLastError();
        }
    }
    (void)cudaGetLastError();
    G->peer_ok = peer || n_devices == 1;
    *out = handle;
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_create_multi")

int ptb_comm_unique_id(void* id_out)
try {
    if(id_out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(!ptb::nccl().ok()) {
        return PTB_ERR_STATE;
    }
    ptb::NcclUniqueId id{};
    if(ptb::nccl().GetUniqueId(&id) != 0) {
        return PTB_ERR_CUDA;
    }
    std::memcpy(id_out, &id, sizeof(id));
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_comm_unique_id")

int ptb_comm_init_rank(ptb_context* ctx, void const* id, int n_ranks, int rank)
try {
    using namespace ptb;
    if(ctx == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr || ctx->comm != nullptr) {
        return api_fail(ctx, PTB_ERR_STATE, "ptb_comm_init_rank: this context already spans several GPUs");
    }
    if(id == nullptr || n_ranks < 1 || n_ranks > kMaxGpus || rank < 0 || rank >= n_ranks) {
        return api_fail(ctx, PTB_ERR_ARGUMENT, "ptb_comm_init_rank: need an id, 1 <= n_ranks <= 16 and 0 <= rank < n_ranks");
    }
    if(!nccl().ok()) {
        return api_fail(ctx, PTB_ERR_STATE, nccl().error.c_str());
    }
    PTB_CU(ctx, cudaSetDevice(ctx->device));
    auto* C = new RankComm{};
    C->n = n_ranks;
    C->rank = rank;
    ctx->comm = C;
    auto bail = [&](int rc) {
        ctx->comm = nullptr;
        cudaFree(C->d_xchg);
        cudaFree(C->d_flag);
        cudaFreeHost(C->h_xchg);
        cudaFreeHost(C->h_flag);
        delete C;
        return rc;
    };
    NcclUniqueId uid{};
    std::memcpy(&uid, id, sizeof(uid));
    int const rc = nccl().CommInitRank(&C->comm, n_ranks, uid, rank);
    if(rc != 0) {
        return bail(fail_nccl(ctx, rc, "ncclCommInitRank"));
    }
    cudaError_t e = cudaMalloc(&C->d_xchg, static_cast<size_t>(n_ranks + 1) * sizeof(PeerRecord));
    e = e == cudaSuccess ? cudaMalloc(&C->d_flag, 2 * sizeof(int)) : e;
    e = e == cudaSuccess ? cudaMallocHost(&C->h_xchg, static_cast<size_t>(n_ranks) * sizeof(PeerRecord)) : e;
    e = e == cudaSuccess ? cudaMallocHost(&C->h_flag, sizeof(int)) : e;
    if(e != cudaSuccess) {
        nccl().CommDestroy(C->comm);
        return bail(api_fail_cuda(ctx, e, "ptb_comm_init_rank: buffers"));
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_comm_init_rank")

int ptb_comm_set_transport(ptb_context* ctx, int transport)
try {
    if(ctx == nullptr || transport < PTB_TRANSPORT_AUTO || transport > PTB_TRANSPORT_PEER) {
        return PTB_ERR_ARGUMENT;
    }
    if(ctx->group != nullptr) {
        ctx->group->transport_pref = transport;
    }
    else if(ctx->comm != nullptr) {
        ctx->comm->transport_pref = transport;
    }
    else {
        return ptb::api_fail(ctx, PTB_ERR_STATE, "ptb_comm_set_transport: a single-GPU context sums nothing");
    }
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_comm_set_transport")

int ptb_comm_info(ptb_context* ctx, int32_t out[6])
try {
    if(ctx == nullptr || out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    int32_t v[6] = { 1, 0, 0, 0, 0, 0 };
    if(ctx->group != nullptr) {
        v[0] = static_cast<int32_t>(ctx->group->members.size());
        v[2] = ctx->group->last_transport;
        v[3] = ctx->group->peer_ok ? 1 : 0;
        v[4] = ctx->group->comms.empty() ? 0 : ptb::nccl().version;
        v[5] = 1;
    }
    else if(ctx->comm != nullptr) {
        v[0] = ctx->comm->n;
        v[1] = ctx->comm->rank;
        v[2] = ctx->comm->last_transport;
        v[3] = ctx->comm->mapped ? 1 : 0;
        v[4] = ptb::nccl().version;
        v[5] = 2;
    }
    std::memcpy(out, v, sizeof(v));
    return PTB_OK;
}
PTB_CATCH(ctx, "ptb_comm_info")

} // extern "C"
