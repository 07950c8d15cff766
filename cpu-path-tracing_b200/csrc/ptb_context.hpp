// ptb_context.hpp -- the context behind include/ptb200.h's opaque handle.  Internal to libptb200.so:
// ptb_api.cpp (single-GPU entry points) and ptb_multi.cpp (several GPUs behind the same entry points).
#pragma once

#include "../../include/ptb200.h"
#include "ptb_jit.hpp"
#include "ptb_kernels.h"

#include <string>
#include <vector>

namespace ptb {
struct Group;
struct RankComm;
} // namespace ptb

using namespace ptb;

struct ptb_context
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr; // the one in use (own or caller's)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    // scene
    int n = 0;
    bool have_scene = false, have_camera = false;
    std::vector<RawSphere> h_spheres;
    RawCamera h_camera{};
    RawSphere* d_spheres = nullptr;
    size_t d_spheres_cap = 0;
    RawCamera* d_camera = nullptr;
    double sb_cam8[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // smallpt camera: position, direction, fov factor, push
    bool have_sbcam = false;
    double* d_sbcam8 = nullptr;
    ConstSceneF32 cs{};
    CameraPair cams{}; // both cameras in the shifted FP32 frame: a kernel argument, not part of the constant block
    double shift[3] = { 0, 0, 0 };
    SceneCounts counts{};
    SmallGeo* d_small = nullptr;
    BigGeo* d_big = nullptr;
    int* d_order = nullptr;
    float4* d_shade = nullptr; // 4 planes of n
    size_t geo_cap = 0;
    // bounding-volume hierarchy over the small spheres (ptb_bvh.hpp); built when they do not fit the constant lists
    float4* d_bvh_nodes = nullptr;
    SmallGeo* d_bvh_geo = nullptr;
    int* d_bvh_pos = nullptr;
    size_t bvh_node_cap = 0, bvh_leaf_cap = 0;
    int bvh_root = 0;
    bool have_bvh = false;
    int bvh_depth = 0, bvh_node_count = 0;

    // image
    int width = 0, height = 0, ns = 0;
    size_t nslots = 0;
    float4* d_accum = nullptr; // owned
    size_t d_accum_bytes = 0;
    float4* ext_accum = nullptr; // caller-owned
    size_t ext_accum_bytes = 0;
    double* d_accum64 = nullptr;
    bool accum64_used = false;
    double* d_rgb = nullptr;
    uint8_t* d_rgb8 = nullptr;

    DeviceCounters* d_counters = nullptr;
    // material scattered in place by the sorted megakernel (0 diffuse, 1 specular): a guess from the geometry at
    // upload, then whichever of the two the previous sorted launch hit more often
    int inline_material = 1;
    unsigned long long seen_diffuse = 0, seen_specular = 0; // counter values already accounted for
    JitCache jit;                 // run-time compiled, scene-specialised sorted megakernels (ptb_jit.hpp)
    bool last_launch_jit = false;
    WavefrontBuffers wf{ nullptr, nullptr, nullptr, 0 };
    ptb_stats stats{};
    uint64_t buffers_epoch = 0; // bumped whenever d_accum / d_accum64 / d_rgb / d_rgb8 are (re)allocated: peers re-map

    // multi-GPU (ptb_multi.cpp).  group: this handle stands for several member contexts in ONE process
    // (ptb_create_multi) and owns no device memory itself.  comm: this context is one rank of a job with one process
    // per GPU (ptb_comm_init_rank).
    ptb::Group* group = nullptr;
    ptb::RankComm* comm = nullptr;
};

namespace ptb {

// ---- ptb_multi.cpp: the same entry points when the handle stands for several GPUs -----------------------------------
// Each returns a ptb_status; the message goes to the handle's err.
void multi_destroy(ptb_context* ctx);
int multi_synchronize(ptb_context* ctx);
int multi_upload_scene(ptb_context* ctx, void const* spheres, size_t count, size_t stride);
int multi_set_camera(ptb_context* ctx, void const* camera, size_t bytes);
int multi_set_smallpt_camera(ptb_context* ctx, double const* cam8);
int multi_set_image(ptb_context* ctx, int width, int height, int ns);
int multi_clear(ptb_context* ctx);
int multi_render(ptb_context* ctx, uint64_t seed, uint32_t first_sample, uint32_t samples, uint32_t flags);
int multi_resolve(ptb_context* ctx, double* rgb_out, uint8_t* rgb8_out, void** device_rgb);
int multi_download_accum(ptb_context* ctx, float* out, size_t floats);
int multi_upload_accum(ptb_context* ctx, float const* in, size_t floats);
int multi_download_accum64(ptb_context* ctx, double* out, size_t doubles);
int multi_upload_accum64(ptb_context* ctx, double const* in, size_t doubles);
int multi_get_stats(ptb_context* ctx, ptb_stats* out);
ptb_context* multi_root(ptb_context* ctx); // the member that owns the image (device list position 0)

// one process per GPU: what changes for a context that is a rank of a job
void comm_destroy(ptb_context* ctx);
int comm_release_peers(ptb_context* ctx); // collective: before buffers other ranks may have mapped are freed
int comm_resolve(ptb_context* ctx, double* rgb_out, uint8_t* rgb8_out, void** device_rgb);

// ---- ptb_api.cpp helpers the multi-GPU layer needs -----------------------------------------------------------------------
int api_zero_accum(ptb_context* ctx); // accumulation buffers only: statistics and counters stay
float4* api_active_accum(ptb_context* ctx);
int api_fail(ptb_context* ctx, int code, char const* what);
// What every `extern "C"` entry point does with a C++ exception (std::bad_alloc from a huge upload, std::system_error from
// a thread that cannot start, ...): nothing may unwind into a caller that is C, Go or ctypes.  `ctx` may be null.
int api_exception(ptb_context* ctx, char const* entry) noexcept;
int api_fail_cuda(ptb_context* ctx, cudaError_t e, char const* what);

} // namespace ptb

// Function-try-block tail of an entry point returning a status: `int ptb_x(...) try { ... } PTB_CATCH(ctx, "ptb_x")`
#define PTB_CATCH(ctx, entry) \
    catch(...) \
    { \
        return ptb::api_exception((ctx), (entry)); \
    }
