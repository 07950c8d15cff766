// ptb_kernels.h -- host-callable launchers of the CUDA kernels (C++ linkage,
// internal to libptb200.so; the public surface is include/ptb200.h).
#pragma once

#include "ptb_types.h"

#include "ptb_scene.cuh"

namespace ptb {

// ---- device-resident counters -------------------------------------------------------------
struct DeviceCounters
{
    unsigned long long tile_cursor; // work distribution of the persistent megakernel
    unsigned long long rays;
    unsigned long long diffuse;
    unsigned long long specular;
    unsigned long long dielectric;
    unsigned long long paths;
};

// ---- FP32 throughput path --------------------------------------------------------------------
struct RenderParamsF32
{
    uint64_t key;           // seed_key(seed)
    uint32_t first_sample;  // absolute index of the first sample of this pass
    uint32_t samples;       // samples per sub-pixel in this pass
    uint32_t chunk;         // samples per work tile
    uint32_t width, height, ns;
    uint32_t nslots;        // width*height*ns*ns
    uint32_t ngroups;       // ceil(nslots/32)
    uint32_t nchunks;       // ceil(samples/chunk)
    uint32_t ntiles;        // ngroups*nchunks
    float4* accum;          // [nslots] {sum r, sum g, sum b, n}
    DeviceCounters* counters;
    ShadePlanes shade;      // global-memory shading planes
    GeoLists geo;           // global-memory geometry lists (generic variant)
    int n_total;
    uint32_t key_mask;      // ~(2^kIdBits - 1), handed over as DATA so that it lives in a register (see closest_hit)
    CameraPair cams;        // thin-lens camera of src/main.cpp and the sandbox's pinhole camera, shifted FP32 frame
};

#ifndef __CUDACC_RTC__ // everything below is host-side; the run-time compiler only needs the records above

// Copy the packed scene into the constant bank the FP32 kernels read.
cudaError_t upload_const_scene(ConstSceneF32 const& cs, cudaStream_t stream);

// Lengths of the four geometry lists of a packed scene (ptb_scene.cuh)
struct SceneCounts
{
    int small_near = 0, small_both = 0, big_near = 0, big_both = 0;
    bool fits_const = true; // both classes fit the __constant__ lists
    int big_x = 0, big_y = 0, big_z = 0; // prefix of the near-only big list: centres on a frame axis
    bool uniform_k = false;              // every big sphere has the same radius
    bool embed_ok = true;                // scene small enough against epsilon for index-in-key (ptb_path_f32.cuh)
    int pair_mask = 0;                   // bit a: the two spheres of axis group a are mirror images (+side listed first)
};
// True when a fully unrolled kernel exists for these list lengths.
bool megakernel_has_specialisation(SceneCounts const& c);
// Launch the persistent megakernel: grid = SM count * resident blocks.
cudaError_t launch_megakernel(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream,
                              int* launches, bool smallpt);

// The material-sorted megakernel (ptb_mega_sorted.cuh): src/main.cpp integrator only.
cudaError_t launch_megakernel_sorted(RenderParamsF32 const& p, SceneCounts const& c, int sm_count, cudaStream_t stream,
                                     int* launches, int inline_material);

// ---- wavefront / material-sorted variant (ptb_wavefront.cuh) ------------------------------------------
struct WavefrontCounters
{
    uint32_t n_active[2];
    uint32_t n_mat[2];
    uint32_t regen_base;
    uint32_t regen_count;
    uint32_t iterations;
    uint32_t pad_;
    unsigned long long cursor;      // next (sample, slot) item
    unsigned long long regen_item0;
};
constexpr int kWfPlanesPerPool = 18; // 2 active streams x 4 planes + 2 material streams x 5 planes
constexpr int kWfWordsPerPool = 4;   // one slot array per stream
struct WavefrontBuffers
{
    float4* planes;   // kWfPlanesPerPool * pool float4
    uint32_t* words;  // kWfWordsPerPool * pool words
    WavefrontCounters* counters;
    uint32_t pool;
};
cudaError_t launch_wavefront(WavefrontBuffers const& buf, RenderParamsF32 const& p, SceneCounts const& c, int sm_count,
                             cudaStream_t stream, int* launches);

struct ProbeParams
{
    uint64_t key;
    uint32_t width, height, ns;
    uint32_t const* x;
    uint32_t const* y;
    uint32_t const* sx;
    uint32_t const* sy;
    uint32_t const* sample;
    uint32_t count;
    int32_t* primary_hit;
    double* radiance; // [count*3]
    double* ray;      // [count*6] or nullptr
    uint32_t* draws;  // [count] or nullptr
    CameraPair cams;  // FP32 probe only
};
cudaError_t launch_probe_f32(ProbeParams const& p, SceneCounts const& c, ShadePlanes const& shade, GeoLists const& geo,
                             cudaStream_t stream, bool smallpt);

// ---- stand-alone smallpt fork, FP64 parity (ptb_smallpt_f64.cu); cam8 is device memory -----------------------------
cudaError_t launch_smallpt_probe_f64(ProbeParams const& p, RawSphere const* spheres, int n, double const* cam8,
                                     cudaStream_t stream);
cudaError_t launch_smallpt_render_f64(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width, uint32_t height,
                                      RawSphere const* spheres, int n, double const* cam8, double* accum64,
                                      DeviceCounters* counters, cudaStream_t stream);

// ---- FP64 parity path (reference operation order, no FMA contraction) ----------------------------------
cudaError_t launch_probe_f64(ProbeParams const& p, RawSphere const* spheres, int n, RawCamera const* cam,
                             cudaStream_t stream);
cudaError_t launch_trail_f64(ProbeParams const& p, RawSphere const* spheres, int n, RawCamera const* cam, int32_t* trail,
                             int trail_len, cudaStream_t stream);
cudaError_t launch_render_f64(uint64_t key, uint32_t first_sample, uint32_t samples, uint32_t width, uint32_t height,
                              uint32_t ns, RawSphere const* spheres, int n, RawCamera const* cam, double* accum64,
                              DeviceCounters* counters, cudaStream_t stream);

// ---- resolve (main.cpp:181,192-196) -----------------------------------------------------------------------
// accum32: float4 per slot; accum64: 4 doubles per slot (either may be null, both are summed when present)
cudaError_t launch_resolve(float4 const* accum32, double const* accum64, uint32_t width, uint32_t height, uint32_t ns,
                           double* rgb_out, uint8_t* rgb8_out, cudaStream_t stream);
// The same kernel over the buffers of SEVERAL GPUs (peer mappings) and a range of pixel rows [y0, y1) in the reference's
// row numbering (before the flip): G such launches, one per GPU, sum + resolve + gather the image (ptb_multi.cpp).
constexpr int kMaxGpus = 16;
struct ResolveSources
{
    float4 const* accum32[kMaxGpus];
    double const* accum64[kMaxGpus];
    int n;
};
cudaError_t launch_resolve_rows(ResolveSources const& src, uint32_t width, uint32_t height, uint32_t ns, uint32_t y0,
                                uint32_t y1, double* rgb_out, uint8_t* rgb8_out, cudaStream_t stream);

// ---- FP32 peak calibration (FFMA loop) ------------------------------------------------------------------------
cudaError_t launch_fp32_peak(int sm_count, int iters, float* scratch, cudaStream_t stream, double* flop_out);

// ---- stream check ---------------------------------------------------------------------------------------------
cudaError_t launch_rng_draws(uint64_t key, uint32_t const* slot, uint32_t const* sample, uint32_t count, int n_draws,
                             double* out, cudaStream_t stream);

#endif // __CUDACC_RTC__

} // namespace ptb
