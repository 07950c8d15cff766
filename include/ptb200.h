/* ptb200.h -- C ABI of the B200-native per-pixel render loop.
 *
 * The reference (AlexandruIca/cpu-path-tracing) has no plugin/FFI boundary: its
 * hot path is the body of main() between scene construction and PPM output,
 * src/main.cpp:214-236 (one taskflow task per image row, each calling
 * render_subpixel -> camera::get_ray -> radiance -> intersect).  This header IS
 * that boundary, cut so that a host program keeps building pt::scene /
 * pt::sphere / pt::camera exactly as the reference does and hands plain pointers
 * across.  Nothing here mentions torch, CUDA types or C++ types.
 *
 * Record layouts accepted (the reference's own, little-endian, FP64):
 *   sphere  : 88 bytes  radius@0 position@8 emission@32 color@56 reflection:int32@80
 *             = pt::sphere (src/sphere.hpp:10-17) = sandbox Sphere (sandbox/main.cpp:62-66)
 *   camera  : 176 bytes position@0 lower_left_corner@24 cam_x_axis@48 cam_y_axis@72
 *             u@96 v@120 w@144 lens_radius@168 = pt::camera (src/camera.hpp:23-32)
 *   config  : 112 bytes = pt::camera_config (src/camera.hpp:11-21)
 *
 * Conventions: every call returns 0 on success or a negative ptb_status; nothing
 * throws across the boundary; the caller owns every host buffer it passes; the
 * library owns all device memory behind the opaque context.  A context made by
 * ptb_create is bound to ONE GPU (ptb_create_multi: to a list of GPUs) and is not thread-safe (one process / one host thread per GPU, as
 * under torchrun); several contexts may be driven from several threads, renders
 * on the same GPU then take turns.  There is NO CPU fallback: without a usable CUDA device
 * ptb_create fails with PTB_ERR_NO_DEVICE.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 4

#define PTB_SPHERE_BYTES 88
#define PTB_CAMERA_BYTES 176
#define PTB_CAMERA_CONFIG_BYTES 112

typedef struct ptb_context ptb_context;

typedef enum ptb_status
{
    PTB_OK = 0,
    PTB_ERR_ARGUMENT = -1,  /* null pointer, bad size, bad stride, bad flag */
    PTB_ERR_NO_DEVICE = -2, /* no CUDA device / device index out of range */
    PTB_ERR_CUDA = -3,      /* a CUDA runtime call failed; see ptb_last_error */
    PTB_ERR_STATE = -4,     /* call order: scene / camera / image not set yet */
    PTB_ERR_IO = -5,        /* file output failed */
    PTB_ERR_MEMORY = -6,    /* host memory exhausted (std::bad_alloc inside the library; nothing unwinds into the caller) */
    PTB_ERR_INTERNAL = -7   /* any other C++ exception inside the library; see ptb_last_error */
} ptb_status;

/* ptb_render flags */
enum
{
    /* integrator variant */
    PTB_VARIANT_MEGAKERNEL = 0x0, /* one persistent kernel, path regeneration in place */
    PTB_VARIANT_WAVEFRONT = 0x1,  /* queue-based: generate / intersect / shade-by-material */
    PTB_VARIANT_MEGAKERNEL_SORTED = 0x2, /* one persistent kernel; every warp sorts its paths by material through
                                            shared-memory rings so that diffuse_ray / dielectric_ray run on full warps
                                            (PTB_INTEGRATOR_PT only) */
    PTB_VARIANT_MASK = 0xF,
    /* arithmetic */
    PTB_PRECISION_FP32 = 0x00, /* throughput mode (the measured mode) */
    PTB_PRECISION_FP64 = 0x10, /* deterministic parity mode: FP64, no FMA contraction,
                                  reference operation order */
    PTB_PRECISION_MASK = 0xF0,
    /* integrator: which of the reference's two programs */
    PTB_INTEGRATOR_PT = 0x000,      /* src/main.cpp: thin-lens camera, sky, roulette after depth 4, limit 100 */
    PTB_INTEGRATOR_SMALLPT = 0x100, /* sandbox/main.cpp (stand-alone smallpt): tent-filter pinhole camera set with
                                       ptb_set_smallpt_camera, black miss, roulette after depth 5, IOR 1.5 glass
                                       that splits into reflection + refraction while depth <= 2, 2x2 sub-pixels */
    PTB_INTEGRATOR_MASK = 0xF00,
    /* closest-hit query of scenes with more than 64 ordinary spheres (FP32 kernels) */
    PTB_ACCEL_AUTO = 0x0000, /* bounding-volume hierarchy built at upload: same hits, O(log n) per ray */
    PTB_ACCEL_SCAN = 0x1000, /* the reference's linear scan over every sphere (src/main.cpp:30-42) */
    PTB_ACCEL_MASK = 0xF000,
    /* code generation of the FP32 megakernels (sorted and in-place) for scenes of up to 16 spheres */
    PTB_CODEGEN_AUTO = 0x00000,        /* compile the kernel for THIS scene at run time (NVRTC, sphere coefficients as
                                          immediates; cached per scene) when libnvrtc is present, else precompiled */
    PTB_CODEGEN_PRECOMPILED = 0x10000, /* always the precompiled kernel (coefficients from the constant bank) */
    PTB_CODEGEN_MASK = 0xF0000
};

typedef struct ptb_stats
{
    uint64_t paths;          /* camera samples traced since the last ptb_clear (src/main.cpp:184) */
    uint64_t rays;           /* closest-hit queries (calls of intersect, src/main.cpp:30) */
    double last_render_ms;   /* device time of the last ptb_render (CUDA events on its stream) */
    double total_render_ms;  /* sum of the above since the last ptb_clear */
    uint64_t kernel_launches; /* kernels launched by this context since creation */
    uint64_t hits_diffuse;   /* scatter events by material since the last ptb_clear */
    uint64_t hits_specular;
    uint64_t hits_dielectric;
    double last_resolve_ms;  /* device time of the last ptb_resolve* on this context, cross-GPU summation included */
} ptb_stats;

/* ---- library ----------------------------------------------------------------- */
int ptb_abi_version(void);
/* Number of CUDA devices visible (0 if none / no driver). */
int ptb_device_count(void);
/* Message for the last failing call on `ctx` (or of ptb_create when ctx is NULL). */
char const* ptb_last_error(ptb_context const* ctx);

/* ---- context ----------------------------------------------------------------- */
/* Replaces tf::Executor / tf::Taskflow construction, src/main.cpp:214-215. */
int ptb_create(int device, ptb_context** out);
void ptb_destroy(ptb_context* ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the stream a
 * framework is recording events on.  NULL is the CUDA legacy default stream (handle 0),
 * which is what torch.cuda.current_stream().cuda_stream is unless a side stream is set.
 * A fresh context runs on a private non-blocking stream; ptb_reset_stream returns to it. */
int ptb_set_stream(ptb_context* ctx, void* cuda_stream);
int ptb_reset_stream(ptb_context* ctx);
int ptb_synchronize(ptb_context* ctx);

/* ---- inputs ------------------------------------------------------------------ */
/* The sphere list of pt::scene (src/scene.hpp:12-16), in index order (the
 * closest-hit tie rule of src/main.cpp:35 depends on it).  `stride` >= 88.
 * count == 0 is valid (an empty pt::scene: every ray sees the sky, main.cpp:114-120).
 * count < 2^24 (PTB_ERR_ARGUMENT otherwise: list positions travel in 24 bits of the path state).
 * Every number must be finite (PTB_ERR_ARGUMENT otherwise); a radius of 0 is a
 * sphere no ray hits, a negative radius counts as its magnitude (the reference
 * only squares it, src/sphere.cpp:11).
 * More than 64 ordinary-sized spheres: a bounding-volume hierarchy is built here
 * (host side, milliseconds) for the FP32 kernels -- see PTB_ACCEL_*. */
int ptb_upload_scene(ptb_context* ctx, void const* spheres, size_t count, size_t stride);
/* A derived pt::camera, i.e. the result of pt::camera::with_config (src/main.cpp:209). */
int ptb_set_camera(ptb_context* ctx, void const* camera, size_t bytes);
/* Camera of the stand-alone smallpt fork, sandbox/main.cpp:235-237,260: cam8 = position(3), viewing
 * direction(3, need not be unit), field-of-view factor (.5135 there), push distance (140 there).
 * cx, cy are derived from the image size like the sandbox does. */
int ptb_set_smallpt_camera(ptb_context* ctx, double const* cam8);
/* Image geometry: width/height of src/main.cpp:204-205, num_subpixels of :202.
 * (Re)allocates and zeroes the accumulation buffer: one float4 {r,g,b,n} of
 * un-clamped radiance SUMS per sub-pixel (slot = ((y*W+x)*ns+sy)*ns+sx). */
int ptb_set_image(ptb_context* ctx, int width, int height, int num_subpixels);
/* Zero the accumulation buffer and the statistics. */
int ptb_clear(ptb_context* ctx);

/* ---- the hot path ---------------------------------------------------------------- */
/* Replaces executor.run(taskflow).wait(), src/main.cpp:217-236: traces samples
 * [first_sample, first_sample + samples_per_subpixel) of EVERY sub-pixel (the `s`
 * loop of src/main.cpp:184) and adds their radiance into the accumulation buffer.
 * Progressive: call again with the next sample range to refine; ranks of a
 * multi-GPU job call it with disjoint ranges.  Blocks until the device is done.
 * `seed` keys the counter-based stream (oracle/ptb_rng.h) that replaces the
 * reference's unseedable per-row mt19937 (src/random_state.cpp:3-7). */
int ptb_render(ptb_context* ctx, uint64_t seed, uint32_t first_sample, uint32_t samples_per_subpixel, uint32_t flags);

/* Per-sub-pixel mean -> clamp to [0,1] -> average of the ns*ns strata, written at
 * the vertically flipped row: src/main.cpp:181,192-196.  rgb_out: width*height*3
 * doubles, the layout of the reference's std::vector<pt::vec3> image (main.cpp:210). */
int ptb_resolve(ptb_context* ctx, double* rgb_out);
/* Same, straight to 8-bit gamma-2.2 values as pt::color_to_int (src/utils.cpp:11-16):
 * width*height*3 bytes. */
int ptb_resolve_rgb8(ptb_context* ctx, uint8_t* rgb8_out);

/* Resolve on the device only: *device_rgb receives the address of width*height*3 doubles
 * owned by the context (valid until the next ptb_set_image / ptb_destroy).  No copy. */
int ptb_resolve_device(ptb_context* ctx, void** device_rgb);

/* ---- several GPUs behind the same calls ---------------------------------------------------------------
 * The reference fills the whole image with ONE blocking call on ONE machine (src/main.cpp:214-236, a taskflow
 * task per row).  Here the work is split over GPUs by SAMPLES: member g of G traces the contiguous share
 * ptb_sample_share(S, G, g) of the samples [first_sample, first_sample + S) of every sub-pixel into its own
 * accumulation buffer -- the stream is keyed by the absolute sample index, so the image does not depend on G --
 * and the un-clamped per-stratum sums (and their counts, in .w) are summed across GPUs before the non-linear
 * resolve (main.cpp:192-196).  Two ways to get there, same entry points afterwards:
 *
 *  (1) ptb_create_multi: ONE host process drives n GPUs -- what a patched src/main.cpp uses.
 *      Every call on the handle applies to all members: ptb_upload_scene / ptb_set_camera / ptb_set_image /
 *      ptb_clear broadcast, ptb_render runs the members' shares concurrently (one host thread per GPU) and
 *      blocks until all are done, ptb_resolve* sum + resolve and fill the caller's image.
 *  (2) ptb_comm_init_rank: ONE process per GPU (torchrun, mpirun).  Rank 0 makes an id with
 *      ptb_comm_unique_id and hands it to the others by whatever side channel the launcher has; after
 *      ptb_comm_init_rank every rank makes the same calls: ptb_render traces this rank's share, ptb_resolve* are
 *      COLLECTIVE (every rank must call; rank 0 receives the image, the others may pass NULL).
 *
 * How the sums travel (PTB_TRANSPORT_*): AUTO picks PEER when every GPU can map the others' memory (NVLink /
 * NVSwitch; CUDA IPC between processes), else NCCL.
 *   PEER  one kernel per GPU does reduce-scatter + resolve + gather at once: GPU g reads rows [g*H/G, (g+1)*H/G)
 *         of ALL members' accumulation buffers through peer loads, resolves them and stores the pixels straight into
 *         the root's image.  Nothing is written but the image; accumulation buffers stay as they are (progressive).
 *   NCCL  ncclReduce(sum) of the buffers into a scratch buffer on the root (library-owned communicator, libnccl.so.2
 *         loaded at run time), then the single-GPU resolve there.
 * ptb_stats.last_resolve_ms is the device time of either, ptb_comm_info says which one ran. */
int ptb_create_multi(int const* devices, int n_devices, ptb_context** out);

#define PTB_COMM_ID_BYTES 128
enum
{
    PTB_TRANSPORT_AUTO = 0,
    PTB_TRANSPORT_NCCL = 1,
    PTB_TRANSPORT_PEER = 2
};
/* 128 opaque bytes (an ncclUniqueId) made on rank 0; needs libnccl.so.2 (PTB_ERR_STATE when it cannot be loaded). */
int ptb_comm_unique_id(void* id_out);
/* Join a job of n_ranks processes as `rank`; collective (blocks until every rank has called it). */
int ptb_comm_init_rank(ptb_context* ctx, void const* id, int n_ranks, int rank);
/* PTB_TRANSPORT_*: a request; PEER falls back to NCCL where peer mapping is impossible.  Must be the same on all ranks. */
int ptb_comm_set_transport(ptb_context* ctx, int transport);
/* out = { number of GPUs (1 for a plain context), this context's rank (0 for a ptb_create_multi handle),
 *         transport of the last resolve (PTB_TRANSPORT_NCCL / _PEER, 0 = none yet), 1 if peer mapping is available,
 *         NCCL version code (0 = not loaded), 1 for a single-process group / 2 for one process per GPU / 0 }. */
int ptb_comm_info(ptb_context* ctx, int32_t out[6]);
/* The share of `total` samples that member `rank` of `n_ranks` traces: contiguous, sizes differ by at most one. */
int ptb_sample_share(uint32_t total, int n_ranks, int rank, uint32_t* first_out, uint32_t* count_out);

/* ---- caller-side plumbing (single-GPU contexts only) ----------------------------------------------------- */
/* Device address and size of the FP32 accumulation buffer, for a caller that runs its own collective layer
 * (e.g. torch.distributed) instead of the calls above.  Slots carry their own sample count in .w, so
 * a plain sum-reduce is all that is needed. */
int ptb_accum_buffer(ptb_context* ctx, void** device_ptr, size_t* bytes);
/* Render into / resolve from a caller-owned device buffer (e.g. a torch tensor)
 * of at least width*height*ns*ns*16 bytes.  NULL returns to the internal one. */
int ptb_set_accum_buffer(ptb_context* ctx, void* device_ptr, size_t bytes);
/* Host copy of the accumulation buffer (float4 per slot). */
int ptb_download_accum(ptb_context* ctx, float* out, size_t floats);

/* ---- checkpoint / resume (the reference's own TODO, README.md:9: "save progress to resume") ------ */
/* ptb_download_accum IS the checkpoint of a progressive FP32 render: un-clamped per-stratum sums and, in .w, how
 * many samples each slot holds.  ptb_upload_accum restores it into a context whose image geometry matches
 * (floats == width*height*ns*ns*4): the buffer is REPLACED, and rendering continues with
 * ptb_render(seed, first_sample = samples already in the checkpoint, ...) -- the stream is keyed by the absolute
 * sample index, so save -> destroy -> create -> restore -> continue gives the slots an uninterrupted run gives.
 * The 64-bit pair does the same for PTB_PRECISION_FP64 renders (4 doubles per slot), bit for bit. */
int ptb_upload_accum(ptb_context* ctx, float const* in, size_t floats);
int ptb_download_accum64(ptb_context* ctx, double* out, size_t doubles);
int ptb_upload_accum64(ptb_context* ctx, double const* in, size_t doubles);

/* ---- introspection ------------------------------------------------------------------ */
int ptb_get_stats(ptb_context* ctx, ptb_stats* out);

/* How the FP32 path packed the uploaded scene (ptb_scene.cuh): out[0..9] = small near-only,
 * small both-roots, big near-only, big both-roots, of the big near-only: on the x / y / z
 * axis of the frame, bit 0: big spheres share one radius | bit 1: index-in-key allowed (scene small
 * against epsilon), lists fit constant memory (0/1),
 * a fully unrolled kernel exists for this layout (0/1). */
int ptb_scene_layout(ptb_context* ctx, int32_t out[10]);
/* Run-time code generation (PTB_CODEGEN_AUTO): out = { available (libnvrtc + libcuda found, PTB_JIT != 0), kernels
 * compiled so far, failed attempts, 1 if the last FP32 megakernel launch ran a run-time compiled kernel,
 * total compile time in ms }.  A failed attempt is not an error of ptb_render: the precompiled kernel runs instead;
 * ptb_jit_last_error says what happened (empty string when nothing did). */
int ptb_jit_info(ptb_context* ctx, int32_t out[5]);
char const* ptb_jit_last_error(ptb_context* ctx);

/* Parity probe: trace `count` individual samples (x, y, sx, sy, sample index in
 * reference loop coordinates) and return, per sample, the sphere index hit by the
 * camera ray (-1 = miss) and the radiance estimate.  PTB_PRECISION_FP64 gives the
 * deterministic mode; PTB_PRECISION_FP32 runs the throughput kernels' arithmetic.
 * radiance_out: count*3 doubles.  ray_out (optional): count*6 doubles origin+direction.
 * draws_out (optional): random draws consumed per sample. */
int ptb_trace_samples(ptb_context* ctx, uint64_t seed, uint32_t const* x, uint32_t const* y, uint32_t const* sx,
                      uint32_t const* sy, uint32_t const* sample, size_t count, uint32_t flags, int32_t* primary_hit_out,
                      double* radiance_out, double* ray_out, uint32_t* draws_out);

/* Parity probe, second part: which sphere each of the `count` samples' paths hits at depth 0, 1, ... trail_len-1
 * (FP64 deterministic mode, src/ integrator): trail_out[count*trail_len], -1 = the ray left to the sky, -2 = the path
 * had ended before.  For classifying samples whose radiance differs from the oracle's: a path is bit-exact as long as
 * it only met mirrors (+ - * / sqrt); after a diffuse or glass bounce it carries libm-vs-CUDA sin/cos/pow ulps. */
int ptb_trace_paths(ptb_context* ctx, uint64_t seed, uint32_t const* x, uint32_t const* y, uint32_t const* sx,
                    uint32_t const* sy, uint32_t const* sample, size_t count, int trail_len, int32_t* trail_out);

/* Raw uniforms of the counter stream for (seed, slot, sample): draws_out[count*n_draws]
 * as doubles, generated ON THE DEVICE.  For checking the stream against oracle/ptb_rng.h. */
int ptb_rng_draws(ptb_context* ctx, uint64_t seed, uint32_t const* slot, uint32_t const* sample, size_t count,
                  int n_draws, double* draws_out);

/* Roofline calibration: runs a dependent-free FFMA loop on every SM of the context's GPU
 * and reports the achieved FP32 rate in TFLOP/s (FFMA = 2 flop).  This is the measured
 * denominator of the FP32-issue roofline the render loop is bound by. */
int ptb_measure_fp32_peak(ptb_context* ctx, double* tflops_out);

/* ---- host-side helpers mirroring the reference's scene layer (no GPU needed) ----------------- */
/* pt::camera::with_config (src/camera.cpp:3-17): 112-byte config -> 176-byte camera. */
int ptb_camera_with_config(void const* camera_config, void* camera_out);
/* Built-in scenes: "simple" (src/simple_scene.hpp:14-52), "box" (src/box_scene.hpp:14-72),
 * "box_mirror" (src/box_mirror_scene.hpp:14-72), "dof_glass" (BASELINE config 4),
 * "spheres10k" (config 5).  Writes up to `capacity` 88-byte spheres, the
 * count, and the 112-byte camera_config.  Call with spheres_out = NULL to query the count. */
int ptb_builtin_scene(char const* name, int width, int height, void* spheres_out, size_t capacity, size_t* count_out,
                      void* camera_config_out);
/* The sandbox's own scene and camera constants (sandbox/main.cpp:94-122,235,260): up to `capacity`
 * 88-byte spheres (10 of them), the count, and cam8 for ptb_set_smallpt_camera. */
int ptb_builtin_smallpt_scene(void* spheres_out, size_t capacity, size_t* count_out, double* cam8_out);
/* ASCII PPM "P3" writer with gamma 2.2, byte-compatible with src/main.cpp:240-247. */
int ptb_write_ppm(char const* path, double const* rgb, int width, int height);
/* Same, with the sandbox's rounding: int(pow(clamp(x), 1/2.2) * 255 + .5), sandbox/main.cpp:130-133,271-275. */
int ptb_write_ppm_smallpt(char const* path, double const* rgb, int width, int height);
/* The output stage for large images (src/main.cpp:240-247 spends its time in 3*W*H pow() calls and integer
 * formatting): write the 8-bit values ptb_resolve_rgb8 produced on the GPU.  binary != 0: "P6" (raw bytes, 3*W*H);
 * binary == 0: the reference's ASCII "P3" layout, token for token, through a 256-entry table. */
int ptb_write_ppm_rgb8(char const* path, uint8_t const* rgb8, int width, int height, int binary);

#ifdef __cplusplus
}
#endif

#endif /* PTB200_H */
