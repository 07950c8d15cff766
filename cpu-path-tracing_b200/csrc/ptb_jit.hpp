// ptb_jit.hpp -- scene-specialised code at run time (internal to libptb200.so).
//
// The sorted megakernel reads a scene's sphere coefficients as constant-bank operands: 7 uniform loads per
// closest-hit scan on the box scenes, and the unrolled tests cannot fold anything.  With the coefficients as
// LITERALS the same source compiles to FFMA / FADD with immediates: measured 4.1 % faster on box_mirror (a hand-made
// literal build, bit-identical paths), while keeping them in registers instead is 3 % slower.
// Scenes arrive through the C ABI with arbitrary numbers, so the literal build has to be made when the scene is
// known: NVRTC compiles ptb_mega_sorted.cuh -- the very source of the precompiled kernels, embedded in the library
// at build time -- with the packed coefficients injected as a hex-float initialiser, for sm_100a, once per distinct
// (layout, coefficients, in-place material); the cubin is loaded through the driver API and cached in the context.
//
// Nothing here is required: libnvrtc / libcuda are opened with dlopen, and every failure (library absent, compile
// error, load error) falls back to the precompiled constant-bank kernel and is reported through ptb_jit_info.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "ptb_kernels.h"

namespace ptb {

struct JitKernel
{
    void* module = nullptr;          // CUmodule
    void* function = nullptr;        // CUfunction
    int blocks_per_sm = 0;
    bool failed = false;             // negative cache entry: do not try this key again
    bool pending = false;            // seen once, not compiled yet
    unsigned long long last_use = 0;
};

class JitCache
{
public:
    JitCache() = default;
    JitCache(JitCache const&) = delete;
    JitCache& operator=(JitCache const&) = delete;
    ~JitCache();

    // The specialised kernel for this packed scene; nullptr = use the precompiled kernel.  Compiling costs ~0.2-0.4 s,
    // so it has to be worth it: `eager` (the caller is about to trace enough paths to amortise it) compiles at once,
    // otherwise a scene is compiled the SECOND time it is asked for -- geometry that changes every frame never is.
    // At most kMaxModules stay loaded (least recently used goes first).
    static constexpr size_t kMaxModules = 32;
    // kind: which kernel -- the sorted megakernel scattering `inline_material` in place, or the in-place megakernel
    // with the src/main.cpp or the sandbox integrator
    enum Kind { kSorted = 0, kInPlacePt = 1, kInPlaceSmallpt = 2 };
    JitKernel const* get(ConstSceneF32 const& cs, SceneCounts const& counts, Kind kind, int inline_material, bool eager);
    // Launch it: same grid policy and semantics as launch_megakernel_sorted.
    cudaError_t launch(JitKernel const& k, RenderParamsF32 const& p, int sm_count, cudaStream_t stream, int* launches);

    bool available();                 // libnvrtc + libcuda found and not disabled (PTB_JIT=0)
    int compiled() const { return compiled_; }
    int failures() const { return failures_; }
    double compile_ms() const { return compile_ms_; }
    std::string const& last_error() const { return error_; }

    // The translation unit handed to NVRTC for a key (also what tests/ compile on the CPU to check the plumbing).
    static std::string translation_unit(ConstSceneF32 const& cs, SceneCounts const& counts, Kind kind);
    static std::string kernel_name(SceneCounts const& counts, Kind kind, int inline_material);

private:
    std::map<std::vector<uint32_t>, JitKernel> cache_;
    int compiled_ = 0, failures_ = 0;
    double compile_ms_ = 0.0;
    std::string error_;
    int state_ = 0; // 0 unknown, 1 available, -1 unavailable
    unsigned long long clock_ = 0;
};

} // namespace ptb
