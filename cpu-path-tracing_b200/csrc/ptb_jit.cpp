// ptb_jit.cpp -- see ptb_jit.hpp.  Host code only; NVRTC and the driver API are reached through dlopen.
#include "ptb_jit.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

#include <cuda.h>
#include <nvrtc.h>

// build/ptb_jit_sources.inc (Makefile): kJitHeaderNames[], kJitHeaderSources[], kJitHeaderCount -- the device
// headers as they were when the library was built
#ifndef PTB_TOOLKIT_LIBDIR
#define PTB_TOOLKIT_LIBDIR "/usr/local/cuda/lib64"
#endif

#include "ptb_jit_sources.inc"

namespace ptb {

namespace {

struct Api
{
    void* nvrtc = nullptr;
    void* cuda = nullptr;
    // NVRTC
    nvrtcResult (*CreateProgram)(nvrtcProgram*, char const*, char const*, int, char const* const*, char const* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, char const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, char const* const*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, char const*, char const**) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    // driver
    CUresult (*ModuleLoadData)(CUmodule*, void const*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, char const*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                             void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, char const**) = nullptr;
    bool ok = false;
};

template<class F>
bool sym(void* lib, char const* name, F& out)
{
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}

Api& api()
{
    static Api a = [] {
        Api x;
        char const* off = std::getenv("PTB_JIT");
        if(off != nullptr && std::strcmp(off, "0") == 0) {
            return x;
        }
        // Which NVRTC: PTB_NVRTC_LIB, else the one of the toolkit this library was built with (by PATH, so that it is that
        // file even when the process already holds another libnvrtc.so.12 -- a Python process that imported torch has torch's
        // bundled 12.8, whose code for this kernel measured 2.8 % slower than 12.9's), else whatever the name resolves to.
        char const* const env_lib = std::getenv("PTB_NVRTC_LIB");
        for(char const* n : { env_lib, PTB_TOOLKIT_LIBDIR "/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12" }) {
            if(n == nullptr || *n == '\0') {
                continue;
            }
            x.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if(x.nvrtc != nullptr) {
                break;
            }
        }
        for(char const* n : { "libcuda.so.1", "libcuda.so" }) {
            x.cuda = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if(x.cuda != nullptr) {
                break;
            }
        }
        if(x.nvrtc == nullptr || x.cuda == nullptr) {
            return x;
        }
        bool ok = sym(x.nvrtc, "nvrtcCreateProgram", x.CreateProgram) && sym(x.nvrtc, "nvrtcDestroyProgram", x.DestroyProgram) &&
                  sym(x.nvrtc, "nvrtcAddNameExpression", x.AddNameExpression) && sym(x.nvrtc, "nvrtcCompileProgram", x.CompileProgram) &&
                  sym(x.nvrtc, "nvrtcGetLoweredName", x.GetLoweredName) && sym(x.nvrtc, "nvrtcGetCUBINSize", x.GetCUBINSize) &&
                  sym(x.nvrtc, "nvrtcGetCUBIN", x.GetCUBIN) && sym(x.nvrtc, "nvrtcGetProgramLogSize", x.GetProgramLogSize) &&
                  sym(x.nvrtc, "nvrtcGetProgramLog", x.GetProgramLog);
        ok = ok && sym(x.cuda, "cuModuleLoadData", x.ModuleLoadData) && sym(x.cuda, "cuModuleUnload", x.ModuleUnload) &&
             sym(x.cuda, "cuModuleGetFunction", x.ModuleGetFunction) && sym(x.cuda, "cuLaunchKernel", x.LaunchKernel) &&
             sym(x.cuda, "cuOccupancyMaxActiveBlocksPerMultiprocessor", x.OccupancyMaxActiveBlocks) &&
             sym(x.cuda, "cuGetErrorString", x.GetErrorString);
        x.ok = ok;
        return x;
    }();
    return a;
}

void put_float(std::string& s, float v)
{
    char buf[48];
    std::snprintf(buf, sizeof(buf), "%af", static_cast<double>(v)); // hex float: exact round trip (C++17 literal)
    s += buf;
}

int small_count(SceneCounts const& c)
{
    return c.small_near + c.small_both;
}
int big_count(SceneCounts const& c)
{
    return c.big_near + c.big_both;
}

} // namespace

std::string JitCache::kernel_name(SceneCounts const& c, Kind kind, int inline_material)
{
    char shape[160];
    std::snprintf(shape, sizeof(shape), "ptb::SceneShape<%d, %d, %d, %d, %d, %d, %d, %s, %s, %d>", c.small_near, c.small_both, c.big_near,
                  c.big_both, c.big_x, c.big_y, c.big_z, c.uniform_k ? "true" : "false", c.embed_ok ? "true" : "false", c.pair_mask);
    char buf[320];
    if(kind == kSorted) {
        std::snprintf(buf, sizeof(buf), "ptb::mega_sorted_kernel<%s, true, %d>", shape, inline_material < 0 ? -1 : inline_material == 0 ? 0 : 1);
    }
    else {
        std::snprintf(buf, sizeof(buf), "ptb::mega_kernel<%s, true, ptb::%s>", shape, kind == kInPlacePt ? "IntegratorPt" : "IntegratorSmallpt");
    }
    return buf;
}

std::string JitCache::translation_unit(ConstSceneF32 const& cs, SceneCounts const& c, Kind kind)
{
    int const ns = small_count(c), nb = big_count(c);
    std::string init = "{ { ";
    for(int i = 0; i < (ns > 0 ? ns : 1); ++i) {
        SmallGeo const g = i < ns ? cs.small_geo[i] : SmallGeo{ 0.0f, 0.0f, 0.0f, 0.0f };
        init += "{ ";
        put_float(init, g.cx);
        init += ", ";
        put_float(init, g.cy);
        init += ", ";
        put_float(init, g.cz);
        init += ", ";
        put_float(init, g.r2);
        init += " }, ";
    }
    init += "}, { ";
    for(int i = 0; i < (nb > 0 ? nb : 1); ++i) {
        BigGeo const g = i < nb ? cs.big_geo[i] : BigGeo{ 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
        float const f[8] = { g.gx, g.gy, g.gz, g.k, g.K, g.two_r, 0.0f, 0.0f };
        init += "{ ";
        for(int k = 0; k < 8; ++k) {
            put_float(init, f[k]);
            init += k < 7 ? ", " : " }, ";
        }
    }
    init += "}, { ";
    for(int i = 0; i < (nb > 0 ? 2 * nb : 2); ++i) {
        put_float(init, i < 2 * nb ? cs.axis_coef[i] : 0.0f);
        init += ", ";
    }
    init += "}, 0, 0, 0, 0 }";

    std::string tu;
    tu += "// generated by ptb_jit.cpp: a megakernel with this scene's coefficients as literals\n";
    tu += "#define PTB_JIT_SCENE_INIT " + init + "\n";
    tu += "#include \"ptb_kernels.h\"\n#include \"ptb_path_f32.cuh\"\n";
    tu += kind == kSorted ? "#include \"ptb_mega_sorted.cuh\"\n" : "#include \"ptb_mega_inplace.cuh\"\n";
    return tu;
}

bool JitCache::available()
{
    if(state_ == 0) {
        state_ = api().ok ? 1 : -1;
        if(state_ < 0) {
            error_ = "run-time compilation unavailable (PTB_JIT=0, or libnvrtc / libcuda not found)";
        }
    }
    return state_ > 0;
}

JitKernel const* JitCache::get(ConstSceneF32 const& cs, SceneCounts const& c, Kind kind, int inline_material, bool eager)
{
    if(!available() || !c.fits_const) {
        return nullptr;
    }
    int const ns = small_count(c), nb = big_count(c);
    // key: layout + in-place material + every coefficient the kernel reads, bit for bit
    std::vector<uint32_t> key = { static_cast<uint32_t>(c.small_near), static_cast<uint32_t>(c.small_both),
                                  static_cast<uint32_t>(c.big_near),   static_cast<uint32_t>(c.big_both),
                                  static_cast<uint32_t>(c.big_x),      static_cast<uint32_t>(c.big_y),
                                  static_cast<uint32_t>(c.big_z),      static_cast<uint32_t>(c.uniform_k),
                                  static_cast<uint32_t>(c.embed_ok),   static_cast<uint32_t>(c.pair_mask),
                                  static_cast<uint32_t>(kind == kSorted ? (inline_material < 0 ? 2 : inline_material == 0 ? 0 : 1) : 0),
                                  static_cast<uint32_t>(kind) };
    auto const push = [&](void const* p, size_t bytes) {
        size_t const n = bytes / sizeof(uint32_t);
        size_t const at = key.size();
        key.resize(at + n);
        std::memcpy(key.data() + at, p, n * sizeof(uint32_t));
    };
    push(cs.small_geo, static_cast<size_t>(ns) * sizeof(SmallGeo));
    push(cs.big_geo, static_cast<size_t>(nb) * sizeof(BigGeo));
    push(cs.axis_coef, static_cast<size_t>(2 * nb) * sizeof(float));
    ++clock_;
    auto it = cache_.find(key);
    bool const seen_before = it != cache_.end();
    if(it != cache_.end() && !it->second.pending) {
        it->second.last_use = clock_;
        return it->second.failed ? nullptr : &it->second;
    }
    if(it != cache_.end()) {
        cache_.erase(it);
    }
    // room (for a compiled module or a "seen once" marker alike -- an animation whose scene changes every frame must
    // not grow the map): drop the least recently used entry
    while(cache_.size() >= kMaxModules) {
        auto victim = cache_.begin();
        for(auto j = cache_.begin(); j != cache_.end(); ++j) {
            if(j->second.last_use < victim->second.last_use) {
                victim = j;
            }
        }
        if(victim->second.module != nullptr) {
            api().ModuleUnload(static_cast<CUmodule>(victim->second.module));
        }
        cache_.erase(victim);
    }
    if(!seen_before && !eager) {
        JitKernel seen;
        seen.pending = true;
        seen.last_use = clock_;
        cache_[key] = seen; // compile when it comes back
        return nullptr;
    }

    Api& a = api();
    JitKernel k;
    auto const t0 = std::chrono::steady_clock::now();
    std::string const tu = translation_unit(cs, c, kind);
    std::string const name = kernel_name(c, kind, inline_material);
    nvrtcProgram prog = nullptr;
    auto const fail = [&](std::string const& what) -> JitKernel const* {
        error_ = what;
        failures_++;
        if(prog != nullptr) {
            a.DestroyProgram(&prog);
        }
        if(k.module != nullptr) {
            a.ModuleUnload(static_cast<CUmodule>(k.module));
        }
        JitKernel bad;
        bad.failed = true;
        bad.last_use = clock_;
        cache_[key] = bad;
        return nullptr;
    };
    if(a.CreateProgram(&prog, tu.c_str(), "ptb_jit_tu.cu", kJitHeaderCount, kJitHeaderSources, kJitHeaderNames) != NVRTC_SUCCESS) {
        return fail("nvrtcCreateProgram failed");
    }
    if(a.AddNameExpression(prog, name.c_str()) != NVRTC_SUCCESS) {
        return fail("nvrtcAddNameExpression failed");
    }
    // development knobs (dev/ab_bench.py, dev/jit_offline.py): PTB_JIT_DEFINES="-DPTB_X=1 -DPTB_Y" adds macro definitions to this
    // process's run-time builds (A/B of #ifdef'ed experiments with ONE library), PTB_JIT_DUMP=dir keeps the generated
    // translation unit, the kernel's name and the cubin
    std::vector<std::string> extra;
    if(char const* defs = std::getenv("PTB_JIT_DEFINES")) {
        std::string const d = defs;
        for(size_t pos = 0; pos < d.size();) {
            size_t const sp = std::min(d.find(' ', pos), d.size());
            if(sp > pos) {
                extra.push_back(d.substr(pos, sp - pos));
            }
            pos = sp + 1;
        }
    }
    std::vector<char const*> opts = { "--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo" };
    for(std::string const& e : extra) {
        opts.push_back(e.c_str());
    }
    nvrtcResult const rc = a.CompileProgram(prog, static_cast<int>(opts.size()), opts.data());
    if(rc != NVRTC_SUCCESS) {
        size_t n = 0;
        a.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if(n > 1) {
            a.GetProgramLog(prog, &log[0]);
        }
        return fail("nvrtcCompileProgram failed: " + log.substr(0, 2000));
    }
    char const* lowered_fn = nullptr;
    if(a.GetLoweredName(prog, name.c_str(), &lowered_fn) != NVRTC_SUCCESS) {
        return fail("nvrtcGetLoweredName failed for " + name);
    }
    size_t bytes = 0;
    if(a.GetCUBINSize(prog, &bytes) != NVRTC_SUCCESS || bytes == 0) {
        return fail("nvrtcGetCUBINSize failed");
    }
    std::vector<char> cubin(bytes);
    if(a.GetCUBIN(prog, cubin.data()) != NVRTC_SUCCESS) {
        return fail("nvrtcGetCUBIN failed");
    }
    auto const cu_fail = [&](char const* what, CUresult e) {
        char const* s = nullptr;
        a.GetErrorString(e, &s);
        return fail(std::string(what) + ": " + (s != nullptr ? s : "unknown driver error"));
    };
    if(char const* dir = std::getenv("PTB_JIT_DUMP")) {
        std::string const base = std::string(dir) + "/jit_" + std::to_string(compiled_ + failures_) + "_" + std::to_string(static_cast<int>(kind));
        if(std::FILE* f = std::fopen((base + ".cu").c_str(), "wb")) {
            std::fwrite(tu.data(), 1, tu.size(), f);
            std::fclose(f);
        }
        if(std::FILE* f = std::fopen((base + ".name").c_str(), "wb")) {
            std::fwrite(name.data(), 1, name.size(), f);
            std::fclose(f);
        }
        if(std::FILE* f = std::fopen((base + ".cubin").c_str(), "wb")) {
            std::fwrite(cubin.data(), 1, cubin.size(), f);
            std::fclose(f);
        }
    }
    CUmodule mod = nullptr;
    CUresult e = a.ModuleLoadData(&mod, cubin.data());
    if(e != CUDA_SUCCESS) {
        return cu_fail("cuModuleLoadData", e);
    }
    k.module = mod;
    CUfunction fn = nullptr;
    if((e = a.ModuleGetFunction(&fn, mod, lowered_fn)) != CUDA_SUCCESS) {
        return cu_fail("cuModuleGetFunction", e);
    }
    int per_sm = 0;
    if((e = a.OccupancyMaxActiveBlocks(&per_sm, fn, 128, 0)) != CUDA_SUCCESS) {
        return cu_fail("cuOccupancyMaxActiveBlocksPerMultiprocessor", e);
    }
    a.DestroyProgram(&prog);
    prog = nullptr;
    k.function = fn;
    k.blocks_per_sm = per_sm < 1 ? 1 : per_sm;
    k.last_use = clock_;
    compile_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    compiled_++;
    auto const ins = cache_.emplace(std::move(key), k);
    return &ins.first->second;
}

cudaError_t JitCache::launch(JitKernel const& k, RenderParamsF32 const& p, int sm_count, cudaStream_t stream, int* launches)
{
    Api& a = api();
    cudaError_t ce = cudaMemsetAsync(&p.counters->tile_cursor, 0, sizeof(unsigned long long), stream);
    if(ce != cudaSuccess) {
        return ce;
    }
    // nothing else to stage: the coefficients are literals in the code, the cameras ride in the kernel argument
    unsigned long long blocks = static_cast<unsigned long long>(sm_count) * static_cast<unsigned long long>(k.blocks_per_sm);
    unsigned long long const needed = (static_cast<unsigned long long>(p.ntiles) + 3ull) / 4ull; // 4 warps per block
    blocks = blocks > needed ? needed : blocks;
    blocks = blocks < 1 ? 1 : blocks;
    RenderParamsF32 params = p;
    void* args[] = { &params };
    if(a.LaunchKernel(static_cast<CUfunction>(k.function), static_cast<unsigned>(blocks), 1, 1, 128, 1, 1, 0, reinterpret_cast<CUstream>(stream),
                      args, nullptr) != CUDA_SUCCESS) {
        return cudaErrorLaunchFailure;
    }
    if(launches != nullptr) {
        *launches += 1;
    }
    return cudaSuccess;
}

JitCache::~JitCache()
{
    Api& a = api();
    if(!a.ok) {
        return;
    }
    for(auto& kv : cache_) {
        if(kv.second.module != nullptr) {
            a.ModuleUnload(static_cast<CUmodule>(kv.second.module));
        }
    }
}

} // namespace ptb
