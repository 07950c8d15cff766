// ptb_main.cpp -- the reference's program with its CPU render loop swapped for the C ABI.
//
// Same command line as /root/reference/src/main.cpp:199-248 (argv[1] = total samples per
// pixel, divided by the 2x2 sub-pixels, main.cpp:206), same defaults (1024x768, the
// box_mirror scene the reference ships with, main.cpp:25,204-205,208), same output
// (./image.ppm, ASCII P3, gamma 2.2).  Only lines 214-236 -- the taskflow row tasks -- are
// replaced: ptb_upload_scene / ptb_set_camera / ptb_set_image / ptb_render / ptb_resolve.
//
//   ptb_main [spp] [--scene simple|box|box_mirror|dof_glass|spheres10k] [--size WxH]
//            [--seed N] [--fp64] [--out image.ppm] [--device N]
//            [--variant sorted|inplace|wavefront] [--precompiled] [--scan]
#include "../../include/ptb200.h"
#include "pt.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace {

auto die(ptb_context* ctx, char const* what, int rc) -> int
{
    std::cerr << what << " failed (" << rc << "): " << ptb_last_error(ctx) << '\n';
    ptb_destroy(ctx);
    return 1;
}

} // namespace

auto main(int argc, char* argv[]) -> int
{
    constexpr int num_subpixels = 2;
    int width = 1024;
    int height = 768;
    int spp = 4;
    int device = 0;
    unsigned long long seed = 1;
    unsigned flags = PTB_VARIANT_MEGAKERNEL_SORTED | PTB_PRECISION_FP32;
    std::string scene_name{ "box_mirror" };
    std::string out{ "image.ppm" };

    for(int i = 1; i < argc; ++i) {
        std::string const a{ argv[i] };
        auto const next = [&]() -> std::string { return i + 1 < argc ? std::string{ argv[++i] } : std::string{}; };
        if(a == "--scene") {
            scene_name = next();
        }
        else if(a == "--size") {
            std::string const s = next();
            if(std::sscanf(s.c_str(), "%dx%d", &width, &height) != 2) {
                std::cerr << "--size wants WxH\n";
                return 2;
            }
        }
        else if(a == "--seed") {
            seed = std::strtoull(next().c_str(), nullptr, 0);
        }
        else if(a == "--device") {
            device = std::atoi(next().c_str());
        }
        else if(a == "--out") {
            out = next();
        }
        else if(a == "--fp64") {
            flags = PTB_VARIANT_MEGAKERNEL | PTB_PRECISION_FP64;
        }
        else if(a == "--variant") {
            std::string const v = next();
            unsigned const var = v == "sorted" ? PTB_VARIANT_MEGAKERNEL_SORTED : (v == "inplace" ? PTB_VARIANT_MEGAKERNEL : (v == "wavefront" ? PTB_VARIANT_WAVEFRONT : 0xFu));
            if(var == 0xFu) {
                std::cerr << "--variant wants sorted, inplace or wavefront\n";
                return 2;
            }
            flags = (flags & ~static_cast<unsigned>(PTB_VARIANT_MASK)) | var;
        }
        else if(a == "--precompiled") { // no run-time compilation for the scene
            flags |= PTB_CODEGEN_PRECOMPILED;
        }
        else if(a == "--scan") { // the reference's linear closest-hit scan even when a hierarchy exists
            flags |= PTB_ACCEL_SCAN;
        }
        else {
            spp = std::stoi(a); // throws on garbage, like the reference's std::stoi (main.cpp:206)
        }
    }
    int const samps = spp / (num_subpixels * num_subpixels);

    pt::scene some_scene{};
    if(scene_name == "simple") {
        some_scene = pt::simple_scene(width, height);
    }
    else if(scene_name == "box") {
        some_scene = pt::box_scene(width, height);
    }
    else if(scene_name == "box_mirror") {
        some_scene = pt::box_mirror_scene(width, height);
    }
    else if(scene_name == "dof_glass") {
        some_scene = pt::dof_glass_scene(width, height);
    }
    else if(scene_name == "spheres10k") {
        some_scene = pt::spheres10k_scene(width, height);
    }
    else {
        std::cerr << "unknown scene " << scene_name << '\n';
        return 2;
    }
    auto const cam = pt::camera::with_config(some_scene.camera_parameters);
    std::vector<pt::vec3> image{};
    image.resize(static_cast<std::size_t>(width) * static_cast<std::size_t>(height), pt::vec3{ 0, 0, 0 });

    ptb_context* ctx = nullptr;
    int rc = ptb_create(device, &ctx);
    if(rc != PTB_OK) {
        std::cerr << "ptb_create failed (" << rc << "): " << ptb_last_error(nullptr) << '\n';
        return 1;
    }
    if((rc = ptb_upload_scene(ctx, some_scene.spheres.data(), some_scene.spheres.size(), sizeof(pt::sphere))) != PTB_OK) {
        return die(ctx, "ptb_upload_scene", rc);
    }
    if((rc = ptb_set_camera(ctx, &cam, sizeof(cam))) != PTB_OK) {
        return die(ctx, "ptb_set_camera", rc);
    }
    if((rc = ptb_set_image(ctx, width, height, num_subpixels)) != PTB_OK) {
        return die(ctx, "ptb_set_image", rc);
    }

    std::cerr << "Rendering (" << samps * num_subpixels * num_subpixels << " spp) " << scene_name << ' ' << width << 'x'
              << height << " on GPU " << device << '\n';
    auto const t0 = std::chrono::steady_clock::now();
    if((rc = ptb_render(ctx, seed, 0, static_cast<unsigned>(samps), flags)) != PTB_OK) {
        return die(ctx, "ptb_render", rc);
    }
    if((rc = ptb_resolve(ctx, reinterpret_cast<double*>(image.data()))) != PTB_OK) {
        return die(ctx, "ptb_resolve", rc);
    }
    auto const t1 = std::chrono::steady_clock::now();

    ptb_stats st{};
    ptb_get_stats(ctx, &st);
    double const wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    int32_t jit[5] = { 0, 0, 0, 0, 0 };
    ptb_jit_info(ctx, jit);
    std::cerr << "  " << (jit[3] != 0 ? "kernel compiled for this scene at run time, " : "precompiled kernel, ")
              << "device " << st.last_render_ms << " ms, wall " << wall_ms << " ms, "
              << static_cast<double>(st.paths) / (st.last_render_ms > 0 ? st.last_render_ms : 1) * 1e-3 << " Mpaths/s, "
              << static_cast<double>(st.rays) / (st.last_render_ms > 0 ? st.last_render_ms : 1) * 1e-3 << " Mrays/s\n";

    rc = ptb_write_ppm(out.c_str(), reinterpret_cast<double const*>(image.data()), width, height);
    ptb_destroy(ctx);
    if(rc != PTB_OK) {
        std::cerr << "cannot write " << out << '\n';
        return 1;
    }
    return 0;
}
