// ptb_main.cpp -- the reference's program with its CPU render loop swapped for the C ABI.
//
// Same command line as /root/reference/src/main.cpp:199-248 (argv[1] = total samples per
// pixel, divided by the 2x2 sub-pixels, main.cpp:206), same defaults (1024x768, the
// box_mirror scene the reference ships with, main.cpp:25,204-205,208), same output
// (./image.ppm, ASCII P3, gamma 2.2).  Only lines 214-236 -- the taskflow row tasks -- are
// replaced: ptb_upload_scene / ptb_set_camera / ptb_set_image / ptb_render / ptb_resolve.
//
//   ptb_main [spp] [--scene simple|box|box_mirror|dof_glass|spheres10k] [--size WxH]
//            [--seed N] [--fp64] [--out image.ppm] [--device N | --devices 0,1,..,7]
//            [--variant sorted|inplace|wavefront] [--precompiled] [--scan] [--p6 | --rgb8]
//
// --devices: ONE process, several GPUs (ptb_create_multi): the samples of every sub-pixel are split over them and summed
// inside the library -- still one blocking ptb_render + one ptb_resolve, as the reference has one executor.run().wait().
// --p6 / --rgb8: the output stage for large images: gamma + 8-bit conversion on the GPU (ptb_resolve_rgb8: 3 bytes per
// pixel come back instead of 24), written as binary "P6" or as the reference's "P3" tokens (main.cpp:240-247).
#include "../../include/ptb200.h"
#include "pt.hpp"

#include <algorithm>
#include <chrono>
#include <cstdint>
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
    std::vector<int> devices{ 0 };
    int output_mode = 0; // 0: FP64 image + the reference's writer, 1: 8-bit image as P6, 2: 8-bit image as P3
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
            devices = { std::atoi(next().c_str()) };
        }
        else if(a == "--devices") {
            devices.clear();
            std::string const list = next();
            for(std::size_t pos = 0; pos <= list.size();) {
                std::size_t const comma = std::min(list.find(',', pos), list.size());
                devices.push_back(std::atoi(list.substr(pos, comma - pos).c_str()));
                pos = comma + 1;
            }
        }
        else if(a == "--p6") {
            output_mode = 1;
        }
        else if(a == "--rgb8") {
            output_mode = 2;
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
    int rc = devices.size() == 1 ? ptb_create(devices[0], &ctx) : ptb_create_multi(devices.data(), static_cast<int>(devices.size()), &ctx);
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
              << height << " on " << devices.size() << " GPU(s), first " << devices[0] << '\n';
    auto const t0 = std::chrono::steady_clock::now();
    if((rc = ptb_render(ctx, seed, 0, static_cast<unsigned>(samps), flags)) != PTB_OK) {
        return die(ctx, "ptb_render", rc);
    }
    std::vector<std::uint8_t> image8{};
    if(output_mode == 0) {
        rc = ptb_resolve(ctx, reinterpret_cast<double*>(image.data()));
    }
    else {
        image8.resize(static_cast<std::size_t>(width) * static_cast<std::size_t>(height) * 3);
        rc = ptb_resolve_rgb8(ctx, image8.data());
    }
    if(rc != PTB_OK) {
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

    int32_t comm[6] = { 1, 0, 0, 0, 0, 0 };
    ptb_comm_info(ctx, comm);
    if(comm[0] > 1) {
        std::cerr << "  " << comm[0] << " GPUs, sum + resolve " << st.last_resolve_ms << " ms over "
                  << (comm[2] == PTB_TRANSPORT_PEER ? "peer mappings (one fused kernel per GPU)" : "NCCL") << '\n';
    }
    auto const w0 = std::chrono::steady_clock::now();
    rc = output_mode == 0 ? ptb_write_ppm(out.c_str(), reinterpret_cast<double const*>(image.data()), width, height)
                          : ptb_write_ppm_rgb8(out.c_str(), image8.data(), width, height, output_mode == 1 ? 1 : 0);
    std::cerr << "  image written in " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count()
              << " ms\n";
    ptb_destroy(ctx);
    if(rc != PTB_OK) {
        std::cerr << "cannot write " << out << '\n';
        return 1;
    }
    return 0;
}
