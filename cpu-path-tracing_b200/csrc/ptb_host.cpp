// ptb_host.cpp -- the entry points of include/ptb200.h that need no GPU: built-in
// scenes, camera derivation, PPM output.  All of it is host/pt.hpp (the mirror of the
// reference's scene layer) behind plain pointers.
#include "../../include/ptb200.h"
#include "../host/pt.hpp"

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

// No exception may unwind into a caller that is C, Go or ctypes (function-try-block tail of every entry point below;
// these have no context to leave a message in).
#define PTB_CATCH(ctx, entry) \
    catch(std::bad_alloc const&) \
    { \
        return PTB_ERR_MEMORY; \
    } \
    catch(...) \
    { \
        return PTB_ERR_INTERNAL; \
    }

namespace {
template<class ToInt>
int write_ppm_with(char const* path, double const* rgb, int width, int height, ToInt to_int)
{
    if(path == nullptr || rgb == nullptr || width <= 0 || height <= 0) {
        return PTB_ERR_ARGUMENT;
    }
    // the text first (it may not fit in memory: nothing is open yet when that throws), then the file
    std::string buf;
    buf.reserve(static_cast<size_t>(width) * static_cast<size_t>(height) * 12 + 32);
    buf += "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255\n";
    size_t const n = static_cast<size_t>(width) * static_cast<size_t>(height) * 3;
    for(size_t i = 0; i < n; ++i) {
        buf += std::to_string(to_int(rgb[i]));
        buf += ' ';
    }
    std::FILE* f = std::fopen(path, "wb");
    if(f == nullptr) {
        return PTB_ERR_IO;
    }
    bool const ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    return (std::fclose(f) == 0 && ok) ? PTB_OK : PTB_ERR_IO;
}
} // namespace

extern "C" {

int ptb_abi_version(void)
try {
    return PTB_ABI_VERSION;
}
PTB_CATCH(nullptr, "ptb_abi_version")

// pt::camera::with_config, /root/reference/src/camera.cpp:3-17
int ptb_camera_with_config(void const* camera_config, void* camera_out)
try {
    if(camera_config == nullptr || camera_out == nullptr) {
        return PTB_ERR_ARGUMENT;
    }
    pt::camera_config cfg{};
    std::memcpy(static_cast<void*>(&cfg), camera_config, sizeof(cfg));
    pt::camera const cam = pt::camera::with_config(cfg);
    std::memcpy(camera_out, static_cast<void const*>(&cam), sizeof(cam));
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_camera_with_config")

int ptb_builtin_scene(char const* name, int width, int height, void* spheres_out, size_t capacity, size_t* count_out,
                      void* camera_config_out)
try {
    if(name == nullptr || width <= 0 || height <= 0) {
        return PTB_ERR_ARGUMENT;
    }
    std::string const which{ name };
    pt::scene scn{};
    if(which == "simple") {
        scn = pt::simple_scene(width, height);
    }
    else if(which == "box") {
        scn = pt::box_scene(width, height);
    }
    else if(which == "box_mirror") {
        scn = pt::box_mirror_scene(width, height);
    }
    else if(which == "dof_glass") {
        scn = pt::dof_glass_scene(width, height);
    }
    else if(which == "spheres10k") {
        scn = pt::spheres10k_scene(width, height);
    }
    else {
        return PTB_ERR_ARGUMENT;
    }
    if(count_out != nullptr) {
        *count_out = scn.spheres.size();
    }
    if(camera_config_out != nullptr) {
        std::memcpy(camera_config_out, static_cast<void const*>(&scn.camera_parameters), sizeof(pt::camera_config));
    }
    if(spheres_out != nullptr) {
        if(capacity < scn.spheres.size()) {
            return PTB_ERR_ARGUMENT;
        }
        std::memcpy(spheres_out, static_cast<void const*>(scn.spheres.data()), sizeof(pt::sphere) * scn.spheres.size());
    }
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_builtin_scene")

int ptb_builtin_smallpt_scene(void* spheres_out, size_t capacity, size_t* count_out, double* cam8_out)
try {
    std::vector<pt::sphere> const spheres = pt::smallpt_scene();
    if(count_out != nullptr) {
        *count_out = spheres.size();
    }
    if(cam8_out != nullptr) {
        pt::smallpt_camera const cam{};
        std::memcpy(cam8_out, static_cast<void const*>(&cam), sizeof(cam));
    }
    if(spheres_out != nullptr) {
        if(capacity < spheres.size()) {
            return PTB_ERR_ARGUMENT;
        }
        std::memcpy(spheres_out, static_cast<void const*>(spheres.data()), sizeof(pt::sphere) * spheres.size());
    }
    return PTB_OK;
}
PTB_CATCH(nullptr, "ptb_builtin_smallpt_scene")

// sandbox/main.cpp:271-275: same P3 layout, toInt rounding
int ptb_write_ppm_smallpt(char const* path, double const* rgb, int width, int height)
try {
    return write_ppm_with(path, rgb, width, height, [](double v) { return pt::smallpt_to_int(v); });
}
PTB_CATCH(nullptr, "ptb_write_ppm_smallpt")

// The 8-bit values of ptb_resolve_rgb8 (pt::color_to_int on the GPU) to disk: raw "P6", or the reference's "P3"
// token layout (main.cpp:240-247) through a table of the 256 possible tokens -- no pow(), no integer formatting.
int ptb_write_ppm_rgb8(char const* path, uint8_t const* rgb8, int width, int height, int binary)
try {
    if(path == nullptr || rgb8 == nullptr || width <= 0 || height <= 0) {
        return PTB_ERR_ARGUMENT;
    }
    size_t const n = static_cast<size_t>(width) * static_cast<size_t>(height) * 3;
    std::string const head = std::string(binary != 0 ? "P6\n" : "P3\n") + std::to_string(width) + " " + std::to_string(height) + "\n255\n";
    // the P3 text first (it may not fit in memory: nothing is open yet when that throws), then the file
    std::vector<char> text;
    size_t m = 0;
    if(binary == 0) {
        char token[256][4];
        unsigned char len[256];
        for(int v = 0; v < 256; ++v) {
            len[v] = static_cast<unsigned char>(std::snprintf(token[v], 4, "%d", v));
            token[v][len[v]++] = ' '; // "{v} ": at most four bytes, no terminator kept
        }
        text.resize(n * 4);
        for(size_t i = 0; i < n; ++i) {
            unsigned const v = rgb8[i];
            std::memcpy(&text[m], token[v], 4);
            m += len[v];
        }
    }
    std::FILE* f = std::fopen(path, "wb");
    if(f == nullptr) {
        return PTB_ERR_IO;
    }
    bool ok = std::fwrite(head.data(), 1, head.size(), f) == head.size();
    ok = ok && (binary != 0 ? std::fwrite(rgb8, 1, n, f) == n : std::fwrite(text.data(), 1, m, f) == m);
    return (std::fclose(f) == 0 && ok) ? PTB_OK : PTB_ERR_IO;
}
PTB_CATCH(nullptr, "ptb_write_ppm_rgb8")

// Same bytes as the writer of /root/reference/src/main.cpp:240-247: header
// "P3\n{w} {h}\n255\n", then "{r} {g} {b} " per pixel, no newlines.
int ptb_write_ppm(char const* path, double const* rgb, int width, int height)
try {
    return write_ppm_with(path, rgb, width, height, [](double v) { return pt::color_to_int(v); });
}
PTB_CATCH(nullptr, "ptb_write_ppm")

} // extern "C"
