// ptb_smallpt_main.cpp -- the reference's stand-alone smallpt fork (sandbox/main.cpp) with its OpenMP
// render loop swapped for the C ABI.
//
// Same command line (argv[1] = total samples per pixel, / 4 per sub-pixel, sandbox/main.cpp:233), same
// 1024x768 frame, same scene array and camera constants (:94-122, :235), same output (./image.ppm, ASCII
// P3 with toInt rounding, :271-275).  Lines 241-269 -- `#pragma omp parallel for` over rows, erand48 per
// row -- are replaced by ptb_upload_scene / ptb_set_smallpt_camera / ptb_set_image /
// ptb_render(PTB_INTEGRATOR_SMALLPT) / ptb_resolve.  The sandbox's `Sphere spheres[]` has pt::sphere's
// layout, so a program that keeps its own array passes it as it is: ptb_upload_scene(ctx, spheres, n,
// sizeof(Sphere)).
#include "../../include/ptb200.h"
#include "pt.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char* argv[])
{
    int constexpr w = 1024;
    int constexpr h = 768;
    int const samps = argc == 2 ? atoi(argv[1]) / 4 : 1; // # samples

    std::vector<pt::sphere> const spheres = pt::smallpt_scene();
    pt::smallpt_camera const cam{};
    std::vector<pt::vec3> c(static_cast<size_t>(w) * h, pt::vec3{ 0, 0, 0 });

    ptb_context* ctx = nullptr;
    if(ptb_create(0, &ctx) != PTB_OK) {
        fprintf(stderr, "%s\n", ptb_last_error(nullptr));
        return 1;
    }
    int rc = ptb_upload_scene(ctx, spheres.data(), spheres.size(), sizeof(pt::sphere));
    rc = rc == PTB_OK ? ptb_set_smallpt_camera(ctx, reinterpret_cast<double const*>(&cam)) : rc;
    rc = rc == PTB_OK ? ptb_set_image(ctx, w, h, 2) : rc;
    fprintf(stderr, "Rendering (%d spp) on the GPU\n", samps * 4);
    rc = rc == PTB_OK ? ptb_render(ctx, 1, 0, static_cast<unsigned>(samps), PTB_INTEGRATOR_SMALLPT | PTB_PRECISION_FP32) : rc;
    rc = rc == PTB_OK ? ptb_resolve(ctx, reinterpret_cast<double*>(c.data())) : rc;
    if(rc != PTB_OK) {
        fprintf(stderr, "ptb call failed (%d): %s\n", rc, ptb_last_error(ctx));
        ptb_destroy(ctx);
        return 1;
    }
    ptb_stats st{};
    ptb_get_stats(ctx, &st);
    fprintf(stderr, "  %.2f ms on the device, %.1f Mpaths/s, %.1f Mrays/s\n", st.last_render_ms,
            static_cast<double>(st.paths) / st.last_render_ms * 1e-3, static_cast<double>(st.rays) / st.last_render_ms * 1e-3);
    rc = ptb_write_ppm_smallpt("image.ppm", reinterpret_cast<double const*>(c.data()), w, h);
    ptb_destroy(ctx);
    return rc == PTB_OK ? 0 : 1;
}
