"""Apply the INTEGRATION.md section 1 patch to the reference's src/main.cpp, writing the result somewhere ELSE
(a build directory; the reference tree is read-only and its sources are never copied into this repository):

    python oracle/ref/make_dropin.py /root/reference/src/main.cpp /tmp/build/main_dropin.cpp

Everything that builds the scene, the camera and the PPM file is kept; the taskflow row-task block
(src/main.cpp:214-236) becomes eight calls into include/ptb200.h (every GPU of the box behind one context).  oracle/Makefile compiles the result with the
reference's own headers and translation units into oracle/_ref/cpu_path_tracer_b200 -- the reference PROGRAM running
on the B200 library -- which tests/test_reference_dropin.py links-checks on CPU and runs on the GPU box.
"""
import re
import sys

ABI_BLOCK = '''    static_assert(sizeof(pt::sphere) == PTB_SPHERE_BYTES && sizeof(pt::camera) == PTB_CAMERA_BYTES,
                  "the reference's own records cross the boundary as they are");
    // every GPU of the machine, as the reference's executor uses every core; PTB_GPUS=n caps it
    int devices[16];
    int n_gpus = ptb_device_count() < 16 ? ptb_device_count() : 16;
    if(char const* cap = std::getenv("PTB_GPUS")) {
        n_gpus = std::atoi(cap) < n_gpus ? std::atoi(cap) : n_gpus;
    }
    n_gpus = n_gpus < 1 ? 1 : n_gpus; // none: let the library say so
    for(int i = 0; i < n_gpus; ++i) {
        devices[i] = i;
    }
    ptb_context* gpu = nullptr;
    if(ptb_create_multi(devices, n_gpus, &gpu) != PTB_OK) {
        std::cerr << ptb_last_error(nullptr) << '\\n';
        return 1;
    }
    int rc = ptb_upload_scene(gpu, some_scene.spheres.data(), some_scene.spheres.size(), sizeof(pt::sphere));
    rc = rc == PTB_OK ? ptb_set_camera(gpu, &cam, sizeof(cam)) : rc;
    rc = rc == PTB_OK ? ptb_set_image(gpu, width, height, num_subpixels) : rc;
    rc = rc == PTB_OK ? ptb_render(gpu, /*seed*/ 1, /*first_sample*/ 0, static_cast<unsigned>(samps),
                                   PTB_VARIANT_MEGAKERNEL_SORTED | PTB_PRECISION_FP32)
                      : rc;
    rc = rc == PTB_OK ? ptb_resolve(gpu, reinterpret_cast<double*>(image.data())) : rc;
    if(rc != PTB_OK) {
        std::cerr << ptb_last_error(gpu) << '\\n';
        return 1;
    }
    ptb_destroy(gpu);
'''


def patch(text: str) -> str:
    if "#include <taskflow/taskflow.hpp>" not in text:
        raise SystemExit("make_dropin: taskflow include not found -- has the reference changed?")
    text = text.replace("#include <taskflow/taskflow.hpp>", "#include <ptb200.h> // C ABI of the B200 render loop\n#include <cstdlib>")
    block = re.compile(r"^    tf::Executor executor\{\};\n.*?^    executor\.run\(taskflow\)\.wait\(\);\n", re.S | re.M)
    text, n = block.subn(lambda _m: ABI_BLOCK, text)
    if n != 1:
        raise SystemExit("make_dropin: the row-task block (src/main.cpp:214-236) was not found exactly once")
    if "tf::" in text:
        raise SystemExit("make_dropin: taskflow is still referenced after the patch")
    return text


SANDBOX_BLOCK = '''    Vec* c = new Vec[w * h];
    static_assert(sizeof(Sphere) == PTB_SPHERE_BYTES && sizeof(Vec) == 3 * sizeof(double), "the program's own records cross the boundary");
    double const cam8[8] = { cam.o.x, cam.o.y, cam.o.z, cam.d.x, cam.d.y, cam.d.z, .5135, 140 };
    int devices[16]; // every GPU of the machine, as the OpenMP loop used every core; PTB_GPUS=n caps it
    int n_gpus = ptb_device_count() < 16 ? ptb_device_count() : 16;
    if(char const* cap = getenv("PTB_GPUS")) {
        n_gpus = atoi(cap) < n_gpus ? atoi(cap) : n_gpus;
    }
    n_gpus = n_gpus < 1 ? 1 : n_gpus; // none: let the library say so
    for(int i = 0; i < n_gpus; ++i) {
        devices[i] = i;
    }
    ptb_context* gpu = nullptr;
    if(ptb_create_multi(devices, n_gpus, &gpu) != PTB_OK) {
        fprintf(stderr, "%s\\n", ptb_last_error(nullptr));
        return 1;
    }
    int rc = ptb_upload_scene(gpu, spheres, sizeof(spheres) / sizeof(Sphere), sizeof(Sphere));
    rc = rc == PTB_OK ? ptb_set_smallpt_camera(gpu, cam8) : rc;
    rc = rc == PTB_OK ? ptb_set_image(gpu, w, h, 2) : rc;
    rc = rc == PTB_OK ? ptb_render(gpu, /*seed*/ 1, 0, static_cast<unsigned>(samps), PTB_INTEGRATOR_SMALLPT | PTB_PRECISION_FP32) : rc;
    rc = rc == PTB_OK ? ptb_resolve(gpu, reinterpret_cast<double*>(c)) : rc;
    if(rc != PTB_OK) {
        fprintf(stderr, "%s\\n", ptb_last_error(gpu));
        return 1;
    }
    ptb_destroy(gpu);

'''


def patch_sandbox(text: str) -> str:
    """sandbox/main.cpp (the stand-alone smallpt fork): its OpenMP row loop (:236-269) becomes the ABI calls of
    INTEGRATION.md section 2; scene array, camera constants and the PPM writer stay."""
    if "#pragma omp parallel for" not in text:
        raise SystemExit("make_dropin: the OpenMP loop of sandbox/main.cpp was not found -- has the reference changed?")
    block = re.compile(r"^    Vec const cx = .*?(?=^    FILE\* f = fopen)", re.S | re.M)
    text, n = block.subn(lambda _m: SANDBOX_BLOCK, text)
    if n != 1:
        raise SystemExit("make_dropin: the render loop of sandbox/main.cpp was not found exactly once")
    if "erand48(Xi)" in text.split("int main")[1]:
        raise SystemExit("make_dropin: the render loop is still there after the patch")
    return "#include <ptb200.h> // C ABI of the B200 render loop\n" + text


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        body = f.read()
    out = patch_sandbox(body) if "--sandbox" in sys.argv[3:] else patch(body)
    with open(dst, "w") as f:
        f.write(out)
