// ptb_resolve.cu -- accumulation buffer -> image.
//
// Replaces the tail of render_subpixel, /root/reference/src/main.cpp:192-196, and the
// row flip of main.cpp:181: per sub-pixel mean, clamp to [0,1] (pt::clamp,
// utils.cpp:6-9), then the average of the ns*ns strata, written at row H-1-y.
// The clamp is NON-linear, which is why the GPUs exchange un-clamped per-stratum
// sums and this kernel runs once, after the reduce.
// Optionally also pt::color_to_int (utils.cpp:11-16) for 8-bit output.
#include "ptb_kernels.h"

namespace ptb {

namespace {

__device__ __forceinline__ double clamp01(double v)
{
    return v < 0.0 ? 0.0 : (1.0 < v ? 1.0 : v);
}

__global__ void resolve_kernel(float4 const* __restrict__ accum32, double const* __restrict__ accum64, uint32_t width,
                               uint32_t height, uint32_t ns, double* __restrict__ rgb_out, uint8_t* __restrict__ rgb8_out)
{
    uint32_t const pix = blockIdx.x * blockDim.x + threadIdx.x;
    if(pix >= width * height) {
        return;
    }
    uint32_t const y = pix / width;
    uint32_t const x = pix - y * width;
    uint32_t const nsub = ns * ns;
    double const w = 1.0 / static_cast<double>(nsub);
    double r = 0.0, g = 0.0, b = 0.0;
    for(uint32_t k = 0; k < nsub; ++k) {
        size_t const slot = static_cast<size_t>(pix) * nsub + k;
        double sr = 0.0, sg = 0.0, sb = 0.0, n = 0.0;
        if(accum32 != nullptr) {
            float4 const a = accum32[slot];
            sr += a.x;
            sg += a.y;
            sb += a.z;
            n += a.w;
        }
        if(accum64 != nullptr) {
            sr += accum64[4 * slot + 0];
            sg += accum64[4 * slot + 1];
            sb += accum64[4 * slot + 2];
            n += accum64[4 * slot + 3];
        }
        if(n > 0.0) {
            double const inv = 1.0 / n;
            r = r + clamp01(sr * inv) * w;
            g = g + clamp01(sg * inv) * w;
            b = b + clamp01(sb * inv) * w;
        }
    }
    size_t const row = static_cast<size_t>(height - y - 1) * width + x;
    if(rgb_out != nullptr) {
        rgb_out[3 * row + 0] = r;
        rgb_out[3 * row + 1] = g;
        rgb_out[3 * row + 2] = b;
    }
    if(rgb8_out != nullptr) {
        rgb8_out[3 * row + 0] = static_cast<uint8_t>(static_cast<int>(round(pow(clamp01(r), 1.0 / 2.2) * 255.0)));
        rgb8_out[3 * row + 1] = static_cast<uint8_t>(static_cast<int>(round(pow(clamp01(g), 1.0 / 2.2) * 255.0)));
        rgb8_out[3 * row + 2] = static_cast<uint8_t>(static_cast<int>(round(pow(clamp01(b), 1.0 / 2.2) * 255.0)));
    }
}

} // namespace

// FP32 roofline calibration: 8 independent FFMA chains per thread, 2048 threads per SM.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
          a7 = a0 + 7.f;
    float const b = seed * 0.5f, c = seed * 0.25f;
    for(int it = 0; it < iters; ++it) {
#pragma unroll
        for(int u = 0; u < 8; ++u) {
            asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                         "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;"
                         : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7)
                         : "f"(b), "f"(c));
        }
    }
    float const r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if(r == 123.456f) {
        out[0] = r;
    }
}

cudaError_t launch_fp32_peak(int sm_count, int iters, float* scratch, cudaStream_t stream, double* flop_out)
{
    int const grid = sm_count * 8;
    fp32_peak_kernel<<<grid, 256, 0, stream>>>(scratch, iters, 1.0f);
    *flop_out = static_cast<double>(grid) * 256.0 * static_cast<double>(iters) * 64.0 * 2.0;
    return cudaGetLastError();
}

cudaError_t launch_resolve(float4 const* accum32, double const* accum64, uint32_t width, uint32_t height, uint32_t ns,
                           double* rgb_out, uint8_t* rgb8_out, cudaStream_t stream)
{
    uint32_t const npix = width * height;
    if(npix == 0) {
        return cudaSuccess;
    }
    unsigned const threads = 256;
    resolve_kernel<<<(npix + threads - 1) / threads, threads, 0, stream>>>(accum32, accum64, width, height, ns, rgb_out,
                                                                           rgb8_out);
    return cudaGetLastError();
}

} // namespace ptb
