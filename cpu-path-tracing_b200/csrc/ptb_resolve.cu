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

// One kernel for one GPU and for many.  `src` lists the accumulation buffers of ALL GPUs of the job (FP32 float4 sums
// and / or FP64 sums, per slot); on another GPU they are peer mappings (NVLink), as is `rgb_out` / `rgb8_out` when
// this GPU is not the one that owns the image.  The launch covers the pixel rows [y0, y1) -- this GPU's share -- so with
// G GPUs the G launches together are reduce-scatter + resolve + gather of the image in one pass over the data: every
// slot is read once from every buffer, nothing but pixels is written.
//
// Threads map to SLOTS for the loads (a warp reads 512 contiguous bytes of each buffer -- full lines over NVLink),
// the per-slot clamped means go through shared memory, then one thread per pixel adds its ns*ns strata in the
// reference's order (sy outer, sx inner: main.cpp:223-227).  Double arithmetic like the reference's.
constexpr int kResolveThreads = 256;

__global__ void __launch_bounds__(kResolveThreads)
    resolve_kernel(ResolveSources const src, uint32_t width, uint32_t height, uint32_t ns, uint32_t y0, uint32_t y1,
                   double* __restrict__ rgb_out, uint8_t* __restrict__ rgb8_out)
{
    __shared__ double s_mean[kResolveThreads * 3];
    uint32_t const nsub = ns * ns;
    uint32_t const ppb = kResolveThreads / nsub; // pixels per block (nsub <= 64)
    uint32_t const pix0 = y0 * width + blockIdx.x * ppb;
    uint32_t const pix_end = y1 * width;
    uint32_t const t = threadIdx.x;
    if(t < ppb * nsub) {
        uint32_t const pix = pix0 + t / nsub;
        double sr = 0.0, sg = 0.0, sb = 0.0, n = 0.0;
        if(pix < pix_end) {
            size_t const slot = static_cast<size_t>(pix0) * nsub + t;
            for(int g = 0; g < src.n; ++g) {
                if(src.accum32[g] != nullptr) {
                    float4 const a = __ldcv(src.accum32[g] + slot); // another GPU wrote it: never from a stale line
                    sr += a.x;
                    sg += a.y;
                    sb += a.z;
                    n += a.w;
                }
                if(src.accum64[g] != nullptr) {
                    double2 const lo = __ldcv(reinterpret_cast<double2 const*>(src.accum64[g] + 4 * slot));
                    double2 const hi = __ldcv(reinterpret_cast<double2 const*>(src.accum64[g] + 4 * slot) + 1);
                    sr += lo.x;
                    sg += lo.y;
                    sb += hi.x;
                    n += hi.y;
                }
            }
        }
        double mr = 0.0, mg = 0.0, mb = 0.0;
        if(n > 0.0) {
            double const inv = 1.0 / n;
            mr = clamp01(sr * inv);
            mg = clamp01(sg * inv);
            mb = clamp01(sb * inv);
        }
        s_mean[3 * t + 0] = mr;
        s_mean[3 * t + 1] = mg;
        s_mean[3 * t + 2] = mb;
    }
    __syncthreads();
    uint32_t const pix = pix0 + t;
    if(t >= ppb || pix >= pix_end) {
        return;
    }
    double const w = 1.0 / static_cast<double>(nsub);
    double r = 0.0, g = 0.0, b = 0.0;
    for(uint32_t k = 0; k < nsub; ++k) {
        r = r + s_mean[3 * (t * nsub + k) + 0] * w;
        g = g + s_mean[3 * (t * nsub + k) + 1] * w;
        b = b + s_mean[3 * (t * nsub + k) + 2] * w;
    }
    uint32_t const y = pix / width;
    uint32_t const x = pix - y * width;
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

cudaError_t launch_resolve_rows(ResolveSources const& src, uint32_t width, uint32_t height, uint32_t ns, uint32_t y0,
                                uint32_t y1, double* rgb_out, uint8_t* rgb8_out, cudaStream_t stream)
{
    if(y1 <= y0 || width == 0 || ns == 0 || ns * ns > 64u) {
        return y1 <= y0 || width == 0 ? cudaSuccess : cudaErrorInvalidValue;
    }
    uint32_t const ppb = static_cast<uint32_t>(kResolveThreads) / (ns * ns);
    uint32_t const npix = (y1 - y0) * width;
    resolve_kernel<<<(npix + ppb - 1) / ppb, kResolveThreads, 0, stream>>>(src, width, height, ns, y0, y1, rgb_out, rgb8_out);
    return cudaGetLastError();
}

cudaError_t launch_resolve(float4 const* accum32, double const* accum64, uint32_t width, uint32_t height, uint32_t ns,
                           double* rgb_out, uint8_t* rgb8_out, cudaStream_t stream)
{
    ResolveSources src{};
    src.n = 1;
    src.accum32[0] = accum32;
    src.accum64[0] = accum64;
    return launch_resolve_rows(src, width, height, ns, 0, height, rgb_out, rgb8_out, stream);
}

} // namespace ptb
