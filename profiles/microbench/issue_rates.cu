// issue_rates.cu -- measured instruction issue rates on the B200 this job runs on.
// The render loop is FP32-issue-bound, and MEASURED_PEAKS.json only
// carries HBM and bf16-tensor peaks, so the FP32 roofline denominator is measured here:
// warp-instructions per clock per SM for the instruction forms the megakernel is made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rates issue_rates.cu && ./issue_rates
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__constant__ float c_k[64];

#define REP8(x) x x x x x x x x
#define REP64(x) REP8(REP8(x))

// each variant: 8 independent accumulators per thread, ITER x 64 x 8 ops
template<int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float seed, double dseed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float const b = seed * 0.5f, c = seed * 0.25f;
    double d0 = dseed + threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    double const db = dseed * 0.5, dc = dseed * 0.25;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7;
    for(int it = 0; it < iters; ++it) {
        if(MODE == 0) { // FFMA reg,reg,reg
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                              "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
        else if(MODE == 1) { // FFMA reg, const, reg
            REP8(a0 = fmaf(a0, c_k[0], a1); a1 = fmaf(a1, c_k[1], a2); a2 = fmaf(a2, c_k[2], a3); a3 = fmaf(a3, c_k[3], a4);
                 a4 = fmaf(a4, c_k[4], a5); a5 = fmaf(a5, c_k[5], a6); a6 = fmaf(a6, c_k[6], a7); a7 = fmaf(a7, c_k[7], a0);)
        }
        else if(MODE == 2) { // FADD
            REP8(asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                              "add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b));)
        }
        else if(MODE == 3) { // MUFU.SQRT
            REP8(asm volatile("sqrt.approx.ftz.f32 %0, %0; sqrt.approx.ftz.f32 %1, %1; sqrt.approx.ftz.f32 %2, %2; sqrt.approx.ftz.f32 %3, %3;"
                              "sqrt.approx.ftz.f32 %4, %4; sqrt.approx.ftz.f32 %5, %5; sqrt.approx.ftz.f32 %6, %6; sqrt.approx.ftz.f32 %7, %7;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));)
        }
        else if(MODE == 4) { // DFMA
            REP8(asm volatile("fma.rn.f64 %0, %0, %4, %5; fma.rn.f64 %1, %1, %4, %5; fma.rn.f64 %2, %2, %4, %5; fma.rn.f64 %3, %3, %4, %5;"
                              "fma.rn.f64 %0, %0, %4, %5; fma.rn.f64 %1, %1, %4, %5; fma.rn.f64 %2, %2, %4, %5; fma.rn.f64 %3, %3, %4, %5;"
                              : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(db), "d"(dc));)
        }
        else if(MODE == 5) { // integer min (VIMNMX, ALU pipe)
            REP8(asm volatile("min.u32 %0, %0, %8; min.u32 %1, %1, %8; min.u32 %2, %2, %8; min.u32 %3, %3, %8;"
                              "min.u32 %4, %4, %8; min.u32 %5, %5, %8; min.u32 %6, %6, %8; min.u32 %7, %7, %8;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(u0 ^ it));)
        }
        else if(MODE == 6) { // 1:1 mix FFMA (fma pipe) + min.u32 (alu pipe)
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; min.u32 %4, %4, %10; fma.rn.f32 %1, %1, %8, %9; min.u32 %5, %5, %10;"
                              "fma.rn.f32 %2, %2, %8, %9; min.u32 %6, %6, %10; fma.rn.f32 %3, %3, %8, %9; min.u32 %7, %7, %10;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c), "r"(u0 ^ it));)
        }
        else if(MODE == 7) { // 3:1 mix FFMA + DFMA
            REP8(asm volatile("fma.rn.f32 %0, %0, %6, %7; fma.rn.f32 %1, %1, %6, %7; fma.rn.f32 %2, %2, %6, %7; fma.rn.f64 %4, %4, %8, %9;"
                              "fma.rn.f32 %3, %3, %6, %7; fma.rn.f32 %0, %0, %6, %7; fma.rn.f32 %1, %1, %6, %7; fma.rn.f64 %5, %5, %8, %9;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+d"(d0), "+d"(d1) : "f"(b), "f"(c), "d"(db), "d"(dc));)
        }
        else if(MODE == 8) { // 7:1 mix FFMA + MUFU.SQRT
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                              "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; sqrt.approx.ftz.f32 %7, %7;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
        else if(MODE == 9) { // FMUL
            REP8(asm volatile("mul.rn.f32 %0, %0, %8; mul.rn.f32 %1, %1, %8; mul.rn.f32 %2, %2, %8; mul.rn.f32 %3, %3, %8;"
                              "mul.rn.f32 %4, %4, %8; mul.rn.f32 %5, %5, %8; mul.rn.f32 %6, %6, %8; mul.rn.f32 %7, %7, %8;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b));)
        }
        else if(MODE == 10) { // 3:1 mix FFMA + MUFU
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; sqrt.approx.ftz.f32 %3, %3;"
                              "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; sqrt.approx.ftz.f32 %7, %7;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(d0 + d1 + d2 + d3) + (float)(u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7);
    if(r == 123.456f) {
        out[0] = r;
    }
}

template<int MODE>
double run(char const* name, int sms, int blocks_per_sm, int iters, double clock_ghz_hint)
{
    float* out;
    cudaMalloc(&out, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int const grid = sms * blocks_per_sm;
    rate_kernel<MODE><<<grid, 256>>>(out, iters / 8, 1.0f, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for(int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        rate_kernel<MODE><<<grid, 256>>>(out, iters, 1.0f, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if(ms < best) best = ms;
    }
    double const warp_instr = (double)grid * 8.0 * (double)iters * 64.0; // 8 warps/block, 64 instr per REP8-of-8
    double const per_s = warp_instr / (best * 1e-3);
    double const per_clk_sm = per_s / (clock_ghz_hint * 1e9) / sms;
    printf("%-34s %8.3f ms  %9.3f Gwarp-instr/s  %6.3f warp-instr/clk/SM @%.3f GHz  (thread-ops %.2f T/s)\n", name, best,
           per_s * 1e-9, per_clk_sm, clock_ghz_hint, per_s * 32 * 1e-12);
    cudaFree(out);
    return per_s;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double const ghz = clk_khz * 1e-6;
    int const sms = p.multiProcessorCount;
    int const iters = argc > 1 ? atoi(argv[1]) : 4096;
    printf("device %s, %d SMs, max clock %.3f GHz; 8 blocks x 256 threads per SM\n", p.name, sms, ghz);
    float h[64];
    for(int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-7f;
    cudaMemcpyToSymbol(c_k, h, sizeof(h));
    double const ffma = run<0>("FFMA reg,reg,reg", sms, 8, iters, ghz);
    run<1>("FFMA reg,const,reg (dependent ring)", sms, 8, iters, ghz);
    run<2>("FADD", sms, 8, iters, ghz);
    run<9>("FMUL", sms, 8, iters, ghz);
    run<3>("MUFU.SQRT", sms, 8, iters / 4, ghz);
    run<4>("DFMA", sms, 8, iters / 2, ghz);
    run<5>("VIMNMX.U32 (alu pipe)", sms, 8, iters, ghz);
    run<6>("FFMA:VIMNMX 1:1", sms, 8, iters, ghz);
    run<7>("FFMA:DFMA 3:1", sms, 8, iters / 2, ghz);
    run<8>("FFMA:MUFU 7:1", sms, 8, iters, ghz);
    run<10>("FFMA:MUFU 3:1", sms, 8, iters / 2, ghz);
    printf("FP32 peak measured: %.2f TFLOP/s (FFMA = 2 flop x 32 lanes)\n", ffma * 64 * 1e-12);
    printf("JSON {\"fp32_ffma_tflops\": %.3f, \"warp_instr_per_s\": %.4e, \"sms\": %d, \"max_clock_ghz\": %.3f}\n", ffma * 64 * 1e-12,
           ffma, sms, ghz);
    return 0;
}
