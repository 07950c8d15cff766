// issue_rates2.cu -- second round of issue-rate measurements on B200 (sm_100a):
// packed FP32 (FFMA2/FADD2/FMUL2, new on sm_100), individual ALU-pipe ops, and the
// cost of MUFU in an FFMA stream.  Same conventions as issue_rates.cu.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x

template<int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float const b = seed * 0.5f, c = seed * 0.25f;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7;
    uint64_t p0, p1, p2, p3, p4, p5, p6, p7, pb, pc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a1), "f"(a2));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a3), "f"(a4));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p4) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p5) : "f"(a5), "f"(a6));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p6) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p7) : "f"(a7), "f"(a0));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(c));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(b));
    unsigned const k = u0 ^ (unsigned)iters;
    for(int it = 0; it < iters; ++it) {
        if(MODE == 0) { // FFMA2
            REP8(asm volatile("fma.rn.f32x2 %0, %0, %8, %9; fma.rn.f32x2 %1, %1, %8, %9; fma.rn.f32x2 %2, %2, %8, %9; fma.rn.f32x2 %3, %3, %8, %9;"
                              "fma.rn.f32x2 %4, %4, %8, %9; fma.rn.f32x2 %5, %5, %8, %9; fma.rn.f32x2 %6, %6, %8, %9; fma.rn.f32x2 %7, %7, %8, %9;"
                              : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(pb), "l"(pc));)
        }
        else if(MODE == 1) { // FFMA2 : IADD 1:1
            REP8(asm volatile("fma.rn.f32x2 %0, %0, %8, %9; add.u32 %4, %4, %5; fma.rn.f32x2 %1, %1, %8, %9; add.u32 %5, %5, %6;"
                              "fma.rn.f32x2 %2, %2, %8, %9; add.u32 %6, %6, %7; fma.rn.f32x2 %3, %3, %8, %9; add.u32 %7, %7, %4;"
                              : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "l"(pb), "l"(pc), "r"(k));)
        }
        else if(MODE == 2) { // IADD (alu)
            REP8(asm volatile("add.u32 %0, %0, %1; add.u32 %1, %1, %2; add.u32 %2, %2, %3; add.u32 %3, %3, %4;"
                              "add.u32 %4, %4, %5; add.u32 %5, %5, %6; add.u32 %6, %6, %7; add.u32 %7, %7, %0;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(k));)
        }
        else if(MODE == 3) { // LOP3 xor
            REP8(asm volatile("and.b32 %0, %0, %1; or.b32 %1, %1, %2; xor.b32 %2, %2, %3; and.b32 %3, %3, %4;"
                              "or.b32 %4, %4, %5; xor.b32 %5, %5, %6; and.b32 %6, %6, %7; or.b32 %7, %7, %0;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(k));)
        }
        else if(MODE == 4) { // IMAD (fma pipe?)
            REP8(asm volatile("mad.lo.u32 %0, %0, %8, %8; mad.lo.u32 %1, %1, %8, %8; mad.lo.u32 %2, %2, %8, %8; mad.lo.u32 %3, %3, %8, %8;"
                              "mad.lo.u32 %4, %4, %8, %8; mad.lo.u32 %5, %5, %8, %8; mad.lo.u32 %6, %6, %8, %8; mad.lo.u32 %7, %7, %8, %8;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(k));)
        }
        else if(MODE == 5) { // SHF
            REP8(asm volatile("shf.r.wrap.b32 %0, %0, %1, %8; shf.r.wrap.b32 %1, %1, %2, %8; shf.r.wrap.b32 %2, %2, %3, %8; shf.r.wrap.b32 %3, %3, %4, %8;"
                              "shf.r.wrap.b32 %4, %4, %5, %8; shf.r.wrap.b32 %5, %5, %6, %8; shf.r.wrap.b32 %6, %6, %7, %8; shf.r.wrap.b32 %7, %7, %0, %8;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(k));)
        }
        else if(MODE == 6) { // min.u32 on independent regs
            REP8(asm volatile("min.u32 %0, %0, %1; max.u32 %1, %1, %2; min.u32 %2, %2, %3; max.u32 %3, %3, %4;"
                              "min.u32 %4, %4, %5; max.u32 %5, %5, %6; min.u32 %6, %6, %7; max.u32 %7, %7, %0;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(k + it));)
        }
        else if(MODE == 7) { // FFMA : MUFU.RCP 15:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                              "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %0, %0, %8, %9;"
                              "fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9; fma.rn.f32 %4, %4, %8, %9;"
                              "fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %0, %0, %8, %9; sqrt.approx.ftz.f32 %7, %7;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
        else if(MODE == 8) { // MUFU.RCP
            REP8(asm volatile("rcp.approx.ftz.f32 %0, %1; rsqrt.approx.ftz.f32 %1, %2; rcp.approx.ftz.f32 %2, %3; rsqrt.approx.ftz.f32 %3, %4;"
                              "rcp.approx.ftz.f32 %4, %5; rsqrt.approx.ftz.f32 %5, %6; rcp.approx.ftz.f32 %6, %7; rsqrt.approx.ftz.f32 %7, %0;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));)
        }
        else if(MODE == 9) { // setp + selp pairs (ISETP + SEL)
            REP8(asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %2, %3, q; setp.lt.u32 q, %1, %2; selp.u32 %1, %3, %0, q;"
                              "setp.lt.u32 q, %2, %3; selp.u32 %2, %0, %1, q; setp.lt.u32 q, %3, %0; selp.u32 %3, %1, %2, q; }"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3));)
        }
        else if(MODE == 10) { // FFMA2 : FFMA 1:1
            REP8(asm volatile("fma.rn.f32x2 %0, %0, %8, %9; fma.rn.f32 %4, %4, %10, %11; fma.rn.f32x2 %1, %1, %8, %9; fma.rn.f32 %5, %5, %10, %11;"
                              "fma.rn.f32x2 %2, %2, %8, %9; fma.rn.f32 %6, %6, %10, %11; fma.rn.f32x2 %3, %3, %8, %9; fma.rn.f32 %7, %7, %10, %11;"
                              : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "l"(pb), "l"(pc), "f"(b), "f"(c));)
        }
        else if(MODE == 11) { // FFMA2 : xor 1:2
            REP8(asm volatile("fma.rn.f32x2 %0, %0, %8, %9; and.b32 %4, %4, %5; or.b32 %5, %5, %6; fma.rn.f32x2 %1, %1, %8, %9; xor.b32 %6, %6, %7; and.b32 %7, %7, %4;"
                              "fma.rn.f32x2 %2, %2, %8, %9; or.b32 %4, %4, %5; xor.b32 %5, %5, %6; fma.rn.f32x2 %3, %3, %8, %9; and.b32 %6, %6, %7; or.b32 %7, %7, %4;"
                              : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "l"(pb), "l"(pc), "r"(k));)
        }
        else if(MODE == 12) { // FADD2
            REP8(asm volatile("add.rn.f32x2 %0, %0, %8; add.rn.f32x2 %1, %1, %8; add.rn.f32x2 %2, %2, %8; add.rn.f32x2 %3, %3, %8;"
                              "add.rn.f32x2 %4, %4, %8; add.rn.f32x2 %5, %5, %8; add.rn.f32x2 %6, %6, %8; add.rn.f32x2 %7, %7, %8;"
                              : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(pb));)
        }
        else if(MODE == 13) { // FFMA : xor 1:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; and.b32 %4, %4, %5; fma.rn.f32 %1, %1, %8, %9; or.b32 %5, %5, %6;"
                              "fma.rn.f32 %2, %2, %8, %9; xor.b32 %6, %6, %7; fma.rn.f32 %3, %3, %8, %9; and.b32 %7, %7, %4;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c), "r"(k));)
        }
        if(MODE == 14) { // FSETP + FSEL ring
            REP8(asm volatile("{ .reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %2, %3, q; setp.lt.f32 q, %1, %2; selp.f32 %1, %3, %0, q;"
                              "setp.lt.f32 q, %2, %3; selp.f32 %2, %0, %1, q; setp.lt.f32 q, %3, %0; selp.f32 %3, %1, %2, q; }"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3));)
        }
        else if(MODE == 15) { // FMNMX ring
            REP8(asm volatile("min.f32 %0, %0, %1; max.f32 %1, %1, %2; min.f32 %2, %2, %3; max.f32 %3, %3, %4;"
                              "min.f32 %4, %4, %5; max.f32 %5, %5, %6; min.f32 %6, %6, %7; max.f32 %7, %7, %0;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));)
        }
        else if(MODE == 16) { // FFMA : MUFU.SQRT 11:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                              "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %0, %0, %8, %9;"
                              "fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9; sqrt.approx.ftz.f32 %7, %7;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
        else if(MODE == 17) { // I2F + F2I? no: int->float mantissa trick cost: SHF + LOP3 + FADD
            REP8(asm volatile("{ .reg .b32 t; shr.u32 t, %0, 9; or.b32 t, t, 0x3f800000; mov.b32 %4, t; add.f32 %4, %4, %5; mov.b32 t, %4; xor.b32 %0, %0, t;"
                              "shr.u32 t, %1, 9; or.b32 t, t, 0x3f800000; mov.b32 %5, t; add.f32 %5, %5, %6; mov.b32 t, %5; xor.b32 %1, %1, t;"
                              "shr.u32 t, %2, 9; or.b32 t, t, 0x3f800000; mov.b32 %6, t; add.f32 %6, %6, %7; mov.b32 t, %6; xor.b32 %2, %2, t;"
                              "shr.u32 t, %3, 9; or.b32 t, t, 0x3f800000; mov.b32 %7, t; add.f32 %7, %7, %4; mov.b32 t, %7; xor.b32 %3, %3, t; }"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));)
        }
        if(MODE == 18) { // FFMA : VIMNMX 1:1 (min ring on u4..u7)
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; min.u32 %4, %4, %5; fma.rn.f32 %1, %1, %8, %9; max.u32 %5, %5, %6;"
                              "fma.rn.f32 %2, %2, %8, %9; min.u32 %6, %6, %7; fma.rn.f32 %3, %3, %8, %9; max.u32 %7, %7, %4;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c));)
        }
        else if(MODE == 19) { // FFMA : VIMNMX 3:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; min.u32 %4, %4, %5;"
                              "fma.rn.f32 %3, %3, %8, %9; fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; max.u32 %5, %5, %6;"
                              "fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9; fma.rn.f32 %0, %0, %8, %9; min.u32 %6, %6, %7;"
                              "fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9; max.u32 %7, %7, %4;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c));)
        }
        else if(MODE == 20) { // FFMA : IMAD 1:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; mad.lo.u32 %4, %4, %5, %6; fma.rn.f32 %1, %1, %8, %9; mad.lo.u32 %5, %5, %6, %7;"
                              "fma.rn.f32 %2, %2, %8, %9; mad.lo.u32 %6, %6, %7, %4; fma.rn.f32 %3, %3, %8, %9; mad.lo.u32 %7, %7, %4, %5;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c));)
        }
        else if(MODE == 21) { // VIMNMX : IMAD 1:1 (alu half-rate + fma half-rate)
            REP8(asm volatile("min.u32 %0, %0, %1; mad.lo.u32 %4, %4, %5, %6; max.u32 %1, %1, %2; mad.lo.u32 %5, %5, %6, %7;"
                              "min.u32 %2, %2, %3; mad.lo.u32 %6, %6, %7, %4; max.u32 %3, %3, %0; mad.lo.u32 %7, %7, %4, %5;"
                              : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7));)
        }
        else if(MODE == 22) { // FFMA : IADD 1:1
            REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; add.u32 %4, %4, %5; fma.rn.f32 %1, %1, %8, %9; add.u32 %5, %5, %6;"
                              "fma.rn.f32 %2, %2, %8, %9; add.u32 %6, %6, %7; fma.rn.f32 %3, %3, %8, %9; add.u32 %7, %7, %4;"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "f"(c));)
        }
        else if(MODE == 23) { // FFMA : FSETP+FSEL 2:2
            REP8(asm volatile("{ .reg .pred q; fma.rn.f32 %0, %0, %8, %9; setp.lt.f32 q, %4, %5; fma.rn.f32 %1, %1, %8, %9; selp.f32 %4, %6, %7, q;"
                              "fma.rn.f32 %2, %2, %8, %9; setp.lt.f32 q, %5, %6; fma.rn.f32 %3, %3, %8, %9; selp.f32 %5, %7, %4, q; }"
                              : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));)
        }
    }
    float2 q0, q1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(q0.x), "=f"(q0.y) : "l"(p0 ^ p1 ^ p2 ^ p3));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(q1.x), "=f"(q1.y) : "l"(p4 ^ p5 ^ p6 ^ p7));
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q0.x + q0.y + q1.x + q1.y + (float)(u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7);
    if(r == 123.456f) {
        out[0] = r;
    }
}

template<int MODE>
void run(char const* name, int sms, int iters, double ghz, int instr_per_rep8)
{
    float* out;
    cudaMalloc(&out, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int const grid = sms * 8;
    rate_kernel<MODE><<<grid, 256>>>(out, iters / 8, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for(int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        rate_kernel<MODE><<<grid, 256>>>(out, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if(ms < best) best = ms;
    }
    double const warp_instr = (double)grid * 8.0 * (double)iters * 8.0 * instr_per_rep8;
    double const per_s = warp_instr / (best * 1e-3);
    printf("%-28s %8.3f ms  %9.3f Gwarp-instr/s  %6.3f warp-instr/clk/SM @%.3f GHz\n", name, best, per_s * 1e-9,
           per_s / (ghz * 1e9) / sms, ghz);
    cudaFree(out);
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
    printf("device %s, %d SMs, max clock %.3f GHz\n", p.name, sms, ghz);
    run<0>("FFMA2", sms, iters, ghz, 8);
    run<12>("FADD2", sms, iters, ghz, 8);
    run<10>("FFMA2:FFMA 1:1", sms, iters, ghz, 8);
    run<1>("FFMA2:IADD 1:1", sms, iters, ghz, 8);
    run<11>("FFMA2:LOP3 1:2", sms, iters, ghz, 12);
    run<13>("FFMA:LOP3 1:1", sms, iters, ghz, 8);
    run<2>("IADD", sms, iters, ghz, 8);
    run<3>("LOP3", sms, iters, ghz, 8);
    run<4>("IMAD", sms, iters, ghz, 8);
    run<5>("SHF", sms, iters, ghz, 8);
    run<6>("VIMNMX", sms, iters, ghz, 8);
    run<9>("ISETP+SEL", sms, iters, ghz, 8);
    run<8>("MUFU.RCP/RSQ", sms, iters / 4, ghz, 8);
    run<7>("FFMA:MUFU.SQRT 15:1", sms, iters / 2, ghz, 16);
    run<16>("FFMA:MUFU.SQRT 11:1", sms, iters / 2, ghz, 12);
    run<14>("FSETP+FSEL", sms, iters, ghz, 8);
    run<15>("FMNMX", sms, iters, ghz, 8);
    run<17>("SHF+LOP3+FADD+LOP3 (x4)", sms, iters, ghz, 16);
    run<18>("FFMA:VIMNMX 1:1", sms, iters, ghz, 8);
    run<19>("FFMA:VIMNMX 3:1", sms, iters / 2, ghz, 16);
    run<20>("FFMA:IMAD 1:1", sms, iters, ghz, 8);
    run<21>("VIMNMX:IMAD 1:1", sms, iters, ghz, 8);
    run<22>("FFMA:IADD 1:1", sms, iters, ghz, 8);
    run<23>("FFMA:(FSETP,FSEL) 1:1", sms, iters, ghz, 8);
    return 0;
}
