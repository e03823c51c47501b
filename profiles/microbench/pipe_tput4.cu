// v3 (round 2: packed FP32 pairs added): one CTA per SM, NW warps, inline PTX so each variant is a clean instruction
// stream; reports warp-instructions issued per clock per SM (max 4.0) and the
// lane-ops per clock per SM of the FIRST instruction class named.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_tput2 pipe_tput2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 1024
#define U 8

template <int OP>
__global__ void __launch_bounds__(1024) bench(long long *cycles, double *sink, int seed)
{
    double d[U]; unsigned x[U], y[U]; unsigned long long w[U];
#pragma unroll
    for (int i = 0; i < U; i++) {
        d[i] = 1.0 + (threadIdx.x + i * 7 + seed) * 1e-3; x[i] = threadIdx.x * 3 + i + seed; y[i] = x[i] * 2654435761u;
        w[i] = (unsigned long long)x[i] * 77;
    }
    double k1 = 0.9999 + seed * 1e-9, k2 = 1e-7 * (seed + 1);
    unsigned m1 = 429497 + seed, m2 = seed + 3;
    unsigned long long wk1 = 0x3F7FF0003F7FF000ull + seed, wk2 = 0x3400000034000000ull + seed;
    __shared__ unsigned short sbuf[1024 * U]; unsigned sa = (unsigned)__cvta_generic_to_shared(sbuf + threadIdx.x * U);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            if (OP == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));
            if (OP == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k1));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));
            if (OP == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(m1));
            if (OP == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m1));
            if (OP == 5) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));
            if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x3254;" : "+r"(x[i]) : "r"(m1));
            if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m1), "r"(m2));
            if (OP == 8) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));
            if (OP == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m1));
            if (OP == 10) asm volatile("bfind.u32 %0, %0;" : "+r"(x[i]));
            if (OP == 11) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));       // DADD + IMAD
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2)); }
            if (OP == 12) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));       // DADD + LOP3
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m1), "r"(m2)); }
            if (OP == 13) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));   // IMAD + LOP3
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 14) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));       // DADD + IMAD + LOP3
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 15) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));       // DADD + 2 IMAD + 2 LOP3
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2));
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m2), "r"(m1));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 16) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(m1)); // WIDE + LOP3
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 17) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));
            if (OP == 18) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
            if (OP == 19) { asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2));  // IDP + LOP3
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 20) { asm volatile("prmt.b32 %0, %0, %1, 0x3254;" : "+r"(x[i]) : "r"(m1));            // PRMT + IMAD
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 21) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&x[i]) : "f"(0.999f), "f"(1e-7f)); }
            if (OP == 30) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
            if (OP == 31) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wk2));
            if (OP == 32) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 33) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 34) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&x[i]) : "f"(0.999f), "f"(1e-7f)); }
            if (OP == 35) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2));
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m1), "r"(m2)); }
            if (OP == 36) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(m1));
            if (OP == 37) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2)); }
            if (OP == 38) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2)); }
            if (OP == 39) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *)&x[i]) : "f"(1e-7f));
            if (OP == 40) { asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(*(float *)&y[i]) : "h"((unsigned short)x[i])); x[i] += y[i]; }
            if (OP == 41) { asm volatile("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.u8 %0, b1;}" : "=f"(*(float *)&y[i]) : "r"(x[i])); }
            if (OP == 42) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m1), "r"(m2));
                            asm volatile("prmt.b32 %0, %0, %1, 0x3254;" : "+r"(x[i]) : "r"(m1)); }
            if (OP == 43) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wk1), "l"(wk2));
                            asm volatile("st.shared.u16 [%0], %1;" :: "r"(sa + i * 2), "h"((unsigned short)x[i]) : "memory"); }
            if (OP == 50) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wk1));
            if (OP == 51) { asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wk1));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i])); }
            if (OP == 52) { asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wk1));
                            asm volatile("mad.lo.u32 %0, %0, 77777, 12345;" : "+r"(y[i])); }
            if (OP == 53) { asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wk1));
                            asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, 0f34000000;" : "+f"(*(float *)&y[i])); }
            if (OP == 54) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wk2));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i])); }
            if (OP == 55) { asm volatile("mad.lo.u32 %0, %0, 77777, 12345;" : "+r"(x[i]));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i])); }
            if (OP == 56) { asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, 0f34000000;" : "+f"(*(float *)&x[i]));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i])); }
            if (OP == 57) { asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, 0f34000000;" : "+f"(*(float *)&x[i]));
                            asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, 0f34000000;" : "+f"(*(float *)&y[i]));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(*(unsigned *)&w[i])); }
            if (OP == 58) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i]));
                            asm volatile("mad.lo.u32 %0, %0, 77777, 12345;" : "+r"(x[i])); }
            if (OP == 59) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(k2));
                            asm volatile("lop3.b32 %0, %0, 0x12345, 0x54321, 0x96;" : "+r"(y[i]));
                            asm volatile("mad.lo.u32 %0, %0, 77777, 12345;" : "+r"(x[i]));
                            asm volatile("prmt.b32 %0, %0, 0x4b000000, 0x7650;" : "+r"(*(unsigned *)&w[i])); }
            if (OP == 22) { asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x[i]) : "r"(m1), "r"(m2)); }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double acc = 0; unsigned long long wa = 0;
#pragma unroll
    for (int i = 0; i < U; i++) { acc += d[i] + x[i] + y[i]; wa += w[i]; }
    if (acc == 12345.678 && wa == 7) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int instrs_per_step)
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cyc; double *sink; cudaMalloc(&cyc, sizeof(long long) * sms); cudaMalloc(&sink, 8);
    for (int nw = 8; nw <= 32; nw *= 2) {
        bench<OP><<<sms, nw * 32>>>(cyc, sink, 1); cudaDeviceSynchronize();
        bench<OP><<<sms, nw * 32>>>(cyc, sink, 2); cudaDeviceSynchronize();
        long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
        double steps = (double)ITERS * U * nw;    // warp-steps per SM
        printf("%-30s warps/SM=%2d cycles=%9.0f  warp-instr/clk/SM=%5.2f  first-op lanes/clk/SM=%6.1f\n",
               name, nw, avg, steps * instrs_per_step / avg, steps * 32 / avg);
    }
    cudaFree(cyc); cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    if (0) { run<0>("DADD", 1); run<1>("DMUL", 1); run<17>("DADD.RZ", 1);
    run<2>("IMAD", 1); run<3>("IMAD.WIDE.U32 (64b acc)", 1); run<4>("IMAD.HI.U32", 1); run<5>("IDP.4A", 1);
    run<6>("PRMT", 1); run<7>("LOP3", 1); run<8>("SHF", 1); run<9>("IADD", 1); run<10>("FLO", 1); run<18>("POPC", 1);
    run<21>("FFMA", 1); run<22>("ISETP+SEL", 2);
    run<11>("DADD+IMAD", 2); run<12>("DADD+LOP3", 2); run<13>("IMAD+LOP3", 2); run<14>("DADD+IMAD+LOP3", 3);
    run<15>("DADD+2IMAD+2LOP3", 5); run<16>("IMAD.WIDE+LOP3", 2); run<19>("IDP.4A+LOP3", 2); run<20>("PRMT+IMAD", 2); }
    run<30>("FFMA2 3 pairs", 1); run<50>("FFMA2 2 pairs", 1);
    run<51>("FFMA2+LOP3imm", 2); run<52>("FFMA2+IMADimm", 2); run<53>("FFMA2+FFMAimm", 2); run<54>("FADD2+LOP3imm", 2);
    run<55>("IMADimm+LOP3imm", 2); run<56>("FFMAimm+LOP3imm", 2); run<57>("2FFMAimm+LOP3imm", 3);
    run<58>("DADD+LOP3imm+IMADimm", 3); run<59>("DADD+LOP3imm+IMADimm+PRMT", 4);
    return 0;
}
