// Instruction-throughput microbenchmark for the op mix the fused JPEG block kernel
// needs (FP64 add/mul without FMA, int<->double conversions, IMAD.WIDE, IDP.4A,
// PRMT, FLO ...).  Prints lane-ops per clock per SM for each op at full occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_tput pipe_tput.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
#define UNROLL 8

template <int OP>
__global__ void __launch_bounds__(256) bench(long long *cycles, double *sink, int seed)
{
    double d[UNROLL]; int x[UNROLL]; unsigned u[UNROLL]; long long w[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++) {
        d[i] = 1.0 + (threadIdx.x + i * 7 + seed) * 1e-3; x[i] = threadIdx.x * 3 + i + seed; u[i] = x[i] * 2654435761u;
        w[i] = (long long)x[i] * 77;
    }
    const double k1 = 0.9999 + seed * 1e-9, k2 = 1e-7 * (seed + 1);
    const int m1 = 429497 + seed, m2 = seed + 3;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (OP == 0) d[i] = __dadd_rn(d[i], k2);
            if (OP == 1) d[i] = __dmul_rn(d[i], k1);
            if (OP == 2) d[i] = __fma_rn(d[i], k1, k2);
            if (OP == 3) { d[i] = (double)x[i]; x[i] += __double2hiint(d[i]) & 3; }            // I2F.F64 (+2 int)
            if (OP == 4) { x[i] = __double2int_rz(d[i]); d[i] = __hiloint2double(0x40000000 + (x[i] & 1), x[i]); }  // F2I.F64 (+ int)
            if (OP == 5) { d[i] = __dadd_rn(__hiloint2double(0x43300000, x[i] ^ 0x80000000), -4503601774854144.0); x[i] += __double2hiint(d[i]) & 3; } // magic I2F
            if (OP == 6) { double t = __dadd_rz(fabs(d[i]), 36028797018963968.0); x[i] = __double2loint(t); d[i] = __hiloint2double(0x40000000 + (x[i] & 1), x[i]); } // magic F2I
            if (OP == 7) x[i] = x[i] * m1 + m2;                                                 // IMAD
            if (OP == 8) w[i] = (long long)x[i] * (long long)m1 + w[i], x[i] = (int)(w[i] >> 32);  // IMAD.WIDE
            if (OP == 9) u[i] = __umulhi(u[i], (unsigned)m1) + m2;                              // IMAD.HI
            if (OP == 10) x[i] = __dp4a(x[i], m1, x[i]);                                        // IDP.4A
            if (OP == 11) u[i] = __byte_perm(u[i], (unsigned)m1, 0x3254 + (u[i] & 1));          // PRMT (+lop)
            if (OP == 12) x[i] = (x[i] ^ m1) + m2;                                              // LOP3 + IADD
            if (OP == 13) x[i] = 31 - __clz(x[i] | 1) + x[i];                                   // FLO (+ int)
            if (OP == 14) x[i] = __funnelshift_r(x[i], m1, x[i] & 31);                          // SHF
            if (OP == 15) x[i] = (x[i] < m1) ? x[i] + m2 : x[i] - m2;                           // ISETP + SEL-ish
            if (OP == 16) x[i] = __popc(x[i]) + x[i];                                           // POPC
            if (OP == 17) { float f = __int_as_float(0x3f800000 | (x[i] & 0x7fffff)); f = fmaf(f, 0.999f, 1e-7f); x[i] = __float_as_int(f); } // FFMA
            if (OP == 18) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1) + 1;                     // SHFL
            if (OP == 19) x[i] = (int)__vabsdiffu4((unsigned)x[i], (unsigned)m1) + m2;         // video SIMD
        }
    }
    long long t1 = clock64();
    double acc = 0; long long wa = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++) { acc += d[i] + x[i] + u[i]; wa += w[i]; }
    if (acc == 12345.678 && wa == 7) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int extra_int_ops)
{
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long *cyc; double *sink; cudaMalloc(&cyc, sizeof(long long) * sms * 8); cudaMalloc(&sink, 8);
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
        int grid = sms * ctas_per_sm;
        bench<OP><<<grid, 256>>>(cyc, sink, 1); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); bench<OP><<<grid, 256>>>(cyc, sink, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
        double ops = (double)ITERS * UNROLL * 256 * ctas_per_sm;   // lane-ops per SM
        printf("%-28s warps/SM=%2d  cycles=%9.0f  lane-ops/clk/SM=%7.2f  (%.3f ms; +%d int ops per op)\n",
               name, 8 * ctas_per_sm, avg, ops / avg, ms, extra_int_ops);
    }
    cudaFree(cyc); cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    run<0>("DADD", 0); run<1>("DMUL", 0); run<2>("DFMA", 0);
    run<3>("I2F.F64.S32 (+2 int)", 2); run<4>("F2I.S32.F64.TRUNC (+3 int)", 3);
    run<5>("magic I2F: LOP+DADD (+2 int)", 2); run<6>("magic F2I: DADD.RZ (+3 int)", 3);
    run<7>("IMAD", 0); run<8>("IMAD.WIDE", 0); run<9>("IMAD.HI (+IADD)", 1); run<10>("IDP.4A", 0);
    run<11>("PRMT (+LOP,IADD)", 2); run<12>("LOP3+IADD", 1); run<13>("FLO (+3 int)", 3); run<14>("SHF (+LOP)", 1);
    run<15>("ISETP+2xIADD+SEL", 3); run<16>("POPC (+IADD)", 1); run<17>("FFMA (+2 LOP)", 2); run<18>("SHFL (+IADD)", 1);
    run<19>("VABSDIFF4 (+IADD)", 1);
    return 0;
}
