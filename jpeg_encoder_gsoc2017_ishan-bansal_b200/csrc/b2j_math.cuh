// Device arithmetic of the encode path: colour conversion, the model's float64
// DCT flow graph, quantisation.  Everything here must reproduce the reference
// software model (/root/reference/test/model/enc_frame.py) bit for bit.
//
// * FP64: only __dadd_rn/__dsub_rn/__dmul_rn are used for the model's float
//   operations, so nothing can be contracted into an FMA (the library is also
//   built with -fmad=false).
// * double -> int uses an exact magic-number add on the FP64 pipe (one DADD.RZ gives
//   floor(|v|/8)); int -> double uses I2F.F64 (see i2d below); pipe rates in
//   profiles/r01/microbench_pipe_tput2.txt.
#pragma once
#include <stdint.h>

namespace b2j {

// math.cos((n*pi)/16), math.sin((n*pi)/16), pi = 3.14159265358979323846, and
// sqrt2 = 1.41421356237309504880, as CPython evaluates enc_frame.py:38-44.
// tests/test_oracle.py checks these literals against math.cos/math.sin.
#define B2J_C1 0x1.f6297cff75cb0p-1
#define B2J_S1 0x1.8f8b83c69a60ap-3
#define B2J_C3 0x1.a9b66290ea1a3p-1
#define B2J_S3 0x1.1c73b39ae68c8p-1
#define B2J_C6 0x1.87de2a6aea964p-2
#define B2J_S6 0x1.d906bcf328d46p-1
#define B2J_SQRT2 0x1.6a09e667f3bcdp+0

// ---------------------------------------------------------------- colour (enc_frame.py:27-35)
// Fixed-point restatement, exact for all 2^24 inputs except flagged candidates
// (tests/test_kernel_identities.py re-derives this from the constants below):
//   S_y  =  5016388 r + 9848226 g + 1912603 b                  (ceil(c * 2^24))
//   S_cb = -2830316 r - 5558291 g + 8388608 b + 128 * 2^24     (floor(c * 2^24))
//   S_cr =  8388608 r - 7024620 g - 1363987 b + 128 * 2^24
// component = S >> 24.  The float64 model differs from the exact rational value
// only where that value is an integer, i.e. where (S & 0x00FFFC00) == 0:
//   Y : 16774 candidate inputs, 3464 real differences -> re-evaluated in FP64;
//   Cb: no differences;
//   Cr: candidates are exactly g == b with r - g even; 688 real differences,
//       all listed by cr_exception() below.
#define B2J_CONST_Y_R 5016388u
#define B2J_CONST_Y_G 9848226u
#define B2J_CONST_Y_B 1912603u
#define B2J_CONST_CB_R 2830316u
#define B2J_CONST_CB_G 5558291u
#define B2J_CONST_CX_HALF 8388608u
#define B2J_CONST_CR_G 7024620u
#define B2J_CONST_CR_B 1363987u
#define B2J_CONST_C_OFFSET 0x80000000u
#define B2J_CONST_EXC_MASK 0x00FFFC00u

__device__ __forceinline__ uint32_t csc_sum_y(uint32_t r, uint32_t g, uint32_t b)
{
    return r * B2J_CONST_Y_R + g * B2J_CONST_Y_G + b * B2J_CONST_Y_B;
}
__device__ __forceinline__ uint32_t csc_sum_cb(uint32_t r, uint32_t g, uint32_t b)
{
    return b * B2J_CONST_CX_HALF - r * B2J_CONST_CB_R - g * B2J_CONST_CB_G + B2J_CONST_C_OFFSET;
}
__device__ __forceinline__ uint32_t csc_sum_cr(uint32_t r, uint32_t g, uint32_t b)
{
    return r * B2J_CONST_CX_HALF - g * B2J_CONST_CR_G - b * B2J_CONST_CR_B + B2J_CONST_C_OFFSET;
}

// Where the float64 model yields Y one below the exact value: 3464 of the 16774 candidates.
// Every (r, g) has at most one candidate b, so one bit per (r, g) decides (gen_y_exceptions.py
// evaluates the model's own expression over all 2^24 inputs); 8 KB, read through L1 on the
// candidate path only.  Candidates are rare in noisy content but include every grey pixel
// (r = g = b), so the path has to stay cheap: a table read, not a float64 re-evaluation.
static __device__ const uint32_t kYExceptionBits[2048] = {
#include "b2j_y_exceptions.inc"
};
// Both exception sets, two bits per (r, g): bit 0 as above, bit 1 = the Cr candidate of (r, g) (b = g, r - g even) is a real
// exception (the 688 inputs cr_exception() lists, here generated from the model's expression).  The pair variant of K1
// keeps this table in shared memory: one lookup patches both components of a candidate pixel, which is what every pixel of
// a grey image is.
static __device__ const uint32_t kYcExceptionBits[4096] = {
#include "b2j_yc_exceptions.inc"
};
__device__ __forceinline__ uint32_t y_exception(uint32_t r, uint32_t g)
{
    const uint32_t idx = r * 256u + g;
    return (__ldg(kYExceptionBits + (idx >> 5)) >> (idx & 31u)) & 1u;
}

// Where the float64 model yields Cr one below the exact value.  Only candidates
// (g == b, r - g even) can differ, and only at eight grey levels g:
//   class 1 (g = 201, 213, 225, 237, 249): every candidate with r <= g;
//   class 2 (g = 164, 176, 188): candidates with r <= g - 128 or g - 30 <= r <= g.
// That is the complete list of the 688 inputs (tests/test_kernel_identities.py).
__device__ __forceinline__ uint32_t cr_exception_class(uint32_t g)
{
    if (g == 201u || g == 213u || g == 225u || g == 237u || g == 249u) return 1u;
    if (g == 164u || g == 176u || g == 188u) return 2u;
    return 0u;
}
__device__ __forceinline__ uint32_t cr_exception_apply(uint32_t cls, uint32_t r, uint32_t g)
{
    if (r > g) return 0u;
    return cls == 1u ? 1u : ((r + 128u <= g || r + 30u >= g) ? 1u : 0u);
}
__device__ __forceinline__ uint32_t cr_exception(uint32_t r, uint32_t g)
{
    const uint32_t cls = cr_exception_class(g);
    return cls ? cr_exception_apply(cls, r, g) : 0u;
}

// Full per-pixel conversion; returns Y | Cb << 8 | Cr << 16.
__device__ __forceinline__ uint32_t rgb2ycbcr_px(uint32_t r, uint32_t g, uint32_t b)
{
    uint32_t sy = csc_sum_y(r, g, b), scb = csc_sum_cb(r, g, b), scr = csc_sum_cr(r, g, b);
    uint32_t y = sy >> 24, cb = scb >> 24, cr = scr >> 24;
    if ((sy & B2J_CONST_EXC_MASK) == 0u) y -= y_exception(r, g);
    if ((scr & B2J_CONST_EXC_MASK) == 0u) cr -= cr_exception(r, g);
    return y | (cb << 8) | (cr << 16);
}

// Gateware colour model, litejpeg/core/csc.py:85-160 with coef_w = 8:
// ca = 46, cb = 15, cc = 166, cd = 141, yoffset 16, coffset 128, saturate 0..255.
__device__ __forceinline__ uint32_t rgb2ycbcr_xapp930_px(int r, int g, int b)
{
    int yraw = ((46 * (r - g) + 15 * (b - g)) >> 8) + g;       // csc.py:87-115
    int y = yraw + 16;                                          // :144-152
    int cb = ((166 * (b - yraw)) >> 8) + 128;
    int cr = ((141 * (r - yraw)) >> 8) + 128;
    y = min(max(y, 0), 255); cb = min(max(cb, 0), 255); cr = min(max(cr, 0), 255);   // :156-160
    return (uint32_t)y | ((uint32_t)cb << 8) | ((uint32_t)cr << 16);
}

// ---------------------------------------------------------------- conversions
// exact (double)x.  I2F.F64.S32 runs on the XU pipe (16 lanes/clk/SM) but is ONE instruction; the
// magic-number form (2^52 + 2^31 + x) - (2^52 + 2^31) stays on the 4x faster FP64 pipe but costs three
// issue slots (LOP3 + MOV of the high word + DADD).  K1 is issue-bound and the XU pipe idle, so the
// hardware conversion wins: 1.556 -> 1.497 ms (profiles/scripts/k1_variants.py).
__device__ __forceinline__ double i2d(int x)
{
#ifdef B2J_I2D_MAGIC
    return __dadd_rn(__hiloint2double(0x43300000, x ^ (int)0x80000000), -4503601774854144.0);
#else
    return __int2double_rn(x);
#endif
}
// floor(|v| / 8) for |v| < 2^34: adding 2^55 leaves an ulp of 8, RZ truncates.
__device__ __forceinline__ uint32_t abs_div8_trunc(double v)
{
    return (uint32_t)__double2loint(__dadd_rz(fabs(v), 36028797018963968.0));
}

// ---------------------------------------------------------------- DCT (enc_frame.py:42-109)
__device__ __forceinline__ void rot(double i0, double i1, double c, double s, double &o0, double &o1)
{
    // enc_frame.py:42-47 with k == 1 (1*x is exact)
    o0 = __dadd_rn(__dmul_rn(i0, c), __dmul_rn(i1, s));
    o1 = __dsub_rn(__dmul_rn(i1, c), __dmul_rn(i0, s));
}

// dct_1d (enc_frame.py:54-84) on 8 doubles, in place: v[k] becomes r[k].
__device__ __forceinline__ void dct_1d(double (&v)[8])
{
    // stage 1
    double a3 = __dadd_rn(v[3], v[4]), a4 = __dsub_rn(v[3], v[4]);
    double a2 = __dadd_rn(v[2], v[5]), a5 = __dsub_rn(v[2], v[5]);
    double a1 = __dadd_rn(v[1], v[6]), a6 = __dsub_rn(v[1], v[6]);
    double a0 = __dadd_rn(v[0], v[7]), a7 = __dsub_rn(v[0], v[7]);
    // stage 2
    double b0 = __dadd_rn(a0, a3), b3 = __dsub_rn(a0, a3);
    double b1 = __dadd_rn(a1, a2), b2 = __dsub_rn(a1, a2);
    double b4, b7, b5, b6;
    rot(a4, a7, B2J_C3, B2J_S3, b4, b7);
    rot(a5, a6, B2J_C1, B2J_S1, b5, b6);
    // stage 3
    double e7 = __dadd_rn(b7, b5), e5 = __dsub_rn(b7, b5);
    double e4 = __dadd_rn(b4, b6), e6 = __dsub_rn(b4, b6);
    double f2, f3;
    rot(b2, b3, B2J_C6, B2J_S6, f2, f3);
    // stage 4
    v[0] = __dadd_rn(b0, b1);
    v[4] = __dsub_rn(b0, b1);
    v[2] = __dmul_rn(B2J_SQRT2, f2);
    v[6] = __dmul_rn(B2J_SQRT2, f3);
    v[1] = __dadd_rn(e7, e4);
    v[7] = __dsub_rn(e7, e4);
    v[5] = __dmul_rn(B2J_SQRT2, e6);
    v[3] = __dmul_rn(B2J_SQRT2, e5);
}

// dct_1d on 8 integers (|x| < 2^20): the butterflies of integers are exact in
// int32, so only the rotations run in FP64.  r0 and r4 stay integers.
__device__ __forceinline__ void dct_1d_int(const int (&x)[8], int &r0, int &r4,
                                           double &r1, double &r2, double &r3,
                                           double &r5, double &r6, double &r7)
{
    int a3 = x[3] + x[4], a4 = x[3] - x[4];
    int a2 = x[2] + x[5], a5 = x[2] - x[5];
    int a1 = x[1] + x[6], a6 = x[1] - x[6];
    int a0 = x[0] + x[7], a7 = x[0] - x[7];
    int b0 = a0 + a3, b3 = a0 - a3;
    int b1 = a1 + a2, b2 = a1 - a2;
    double b4, b7, b5, b6;
    rot(i2d(a4), i2d(a7), B2J_C3, B2J_S3, b4, b7);
    rot(i2d(a5), i2d(a6), B2J_C1, B2J_S1, b5, b6);
    double e7 = __dadd_rn(b7, b5), e5 = __dsub_rn(b7, b5);
    double e4 = __dadd_rn(b4, b6), e6 = __dsub_rn(b4, b6);
    double f2, f3;
    rot(i2d(b2), i2d(b3), B2J_C6, B2J_S6, f2, f3);
    r0 = b0 + b1;
    r4 = b0 - b1;
    r2 = __dmul_rn(B2J_SQRT2, f2);
    r6 = __dmul_rn(B2J_SQRT2, f3);
    r1 = __dadd_rn(e7, e4);
    r7 = __dsub_rn(e7, e4);
    r5 = __dmul_rn(B2J_SQRT2, e6);
    r3 = __dmul_rn(B2J_SQRT2, e5);
}

// exact (double)(int16) of the low / high half of a word: one I2F.F64.S16 with a half-word selector
__device__ __forceinline__ double half_lo_to_double(uint32_t w)
{
    double d;
    asm("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f64.s16 %0, lo;}" : "=d"(d) : "r"(w));
    return d;
}
__device__ __forceinline__ double half_hi_to_double(uint32_t w)
{
    double d;
    asm("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f64.s16 %0, hi;}" : "=d"(d) : "r"(w));
    return d;
}

// The same row transform straight from the eight sample BYTES of a row (px.x = samples 0..3, px.y = 4..7,
// unsigned, the -128 level shift folded in): the integer butterflies of stages 1 and 2 run on pairs of 16-bit
// fields (sums stay below 2^16; differences are two's complement halves whose low-half borrow is undone by
// adding 0x8000 before the high half is read), and the int -> double conversions read the halves directly
// (I2F.F64.S16 with a half-word selector).  17 integer instructions instead of 23; same values as dct_1d_int.
__device__ __forceinline__ void dct_1d_bytes(uint2 px, int &r0, int &r4, double &r1, double &r2, double &r3,
                                             double &r5, double &r6, double &r7)
{
    const uint32_t A1 = __byte_perm(px.x, 0u, 0x4140), A2 = __byte_perm(px.x, 0u, 0x4243);   // (x0, x1), (x3, x2)
    const uint32_t B1 = __byte_perm(px.y, 0u, 0x4243), B2 = __byte_perm(px.y, 0u, 0x4140);   // (x7, x6), (x4, x5)
    const uint32_t S1 = A1 + B1, S2 = A2 + B2;          // (a0, a1) + 256 each, (a3, a2) + 256 each
    const uint32_t D1 = A1 - B1, D2 = A2 - B2;          // (a7, a6), (a4, a5)
    const uint32_t T = S1 + S2, U = S1 - S2;            // (b0, b1) + 512 each, (b3, b2)
    const uint32_t D1h = D1 + 0x8000u, D2h = D2 + 0x8000u, Uh = U + 0x8000u;
    const int b0 = (int)(T & 0xFFFFu), b1 = (int)(T >> 16);
    double b4, b7, b5, b6;
    rot(half_lo_to_double(D2), half_lo_to_double(D1), B2J_C3, B2J_S3, b4, b7);          // a4, a7
    rot(half_hi_to_double(D2h), half_hi_to_double(D1h), B2J_C1, B2J_S1, b5, b6);        // a5, a6
    double e7 = __dadd_rn(b7, b5), e5 = __dsub_rn(b7, b5);
    double e4 = __dadd_rn(b4, b6), e6 = __dsub_rn(b4, b6);
    double f2, f3;
    rot(half_hi_to_double(Uh), half_lo_to_double(U), B2J_C6, B2J_S6, f2, f3);           // b2, b3
    r0 = b0 + b1 - 1024;
    r4 = b0 - b1;
    r2 = __dmul_rn(B2J_SQRT2, f2);
    r6 = __dmul_rn(B2J_SQRT2, f3);
    r1 = __dadd_rn(e7, e4);
    r7 = __dsub_rn(e7, e4);
    r5 = __dmul_rn(B2J_SQRT2, e6);
    r3 = __dmul_rn(B2J_SQRT2, e5);
}

// [ext] "fixed" DCT mode: one pass of the gateware datapath
// litejpeg/core/new_dct6.py:171-396 - its constants and shift points (:219-230, :345, :370),
// every product floor(c * x / 2^s) - with the two wiring slips of the source (:262-263, :294)
// corrected so that the graph is the model's, and 32-bit sums instead of wrapping 12-bit ones.
// Not parity with the float64 model (<= 4 LSB away); oracle: oracle/gateware.py dct_1d_fixed.
__device__ __forceinline__ void dct_1d_fixed(int (&v)[8])
{
#define B2J_BF(i, j) do { const int a_ = v[i] + v[j], b_ = v[i] - v[j]; v[i] = a_; v[j] = b_; } while (0)
    B2J_BF(3, 4); B2J_BF(2, 5); B2J_BF(1, 6); B2J_BF(0, 7);
    B2J_BF(0, 3); B2J_BF(1, 2);
    {
        const int v4 = v[4], v7 = v[7], v5 = v[5], v6 = v[6];
        v[4] = ((213 * v4) >> 8) + ((569 * v7) >> 10);
        v[7] = ((213 * v7) >> 8) - ((569 * v4) >> 10);
        v[5] = ((1004 * v5) >> 10) + ((25 * v6) >> 7);
        v[6] = ((1004 * v6) >> 10) - ((25 * v5) >> 7);
    }
    B2J_BF(7, 5); B2J_BF(4, 6); B2J_BF(0, 1);
    {
        const int fa = ((49 * v[2]) >> 7) + ((473 * v[3]) >> 9);
        const int fs = ((49 * v[3]) >> 7) - ((473 * v[2]) >> 9);
        v[2] = (181 * fa) >> 7;
        v[3] = (181 * fs) >> 7;
    }
    const int r1 = v[7] + v[4], r7 = v[7] - v[4];
    const int r5 = (181 * v[6]) >> 7, r3 = (181 * v[5]) >> 7;
    const int r6 = v[3], r2 = v[2], r4 = v[1], r0 = v[0];
    v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3; v[4] = r4; v[5] = r5; v[6] = r6; v[7] = r7;
#undef B2J_BF
}

// new_dct6.py:107-145 with dct_1d_fixed: -128, rows, columns, arithmetic >> 3.
// Emit(natural_index, |coefficient|, negative?) like dct_block below.
template <class Emit>
__device__ __forceinline__ void dct_block_fixed(const uint2 (&px)[8], Emit &&emit)
{
    int m[8][8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        int s[8];
        s[0] = (int)__byte_perm(px[x].x, 0u, 0x4440) - 128; s[1] = (int)__byte_perm(px[x].x, 0u, 0x4441) - 128;
        s[2] = (int)__byte_perm(px[x].x, 0u, 0x4442) - 128; s[3] = (int)__byte_perm(px[x].x, 0u, 0x4443) - 128;
        s[4] = (int)__byte_perm(px[x].y, 0u, 0x4440) - 128; s[5] = (int)__byte_perm(px[x].y, 0u, 0x4441) - 128;
        s[6] = (int)__byte_perm(px[x].y, 0u, 0x4442) - 128; s[7] = (int)__byte_perm(px[x].y, 0u, 0x4443) - 128;
        dct_1d_fixed(s);
#pragma unroll
        for (int u = 0; u < 8; u++) m[u][x] = s[u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
        dct_1d_fixed(m[u]);
#pragma unroll
        for (int vv = 0; vv < 8; vv++) {
            const int o = m[u][vv] >> 3;
            emit(vv * 8 + u, (uint32_t)abs(o), o < 0);
        }
    }
}

// One 8x8 block held by one thread.  px[row] = the 8 samples of a row as two
// little-endian words.  Emit(natural_index, magnitude of int(v/8), negative?)
// is called for all 64 coefficients.
template <class Emit>
__device__ __forceinline__ void dct_block(const uint2 (&px)[8], Emit &&emit)
{
    int m0[8], m4[8];             // integer columns 0 and 4 after the row pass
    double m[6][8];               // columns 1,2,3,5,6,7
#pragma unroll
    for (int x = 0; x < 8; x++) {
        int s[8];
        s[0] = (int)(px[x].x & 0xFF) - 128; s[1] = (int)((px[x].x >> 8) & 0xFF) - 128;
        s[2] = (int)((px[x].x >> 16) & 0xFF) - 128; s[3] = (int)(px[x].x >> 24) - 128;
        s[4] = (int)(px[x].y & 0xFF) - 128; s[5] = (int)((px[x].y >> 8) & 0xFF) - 128;
        s[6] = (int)((px[x].y >> 16) & 0xFF) - 128; s[7] = (int)(px[x].y >> 24) - 128;
        dct_1d_int(s, m0[x], m4[x], m[0][x], m[1][x], m[2][x], m[3][x], m[4][x], m[5][x]);
    }
    // column pass; column u of the matrix is row-pass output index u
    {
        int r0, r4; double r1, r2, r3, r5, r6, r7;
        dct_1d_int(m0, r0, r4, r1, r2, r3, r5, r6, r7);
        emit(0 * 8 + 0, (uint32_t)abs(r0) >> 3, r0 < 0);
        emit(4 * 8 + 0, (uint32_t)abs(r4) >> 3, r4 < 0);
        emit(1 * 8 + 0, abs_div8_trunc(r1), __double2hiint(r1) < 0);
        emit(2 * 8 + 0, abs_div8_trunc(r2), __double2hiint(r2) < 0);
        emit(3 * 8 + 0, abs_div8_trunc(r3), __double2hiint(r3) < 0);
        emit(5 * 8 + 0, abs_div8_trunc(r5), __double2hiint(r5) < 0);
        emit(6 * 8 + 0, abs_div8_trunc(r6), __double2hiint(r6) < 0);
        emit(7 * 8 + 0, abs_div8_trunc(r7), __double2hiint(r7) < 0);
        dct_1d_int(m4, r0, r4, r1, r2, r3, r5, r6, r7);
        emit(0 * 8 + 4, (uint32_t)abs(r0) >> 3, r0 < 0);
        emit(4 * 8 + 4, (uint32_t)abs(r4) >> 3, r4 < 0);
        emit(1 * 8 + 4, abs_div8_trunc(r1), __double2hiint(r1) < 0);
        emit(2 * 8 + 4, abs_div8_trunc(r2), __double2hiint(r2) < 0);
        emit(3 * 8 + 4, abs_div8_trunc(r3), __double2hiint(r3) < 0);
        emit(5 * 8 + 4, abs_div8_trunc(r5), __double2hiint(r5) < 0);
        emit(6 * 8 + 4, abs_div8_trunc(r6), __double2hiint(r6) < 0);
        emit(7 * 8 + 4, abs_div8_trunc(r7), __double2hiint(r7) < 0);
    }
    constexpr int kCol[6] = {1, 2, 3, 5, 6, 7};
#pragma unroll
    for (int j = 0; j < 6; j++) {
        dct_1d(m[j]);
#pragma unroll
        for (int x = 0; x < 8; x++)
            emit(x * 8 + kCol[j], abs_div8_trunc(m[j][x]), __double2hiint(m[j][x]) < 0);
    }
}

// round-half-away-from-zero of a / d for a <= 1024 (enc_frame.py:111-129);
// mul = ceil(2^20 / d).
__device__ __forceinline__ uint32_t quant_mag(uint32_t a, uint32_t mul)
{
    return (a * mul + (1u << 19)) >> 20;
}

// exact for any |v| <= 32767, d in 1..255 (stage operator, not the fused path)
__device__ __forceinline__ int quant_round(int v, int d)
{
    int a = abs(v);
    a = (2 * a + d) / (2 * d);
    return v < 0 ? -a : a;
}
__device__ __forceinline__ int quant_trunc(int v, int d)      // test/model/encoder.py:84-88
{
    int a = abs(v) / d;
    return v < 0 ? -a : a;
}
__device__ __forceinline__ int quant_reciprocal(int v, int d)  // core/quantization.py:58,124-168
{
    uint32_t inv = 65536u / (uint32_t)d;
    uint32_t m = (uint32_t)abs(v) * inv;
    int mag = (int)(((m >> 16) & 0xFFFu) + ((m >> 15) & 1u));
    return v < 0 ? -mag : mag;
}

__device__ __forceinline__ int bit_size(int v)                 // rle/entropycoder.py:55-65
{
    return 32 - __clz(abs(v));
}

}  // namespace b2j
