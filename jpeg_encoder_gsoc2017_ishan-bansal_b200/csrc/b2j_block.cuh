// The per-block pipeline of the one-thread-per-block kernels of b2j_fused.cu (simple and tiled variants):
// 8x8 samples -> DCT -> quantise -> zig-zag, entirely in registers.
#pragma once
#include "b2j_common.cuh"
#include "b2j_math.cuh"

namespace b2j {

__device__ __forceinline__ constexpr int zz_of(int i)
{
    constexpr uint8_t t[64] = B2J_ZIGZAG_INIT;
    return t[i];
}

// Collects the 64 quantised coefficients of one block into 32 packed words in
// zig-zag order (all indices are compile-time constants -> registers).
struct BlockPacker {
    uint32_t w[32];
    const uint32_t *mul;          // 64 multipliers of this block's table (shared memory)
    __device__ __forceinline__ void init(const uint32_t *m)
    {
        mul = m;
#pragma unroll
        for (int i = 0; i < 32; i++) w[i] = 0u;
    }
    template <int IDX>
    __device__ __forceinline__ void put(uint32_t mag, bool neg)
    {
        constexpr int z = zz_of(IDX);
        uint32_t q = quant_mag(mag, mul[IDX]);
        int r = neg ? -(int)q : (int)q;
        if (z & 1) w[z >> 1] = (uint32_t)r * 65536u + w[z >> 1];
        else w[z >> 1] = ((uint32_t)r & 0xFFFFu) | w[z >> 1];
    }
    __device__ __forceinline__ void store(int16_t *dst) const
    {
        uint4 *o = reinterpret_cast<uint4 *>(dst);
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
};

template <int U, class P>
__device__ __forceinline__ void emit_col_d(const double (&c)[8], P &p)
{
    p.template put<0 * 8 + U>(abs_div8_trunc(c[0]), __double2hiint(c[0]) < 0);
    p.template put<1 * 8 + U>(abs_div8_trunc(c[1]), __double2hiint(c[1]) < 0);
    p.template put<2 * 8 + U>(abs_div8_trunc(c[2]), __double2hiint(c[2]) < 0);
    p.template put<3 * 8 + U>(abs_div8_trunc(c[3]), __double2hiint(c[3]) < 0);
    p.template put<4 * 8 + U>(abs_div8_trunc(c[4]), __double2hiint(c[4]) < 0);
    p.template put<5 * 8 + U>(abs_div8_trunc(c[5]), __double2hiint(c[5]) < 0);
    p.template put<6 * 8 + U>(abs_div8_trunc(c[6]), __double2hiint(c[6]) < 0);
    p.template put<7 * 8 + U>(abs_div8_trunc(c[7]), __double2hiint(c[7]) < 0);
}

template <int U, class P>
__device__ __forceinline__ void emit_col_i(const int (&m)[8], P &p)
{
    int r0, r4;
    double c[8];
    dct_1d_int(m, r0, r4, c[1], c[2], c[3], c[5], c[6], c[7]);
    p.template put<0 * 8 + U>((uint32_t)abs(r0) >> 3, r0 < 0);
    p.template put<4 * 8 + U>((uint32_t)abs(r4) >> 3, r4 < 0);
    p.template put<1 * 8 + U>(abs_div8_trunc(c[1]), __double2hiint(c[1]) < 0);
    p.template put<2 * 8 + U>(abs_div8_trunc(c[2]), __double2hiint(c[2]) < 0);
    p.template put<3 * 8 + U>(abs_div8_trunc(c[3]), __double2hiint(c[3]) < 0);
    p.template put<5 * 8 + U>(abs_div8_trunc(c[5]), __double2hiint(c[5]) < 0);
    p.template put<6 * 8 + U>(abs_div8_trunc(c[6]), __double2hiint(c[6]) < 0);
    p.template put<7 * 8 + U>(abs_div8_trunc(c[7]), __double2hiint(c[7]) < 0);
}

// 8x8 samples (8 rows of 8 bytes) -> quantised zig-zag block in `p`.
// Row pass: dct_1d on every row (enc_frame.py:92-100); columns 0 and 4 of the
// intermediate matrix are integers.  Column pass: enc_frame.py:103-108.
template <class P>
__device__ __forceinline__ void dct_quant_block(const uint2 (&px)[8], P &p)
{
    int m0[8], m4[8];
    double m1[8], m2[8], m3[8], m5[8], m6[8], m7[8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        int s[8];
#ifndef B2J_ROW_SHIFT_MASK      // PRMT extraction: 1 op per byte instead of shift+mask (1.501 -> 1.476 ms)
        s[0] = (int)__byte_perm(px[x].x, 0u, 0x4440) - 128; s[1] = (int)__byte_perm(px[x].x, 0u, 0x4441) - 128;
        s[2] = (int)__byte_perm(px[x].x, 0u, 0x4442) - 128; s[3] = (int)__byte_perm(px[x].x, 0u, 0x4443) - 128;
        s[4] = (int)__byte_perm(px[x].y, 0u, 0x4440) - 128; s[5] = (int)__byte_perm(px[x].y, 0u, 0x4441) - 128;
        s[6] = (int)__byte_perm(px[x].y, 0u, 0x4442) - 128; s[7] = (int)__byte_perm(px[x].y, 0u, 0x4443) - 128;
#else
        s[0] = (int)(px[x].x & 0xFFu) - 128; s[1] = (int)((px[x].x >> 8) & 0xFFu) - 128;
        s[2] = (int)((px[x].x >> 16) & 0xFFu) - 128; s[3] = (int)(px[x].x >> 24) - 128;
        s[4] = (int)(px[x].y & 0xFFu) - 128; s[5] = (int)((px[x].y >> 8) & 0xFFu) - 128;
        s[6] = (int)((px[x].y >> 16) & 0xFFu) - 128; s[7] = (int)(px[x].y >> 24) - 128;
#endif
        dct_1d_int(s, m0[x], m4[x], m1[x], m2[x], m3[x], m5[x], m6[x], m7[x]);
    }
    emit_col_i<0>(m0, p);
    emit_col_i<4>(m4, p);
    dct_1d(m1); emit_col_d<1>(m1, p);
    dct_1d(m2); emit_col_d<2>(m2, p);
    dct_1d(m3); emit_col_d<3>(m3, p);
    dct_1d(m5); emit_col_d<5>(m5, p);
    dct_1d(m6); emit_col_d<6>(m6, p);
    dct_1d(m7); emit_col_d<7>(m7, p);
}

// [ext] the same block in the gateware's fixed-point arithmetic (b2j_math.cuh dct_1d_fixed):
// rows, columns, arithmetic >> 3 (new_dct6.py:107-140), then the model's quantiser.
template <int U, class P>
__device__ __forceinline__ void emit_col_fixed(const int (&c)[8], P &p)
{
#define B2J_EMITF(VV) do { const int o_ = c[VV] >> 3; p.template put<VV * 8 + U>((uint32_t)abs(o_), o_ < 0); } while (0)
    B2J_EMITF(0); B2J_EMITF(1); B2J_EMITF(2); B2J_EMITF(3); B2J_EMITF(4); B2J_EMITF(5); B2J_EMITF(6); B2J_EMITF(7);
#undef B2J_EMITF
}

template <class P>
__device__ __forceinline__ void dct_quant_block_fixed(const uint2 (&px)[8], P &p)
{
    int m[8][8];                   // m[u][x]: row-pass output u of row x
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
    dct_1d_fixed(m[0]); emit_col_fixed<0>(m[0], p);
    dct_1d_fixed(m[1]); emit_col_fixed<1>(m[1], p);
    dct_1d_fixed(m[2]); emit_col_fixed<2>(m[2], p);
    dct_1d_fixed(m[3]); emit_col_fixed<3>(m[3], p);
    dct_1d_fixed(m[4]); emit_col_fixed<4>(m[4], p);
    dct_1d_fixed(m[5]); emit_col_fixed<5>(m[5], p);
    dct_1d_fixed(m[6]); emit_col_fixed<6>(m[6], p);
    dct_1d_fixed(m[7]); emit_col_fixed<7>(m[7], p);
}

}  // namespace b2j
