/* CPU oracle, plain C.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restates, in the same float64 operation order, the reference software model
 * /root/reference/test/model/enc_frame.py (rgb2ycbcr :27-35, dct :42-109,
 * quantize :111-129, zigzag :131-139/:161-165, DC differencing :219-230) and the
 * [ext] stages defined in oracle/jpeg_ext.py (subsampling, IJG quality tables,
 * gateware RLE symbol rules litejpeg/core/rle/rlecore.py:84-132, Annex-K Huffman
 * coding, bit packing, 0xFF stuffing, JFIF container, baseline decoder).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, which
 * would change float64 rounding).  Pinned by tests/test_oracle_c.py against
 * oracle/model_np.py and the golden fixtures generated from the real model.
 * Used by the GPU parity tests at full frame sizes and as the timed CPU
 * baseline (kind "port") in bench.py.  Never linked into the product library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* math.cos((n*pi)/16), math.sin((n*pi)/16) with pi = 3.14159265358979323846 as
 * CPython evaluates them (enc_frame.py:38-44); hex literals keep them exact. */
static const double C1 = 0x1.f6297cff75cb0p-1, S1 = 0x1.8f8b83c69a60ap-3;
static const double C3 = 0x1.a9b66290ea1a3p-1, S3 = 0x1.1c73b39ae68c8p-1;
static const double C6 = 0x1.87de2a6aea964p-2, S6 = 0x1.d906bcf328d46p-1;
static const double SQRT2 = 0x1.6a09e667f3bcdp+0;   /* enc_frame.py:39 */

static const uint8_t ZIGZAG[64] = {   /* enc_frame.py:131-139 */
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

static const uint8_t BASE_LUMA[64] = {   /* Annex K.1 == quantization.py:20-27 */
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t BASE_CHROMA[64] = { /* Annex K.2 == test/common.py:329-336 */
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* ------------------------------------------------------------ stages */

/* enc_frame.py:27-35: left-to-right float64, int() truncation. */
static inline void px_rgb2ycbcr(int r, int g, int b, int *y, int *cb, int *cr)
{
    double t;
    t = 0.299 * r; t = t + 0.587 * g; t = t + 0.114 * b;              *y = (int)t;
    t = -0.1687 * r; t = t - 0.3313 * g; t = t + 0.5 * b; t = t + 128; *cb = (int)t;
    t = 0.5 * r; t = t - 0.4187 * g; t = t - 0.0813 * b; t = t + 128;  *cr = (int)t;
}

/* enc_frame.py:54-84 */
static void dct_1d(double *v /* 8, in place -> r */)
{
    double a, b, t0, t1;
#define BF(i, j) do { a = v[i] + v[j]; b = v[i] - v[j]; v[i] = a; v[j] = b; } while (0)
#define ROT(i, j, k, c, s) do { t0 = (k) * (v[i] * (c) + v[j] * (s)); \
                                t1 = (k) * (v[j] * (c) - v[i] * (s)); v[i] = t0; v[j] = t1; } while (0)
    BF(3, 4); BF(2, 5); BF(1, 6); BF(0, 7);                 /* stage 1 */
    BF(0, 3); BF(1, 2); ROT(4, 7, 1.0, C3, S3); ROT(5, 6, 1.0, C1, S1);  /* stage 2 */
    BF(7, 5); BF(4, 6); BF(0, 1); ROT(2, 3, SQRT2, C6, S6);  /* stage 3 */
    {   /* stage 4 */
        double r[8];
        r[1] = v[7] + v[4]; r[7] = v[7] - v[4];
        r[5] = SQRT2 * v[6]; r[3] = SQRT2 * v[5];
        r[6] = v[3]; r[2] = v[2]; r[4] = v[1]; r[0] = v[0];
        memcpy(v, r, sizeof r);
    }
#undef BF
#undef ROT
}

/* enc_frame.py:86-109: samples 0..255 row-major -> int(v/8), row-major [v][u]. */
static void dct_block(const int *in, int *out)
{
    double m[64], vec[8];
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) vec[y] = (double)(in[x * 8 + y] - 128);
        dct_1d(vec);
        for (int y = 0; y < 8; y++) m[x * 8 + y] = vec[y];
    }
    for (int y = 0; y < 8; y++) {
        for (int x = 0; x < 8; x++) vec[x] = m[x * 8 + y];
        dct_1d(vec);
        for (int x = 0; x < 8; x++) out[x * 8 + y] = (int)(vec[x] / 8);
    }
}

/* [ext] "fixed" DCT mode: litejpeg/core/new_dct6.py:171-396 with its own constants and shift points,
 * the two wiring slips (:262-263, :294) corrected and wide sums; see oracle/gateware.py dct_1d_fixed.
 * `>>` on negative ints is an arithmetic shift with gcc (floor), as in the gateware. */
static void dct_1d_fixed(int *v)
{
    int a, b;
#define BF(i, j) do { a = v[i] + v[j]; b = v[i] - v[j]; v[i] = a; v[j] = b; } while (0)
    BF(3, 4); BF(2, 5); BF(1, 6); BF(0, 7);
    BF(0, 3); BF(1, 2);
    {
        const int v4 = v[4], v7 = v[7], v5 = v[5], v6 = v[6];
        v[4] = ((213 * v4) >> 8) + ((569 * v7) >> 10);
        v[7] = ((213 * v7) >> 8) - ((569 * v4) >> 10);
        v[5] = ((1004 * v5) >> 10) + ((25 * v6) >> 7);
        v[6] = ((1004 * v6) >> 10) - ((25 * v5) >> 7);
    }
    BF(7, 5); BF(4, 6); BF(0, 1);
    {
        const int fa = ((49 * v[2]) >> 7) + ((473 * v[3]) >> 9);
        const int fs = ((49 * v[3]) >> 7) - ((473 * v[2]) >> 9);
        v[2] = (181 * fa) >> 7; v[3] = (181 * fs) >> 7;
    }
    {
        int r[8];
        r[1] = v[7] + v[4]; r[7] = v[7] - v[4];
        r[5] = (181 * v[6]) >> 7; r[3] = (181 * v[5]) >> 7;
        r[6] = v[3]; r[2] = v[2]; r[4] = v[1]; r[0] = v[0];
        memcpy(v, r, sizeof r);
    }
#undef BF
}

/* new_dct6.py:107-145: -128, rows, columns, arithmetic >> 3 */
static void dct_block_fixed(const int *in, int *out)
{
    int m[64], vec[8];
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) vec[y] = in[x * 8 + y] - 128;
        dct_1d_fixed(vec);
        for (int y = 0; y < 8; y++) m[x * 8 + y] = vec[y];
    }
    for (int y = 0; y < 8; y++) {
        for (int x = 0; x < 8; x++) vec[x] = m[x * 8 + y];
        dct_1d_fixed(vec);
        for (int x = 0; x < 8; x++) out[x * 8 + y] = vec[x] >> 3;
    }
}

/* enc_frame.py:111-129 == sign(v) * ((2|v| + d) / (2d)) */
static inline int quant1(int v, int d)
{
    int a = v < 0 ? -v : v;
    a = (2 * a + d) / (2 * d);
    return v < 0 ? -a : a;
}

int b2jo_rgb2ycbcr(const uint8_t *rgb, int64_t npx, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    for (int64_t i = 0; i < npx; i++) {
        int a, b, c;
        px_rgb2ycbcr(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], &a, &b, &c);
        y[i] = (uint8_t)a; cb[i] = (uint8_t)b; cr[i] = (uint8_t)c;
    }
    return 0;
}

int b2jo_dct(const int32_t *in, int64_t nblk, int32_t *out)
{
    for (int64_t i = 0; i < nblk; i++) dct_block(in + 64 * i, out + 64 * i);
    return 0;
}

int b2jo_dct_fixed(const int32_t *in, int64_t nblk, int32_t *out)
{
    for (int64_t i = 0; i < nblk; i++) dct_block_fixed(in + 64 * i, out + 64 * i);
    return 0;
}

int b2jo_quantize(const int32_t *in, int64_t nblk, const int32_t *table, int32_t *out)
{
    for (int64_t i = 0; i < nblk * 64; i++) out[i] = quant1(in[i], table[i & 63]);
    return 0;
}

int b2jo_zigzag(const int32_t *in, int64_t nblk, int32_t *out)
{
    for (int64_t i = 0; i < nblk; i++)
        for (int k = 0; k < 64; k++) out[64 * i + ZIGZAG[k]] = in[64 * i + k];
    return 0;
}

void b2jo_quant_tables(int quality, uint8_t *qy, uint8_t *qc)
{
    int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int i = 0; i < 64; i++) {
        int a = (BASE_LUMA[i] * s + 50) / 100, b = (BASE_CHROMA[i] * s + 50) / 100;
        qy[i] = (uint8_t)(a < 1 ? 1 : a > 255 ? 255 : a);
        qc[i] = (uint8_t)(b < 1 ? 1 : b > 255 ? 255 : b);
    }
}

/* ------------------------------------------------------------ geometry */

typedef struct {
    int height, width, h, v, mcus_x, mcus_y, bpm, n_mcu, padded_w, padded_h;
} geo_t;

static int make_geo(int height, int width, int sub, geo_t *g)
{
    if (sub == 444) { g->h = 1; g->v = 1; }
    else if (sub == 422) { g->h = 2; g->v = 1; }
    else if (sub == 420) { g->h = 2; g->v = 2; }
    else return -1;
    g->height = height; g->width = width;
    g->mcus_x = (width + 8 * g->h - 1) / (8 * g->h);
    g->mcus_y = (height + 8 * g->v - 1) / (8 * g->v);
    g->bpm = g->h * g->v + 2;
    g->n_mcu = g->mcus_x * g->mcus_y;
    g->padded_w = g->mcus_x * 8 * g->h;
    g->padded_h = g->mcus_y * 8 * g->v;
    return 0;
}

/* RGB (H, W, 3) -> zig-zag quantised coefficients, scan (MCU) order,
 * [n_mcu][bpm][64] int16, DC not differenced. */
int b2jo_encode_coefficients_ex(const uint8_t *rgb, int height, int width, int sub,
                                const uint8_t *qy, const uint8_t *qc, int dct_fixed, int16_t *coefs);
int b2jo_encode_coefficients(const uint8_t *rgb, int height, int width, int sub,
                             const uint8_t *qy, const uint8_t *qc, int16_t *coefs)
{
    return b2jo_encode_coefficients_ex(rgb, height, width, sub, qy, qc, 0, coefs);
}

/* dct_fixed != 0: the [ext] fixed-point DCT mode (dct_block_fixed) instead of the model's float64 DCT */
int b2jo_encode_coefficients_ex(const uint8_t *rgb, int height, int width, int sub,
                                const uint8_t *qy, const uint8_t *qc, int dct_fixed, int16_t *coefs)
{
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    const int mw = 8 * g.h, mh = 8 * g.v;
#pragma omp parallel for schedule(static)
    for (int my = 0; my < g.mcus_y; my++) {
        int Y[16 * 16], CB[16 * 16], CR[16 * 16], blk[64], d[64];
        for (int mx = 0; mx < g.mcus_x; mx++) {
            for (int py = 0; py < mh; py++) {
                int sy = my * mh + py; if (sy >= height) sy = height - 1;   /* edge replicate */
                for (int px = 0; px < mw; px++) {
                    int sx = mx * mw + px; if (sx >= width) sx = width - 1;
                    const uint8_t *p = rgb + ((int64_t)sy * width + sx) * 3;
                    px_rgb2ycbcr(p[0], p[1], p[2], &Y[py * 16 + px], &CB[py * 16 + px], &CR[py * 16 + px]);
                }
            }
            int16_t *out = coefs + ((int64_t)my * g.mcus_x + mx) * g.bpm * 64;
            for (int b = 0; b < g.bpm; b++) {
                const uint8_t *q;
                if (b < g.h * g.v) {
                    int bx = b % g.h, by = b / g.h;
                    for (int i = 0; i < 64; i++) blk[i] = Y[(by * 8 + i / 8) * 16 + bx * 8 + i % 8];
                    q = qy;
                } else {
                    const int *C = (b == g.h * g.v) ? CB : CR;
                    for (int i = 0; i < 64; i++) {
                        int yy = (i / 8) * g.v, xx = (i % 8) * g.h, s;
                        /* crs.py:39-54 pair mean, horizontally then vertically */
                        if (g.h == 2) {
                            s = (C[yy * 16 + xx] + C[yy * 16 + xx + 1]) >> 1;
                            if (g.v == 2)
                                s = (s + ((C[(yy + 1) * 16 + xx] + C[(yy + 1) * 16 + xx + 1]) >> 1)) >> 1;
                        } else s = C[yy * 16 + xx];
                        blk[i] = s;
                    }
                    q = qc;
                }
                if (dct_fixed) dct_block_fixed(blk, d); else dct_block(blk, d);
                for (int i = 0; i < 64; i++) out[b * 64 + ZIGZAG[i]] = (int16_t)quant1(d[i], q[i]);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------ Huffman */

static const uint8_t DC_L_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t DC_C_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t AC_L_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t AC_L_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};
static const uint8_t AC_C_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t AC_C_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};

typedef struct {
    uint16_t code[256]; uint8_t len[256];
    int mincode[17], maxcode[17], valptr[17]; const uint8_t *vals;   /* T.81 F.2.2.3 decode aids */
} hufftab_t;

static void derive(const uint8_t *bits, const uint8_t *vals, hufftab_t *t)
{
    memset(t, 0, sizeof *t);
    t->vals = vals;
    int c = 0, k = 0;
    for (int ln = 1; ln <= 16; ln++) {
        t->valptr[ln] = k; t->mincode[ln] = c; t->maxcode[ln] = bits[ln - 1] ? c + bits[ln - 1] - 1 : -1;
        for (int i = 0; i < bits[ln - 1]; i++, k++, c++) { t->code[vals[k]] = (uint16_t)c; t->len[vals[k]] = (uint8_t)ln; }
        c <<= 1;
    }
}

static hufftab_t T_DC_L, T_DC_C, T_AC_L, T_AC_C;
static int tables_ready = 0;
static void init_tables(void)
{
    if (tables_ready) return;
#pragma omp critical
    {
        if (!tables_ready) {
            derive(DC_L_BITS, DC_VALS, &T_DC_L); derive(DC_C_BITS, DC_VALS, &T_DC_C);
            derive(AC_L_BITS, AC_L_VALS, &T_AC_L); derive(AC_C_BITS, AC_C_VALS, &T_AC_C);
            tables_ready = 1;
        }
    }
}

typedef struct { uint8_t *p, *end; uint64_t acc; int n; int64_t bits; int overflow; } bitw_t;

static inline void put_byte(bitw_t *w, int b)
{
    if (w->p < w->end) *w->p = (uint8_t)b; else w->overflow = 1;
    w->p++;
    if (b == 0xFF) { if (w->p < w->end) *w->p = 0; else w->overflow = 1; w->p++; }   /* stuffing */
}

static inline void put_bits(bitw_t *w, uint32_t val, int n)
{
    if (!n) return;
    w->acc = (w->acc << n) | (val & ((1u << n) - 1));
    w->n += n; w->bits += n;
    while (w->n >= 8) { put_byte(w, (int)((w->acc >> (w->n - 8)) & 0xFF)); w->n -= 8; }
}

static inline int bitlen(int a) { int n = 0; while (a) { n++; a >>= 1; } return n; }

/* One block: symbols per rlecore.py:84-132 (zrl=0 eager) or libjpeg (zrl=1),
 * sizes per entropycoder.py:55-65, codes Annex K.  w may be NULL (count only). */
static int encode_block(bitw_t *w, const int16_t *c, int diff, const hufftab_t *dc,
                        const hufftab_t *ac, int zrl)
{
    int total = 0;
#define EMIT(v, n) do { total += (n); if (w) put_bits(w, (uint32_t)(v), (n)); } while (0)
    int s = bitlen(diff < 0 ? -diff : diff);
    EMIT(dc->code[s], dc->len[s]);
    EMIT(diff < 0 ? diff - 1 : diff, s);
    int run = 0;
    for (int k = 1; k < 64; k++) {
        int v = c[k];
        if (v == 0) {
            if (zrl == 0) {
                if (run == 15) { EMIT(ac->code[0xF0], ac->len[0xF0]); run = 0; }
                else if (k == 63) EMIT(ac->code[0], ac->len[0]);
                else run++;
            } else {
                if (k == 63) EMIT(ac->code[0], ac->len[0]);
                else run++;
            }
        } else {
            while (run > 15) { EMIT(ac->code[0xF0], ac->len[0xF0]); run -= 16; }
            s = bitlen(v < 0 ? -v : v);
            int rs = (run << 4) | s;
            EMIT(ac->code[rs], ac->len[rs]);
            EMIT(v < 0 ? v - 1 : v, s);
            run = 0;
        }
    }
#undef EMIT
    return total;
}

/* bits per block, [n_mcu][bpm] */
int b2jo_block_bits(const int16_t *coefs, int height, int width, int sub, int zrl, int32_t *bits)
{
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    init_tables();
    int pred[3] = {0, 0, 0};
    for (int64_t m = 0; m < g.n_mcu; m++)
        for (int b = 0; b < g.bpm; b++) {
            int comp = b < g.h * g.v ? 0 : b - g.h * g.v + 1;
            const int16_t *c = coefs + (m * g.bpm + b) * 64;
            int diff = c[0] - pred[comp]; pred[comp] = c[0];
            bits[m * g.bpm + b] = encode_block(NULL, c, diff, comp ? &T_DC_C : &T_DC_L,
                                               comp ? &T_AC_C : &T_AC_L, zrl);
        }
    return 0;
}

/* Scan data: stuffed, final byte padded with 1s.  Returns bytes needed
 * (> cap means the output was truncated). */
int64_t b2jo_entropy_encode(const int16_t *coefs, int height, int width, int sub, int zrl,
                            uint8_t *out, int64_t cap, int64_t *total_bits)
{
    /* [ext] zrl >> 16 = restart interval in MCUs (T.81 B.2.4.4, E.1.4): every interval the byte is
     * padded with 1s, RSTm (0xFFD0 + m mod 8) is written unstuffed, the DC predictors restart. */
    const int ri = (zrl >> 16) & 0xFFFF;
    zrl &= 0xFF;
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    init_tables();
    bitw_t w = {out, out + cap, 0, 0, 0, 0};
    int pred[3] = {0, 0, 0};
    for (int64_t m = 0; m < g.n_mcu; m++) {
        if (ri && m && m % ri == 0) {
            if (w.n) { int pad = 8 - w.n; put_bits(&w, (1u << pad) - 1, pad); w.bits -= pad; }
            if (w.p < w.end) *w.p = 0xFF; else w.overflow = 1;
            w.p++;
            if (w.p < w.end) *w.p = (uint8_t)(0xD0 + ((m / ri - 1) & 7)); else w.overflow = 1;
            w.p++;
            pred[0] = pred[1] = pred[2] = 0;
        }
        for (int b = 0; b < g.bpm; b++) {
            int comp = b < g.h * g.v ? 0 : b - g.h * g.v + 1;
            const int16_t *c = coefs + (m * g.bpm + b) * 64;
            int diff = c[0] - pred[comp]; pred[comp] = c[0];
            encode_block(&w, c, diff, comp ? &T_DC_C : &T_DC_L, comp ? &T_AC_C : &T_AC_L, zrl);
        }
    }
    if (total_bits) *total_bits = w.bits;
    if (w.n) { int pad = 8 - w.n; put_bits(&w, (1u << pad) - 1, pad); }
    return (int64_t)(w.p - out);
}

/* ------------------------------------------------------------ container */

static uint8_t *seg(uint8_t *p, int marker, const uint8_t *payload, int n)
{
    *p++ = 0xFF; *p++ = (uint8_t)marker; *p++ = (uint8_t)((n + 2) >> 8); *p++ = (uint8_t)((n + 2) & 0xFF);
    memcpy(p, payload, (size_t)n);
    return p + n;
}

/* SOI APP0 DQT DQT SOF0 DHT x4 SOS; returns length (623 bytes). out >= 1024 B. */
int b2jo_jfif_header_ex(int height, int width, int sub, const uint8_t *qy, const uint8_t *qc, int restart_interval,
                        uint8_t *out);
int b2jo_jfif_header(int height, int width, int sub, const uint8_t *qy, const uint8_t *qc, uint8_t *out)
{
    return b2jo_jfif_header_ex(height, width, sub, qy, qc, 0, out);
}

/* restart_interval > 0 adds a DRI segment in front of SOS (629 bytes) */
int b2jo_jfif_header_ex(int height, int width, int sub, const uint8_t *qy, const uint8_t *qc, int restart_interval,
                        uint8_t *out)
{
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    uint8_t *p = out, buf[256];
    *p++ = 0xFF; *p++ = 0xD8;
    { const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0}; p = seg(p, 0xE0, app0, 14); }
    for (int t = 0; t < 2; t++) {
        const uint8_t *q = t ? qc : qy;
        buf[0] = (uint8_t)t;
        for (int i = 0; i < 64; i++) buf[1 + ZIGZAG[i]] = q[i];
        p = seg(p, 0xDB, buf, 65);
    }
    {
        uint8_t sof[15] = {8, (uint8_t)(height >> 8), (uint8_t)height, (uint8_t)(width >> 8), (uint8_t)width, 3,
                           1, (uint8_t)((g.h << 4) | g.v), 0, 2, 0x11, 1, 3, 0x11, 1};
        p = seg(p, 0xC0, sof, 15);
    }
    {
        const uint8_t *bits[4] = {DC_L_BITS, AC_L_BITS, DC_C_BITS, AC_C_BITS};
        const uint8_t *vals[4] = {DC_VALS, AC_L_VALS, DC_VALS, AC_C_VALS};
        const int nv[4] = {12, 162, 12, 162};
        const uint8_t id[4] = {0x00, 0x10, 0x01, 0x11};
        for (int t = 0; t < 4; t++) {
            buf[0] = id[t]; memcpy(buf + 1, bits[t], 16); memcpy(buf + 17, vals[t], (size_t)nv[t]);
            p = seg(p, 0xC4, buf, 17 + nv[t]);
        }
    }
    if (restart_interval > 0) {
        const uint8_t dri[2] = {(uint8_t)(restart_interval >> 8), (uint8_t)restart_interval};
        p = seg(p, 0xDD, dri, 2);
    }
    { const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0}; p = seg(p, 0xDA, sos, 10); }
    return (int)(p - out);
}

/* RGB -> JFIF file.  Returns bytes needed (> cap: truncated), < 0 on error.
 * zrl: 0 eager / 1 standard, optionally | 0x100 for the [ext] fixed-point DCT mode,
 * optionally | (restart interval in MCUs << 16). */
int64_t b2jo_encode_jpeg(const uint8_t *rgb, int height, int width, int sub, int quality, int zrl,
                         uint8_t *out, int64_t cap)
{
    const int dct_fixed = (zrl & 0x100) != 0;
    zrl &= ~0x100;
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    if (cap < 1024) return -2;
    uint8_t qy[64], qc[64];
    b2jo_quant_tables(quality, qy, qc);
    int16_t *coefs = (int16_t *)malloc((size_t)g.n_mcu * g.bpm * 64 * sizeof(int16_t));
    if (!coefs) return -3;
    b2jo_encode_coefficients_ex(rgb, height, width, sub, qy, qc, dct_fixed, coefs);
    int hl = b2jo_jfif_header_ex(height, width, sub, qy, qc, (zrl >> 16) & 0xFFFF, out);
    int64_t n = b2jo_entropy_encode(coefs, height, width, sub, zrl, out + hl, cap - hl - 2 > 0 ? cap - hl - 2 : 0, NULL);
    free(coefs);
    int64_t total = hl + n + 2;
    if (total <= cap) { out[hl + n] = 0xFF; out[hl + n + 1] = 0xD9; }
    return total;
}

/* Batch of equally sized frames, one slot of `cap` bytes per frame.  Frames are
 * spread over OpenMP threads (the model itself is single-threaded Python; this
 * is the strongest honest CPU arrangement of the same arithmetic). */
int b2jo_encode_batch(const uint8_t *rgb, int n, int height, int width, int sub, int quality, int zrl,
                      uint8_t *out, int64_t cap, int64_t *sizes, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    omp_set_max_active_levels(1);
#endif
    init_tables();
    int64_t fsz = (int64_t)height * width * 3;
    if (n == 1) { sizes[0] = b2jo_encode_jpeg(rgb, height, width, sub, quality, zrl, out, cap); return 0; }
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; i++)
        sizes[i] = b2jo_encode_jpeg(rgb + i * fsz, height, width, sub, quality, zrl, out + i * cap, cap);
    return 0;
}

int b2jo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------ decoder */

typedef struct { const uint8_t *p, *end; uint32_t acc; int n; int err; } bitr_t;

static inline int get_bit(bitr_t *r)
{
    if (r->n == 0) {
        if (r->p >= r->end) { r->err = 1; return 0; }
        int b = *r->p++;
        if (b == 0xFF) { if (r->p >= r->end || *r->p != 0) { r->err = 2; return 0; } r->p++; }
        r->acc = (uint32_t)b; r->n = 8;
    }
    r->n--;
    return (int)((r->acc >> r->n) & 1);
}

static int get_sym(bitr_t *r, const hufftab_t *t)
{
    int code = 0;
    for (int ln = 1; ln <= 16; ln++) {
        code = (code << 1) | get_bit(r);
        if (code <= t->maxcode[ln]) return t->vals[t->valptr[ln] + code - t->mincode[ln]];
    }
    r->err = 3;
    return 0;
}

static int get_ext(bitr_t *r, int s)
{
    int v = 0;
    for (int i = 0; i < s; i++) v = (v << 1) | get_bit(r);
    return (s == 0 || v >= (1 << (s - 1))) ? v : v - (1 << s) + 1;
}

/* Decodes a file written by this oracle / the CUDA path (Annex-K tables are
 * assumed, header fields are re-derived from the arguments).  coefs gets DC
 * prediction undone.  Returns 0, or a positive error code. */
int b2jo_decode_coefficients(const uint8_t *jpg, int64_t n, int height, int width, int sub, int16_t *coefs)
{
    geo_t g;
    if (make_geo(height, width, sub, &g)) return -1;
    init_tables();
    /* find SOS */
    int64_t pos = 2;
    int ri = 0;
    while (pos + 4 <= n) {
        if (jpg[pos] != 0xFF) return 10;
        int m = jpg[pos + 1], ln = (jpg[pos + 2] << 8) | jpg[pos + 3];
        if (m == 0xDD) ri = (jpg[pos + 4] << 8) | jpg[pos + 5];
        pos += 2 + ln;
        if (m == 0xDA) break;
    }
    bitr_t r = {jpg + pos, jpg + n, 0, 0, 0};
    int pred[3] = {0, 0, 0};
    memset(coefs, 0, (size_t)g.n_mcu * g.bpm * 64 * sizeof(int16_t));
    for (int64_t m = 0; m < g.n_mcu; m++) {
        if (ri && m && m % ri == 0) {                       /* 1-padding, RSTm, predictors restart */
            while (r.n) if (get_bit(&r) != 1) return 14;
            if (r.p + 2 > r.end || r.p[0] != 0xFF || r.p[1] != 0xD0 + ((m / ri - 1) & 7)) return 15;
            r.p += 2;
            pred[0] = pred[1] = pred[2] = 0;
        }
        for (int b = 0; b < g.bpm; b++) {
            int comp = b < g.h * g.v ? 0 : b - g.h * g.v + 1;
            int16_t *c = coefs + (m * g.bpm + b) * 64;
            int s = get_sym(&r, comp ? &T_DC_C : &T_DC_L);
            pred[comp] += get_ext(&r, s);
            c[0] = (int16_t)pred[comp];
            for (int k = 1; k < 64;) {
                int rs = get_sym(&r, comp ? &T_AC_C : &T_AC_L);
                int run = rs >> 4; s = rs & 15;
                if (s == 0) { if (run == 15) { k += 16; continue; } break; }
                k += run;
                if (k > 63) return 11;
                c[k++] = (int16_t)get_ext(&r, s);
            }
            if (r.err) return r.err;
        }
    }
    while (r.n) if (get_bit(&r) != 1) return 12;
    if (r.p + 2 > r.end || r.p[0] != 0xFF || r.p[1] != 0xD9) return 13;
    return 0;
}
