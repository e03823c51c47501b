// Host-side helpers: error text, geometry, constant tables, JFIF header.
#include <stdarg.h>

#include "b2j_common.cuh"

namespace b2j {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

const uint8_t kZigzag[64] = B2J_ZIGZAG_INIT;

// ITU-T T.81 Annex K.1 / K.2 (the same numbers as litejpeg/core/quantization.py:20-27
// and test/common.py:319-336 of the reference).
const uint8_t kBaseLuma[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kBaseChroma[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// Annex K.3.3 typical Huffman tables.
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
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
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
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

// T.81 Annex C canonical code assignment.
static void derive(const uint8_t *bits, const uint8_t *vals, uint32_t *lut)
{
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < bits[len - 1]; i++, k++, code++) lut[vals[k]] = ((uint32_t)len << 16) | code;
        code <<= 1;
    }
}

const HuffLut &host_huff_lut()
{
    static const HuffLut lut = [] {
        HuffLut l;
        memset(&l, 0, sizeof l);
        uint32_t tmp[256];
        memset(tmp, 0, sizeof tmp); derive(kDcLumaBits, kDcVals, tmp); memcpy(l.dc[0], tmp, sizeof l.dc[0]);
        memset(tmp, 0, sizeof tmp); derive(kDcChromaBits, kDcVals, tmp); memcpy(l.dc[1], tmp, sizeof l.dc[1]);
        derive(kAcLumaBits, kAcLumaVals, l.ac[0]);
        derive(kAcChromaBits, kAcChromaVals, l.ac[1]);
        return l;
    }();
    return lut;
}

int make_geo(const b2j_image_desc *d, Geo *g)
{
    B2J_REQUIRE(d && g, B2J_E_ARG, "null image descriptor");
    B2J_REQUIRE(d->width >= 1 && d->width <= 65535 && d->height >= 1 && d->height <= 65535, B2J_E_SHAPE,
                "image side out of range 1..65535: %d x %d", d->width, d->height);
    B2J_REQUIRE(d->n_frames >= 1, B2J_E_ARG, "n_frames must be >= 1");
    g->layout = (d->subsampling >> 16) & 0x7FFF;
    B2J_REQUIRE(g->layout == B2J_LAYOUT_RGB || g->layout == B2J_LAYOUT_BGR || g->layout == B2J_LAYOUT_RGBA || g->layout == B2J_LAYOUT_BGRA,
                B2J_E_ARG, "descriptor layout must be RGB, BGR, RGBA or BGRA (got %d)", g->layout);
    g->bpp = (g->layout == B2J_LAYOUT_RGBA || g->layout == B2J_LAYOUT_BGRA) ? 4 : 3;
    switch (d->subsampling & 0xFFFF) {
    case B2J_SUB_444: g->h = 1; g->v = 1; break;
    case B2J_SUB_422: g->h = 2; g->v = 1; break;
    case B2J_SUB_420: g->h = 2; g->v = 2; break;
    default: B2J_REQUIRE(false, B2J_E_ARG, "subsampling must be 444, 422 or 420 (got %d)", d->subsampling & 0xFFFF);
    }
    g->width = d->width; g->height = d->height; g->n_frames = d->n_frames;
    g->mcu_w = 8 * g->h; g->mcu_h = 8 * g->v;
    g->mcus_x = (d->width + g->mcu_w - 1) / g->mcu_w;
    g->mcus_y = (d->height + g->mcu_h - 1) / g->mcu_h;
    g->bpm = g->h * g->v + 2;
    const long long row_bytes = (long long)g->bpp * d->width;
    g->row_stride = d->row_stride ? d->row_stride : row_bytes;
    g->frame_stride = d->frame_stride ? d->frame_stride : g->row_stride * d->height;
    B2J_REQUIRE(g->row_stride >= row_bytes, B2J_E_ARG, "row_stride smaller than %d*width", g->bpp);
    B2J_REQUIRE(g->frame_stride >= g->row_stride * (long long)(d->height - 1) + row_bytes, B2J_E_ARG,
                "frame_stride smaller than one frame");
    g->blocks_per_frame = (long long)g->mcus_x * g->mcus_y * g->bpm;
    // bit offsets inside one frame are 32-bit: worst case 64 * 27 bits per block
    B2J_REQUIRE(g->blocks_per_frame * 1728LL < 0xFFFFFFFFLL, B2J_E_SHAPE, "frame too large for 32-bit bit offsets");
    return B2J_OK;
}

void make_quant_mul(const uint8_t *qy, const uint8_t *qc, QuantMul *qm)
{
    for (int i = 0; i < 64; i++) {
        qm->mul[0][i] = ((1u << 20) + qy[i] - 1) / qy[i];
        qm->mul[1][i] = ((1u << 20) + qc[i] - 1) / qc[i];
    }
}

bool make_quant_f32(const uint8_t *qy, const uint8_t *qc, QuantF32 *qf, double err)
{
    bool ok = true;
    const double sqrt2 = 1.41421356237309504880;
    for (int c = 0; c < 2; c++)
        for (int x = 0; x < 8; x++)
            for (int u = 0; u < 8; u++) {
                const int i = x * 8 + u, d = c ? qc[i] : qy[i];
                if (d < 2) ok = false;
                const double s = ((x == 3 || x == 5) ? sqrt2 : 1.0) * ((u == 3 || u == 5) ? sqrt2 : 1.0);
                const bool exact = (x & 3) == 0 && (u & 3) == 0;
                const double h = ((d & 1) ? 0.5 : 0.0) - (exact ? 1.0 / 16 : 0.0), e = exact ? 0.0 : err;
                qf->r[c][i] = (float)(s / (8.0 * (d ? d : 1)));
                qf->ha[c][i] = (float)((h - e) * 8.0 / s);
                qf->hb[c][i] = (float)((h + e) * 8.0 / s);
            }
    return ok;
}

static uint8_t *seg(uint8_t *p, int marker, const uint8_t *payload, int n)
{
    *p++ = 0xFF; *p++ = (uint8_t)marker; *p++ = (uint8_t)((n + 2) >> 8); *p++ = (uint8_t)((n + 2) & 0xFF);
    memcpy(p, payload, (size_t)n);
    return p + n;
}

// [ext] SOI, APP0 (JFIF 1.01), DQT x2 (zig-zag order), SOF0, DHT x4, SOS.  The
// reference only sketches this (test/model/enc_frame.py:189-200).
int jfif_header(const Geo &g, const uint8_t *qy, const uint8_t *qc, uint8_t *out, int restart_interval)
{
    uint8_t *p = out, buf[256];
    *p++ = 0xFF; *p++ = 0xD8;
    { const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0}; p = seg(p, 0xE0, app0, 14); }
    for (int t = 0; t < 2; t++) {
        const uint8_t *q = t ? qc : qy;
        buf[0] = (uint8_t)t;
        for (int i = 0; i < 64; i++) buf[1 + kZigzag[i]] = q[i];
        p = seg(p, 0xDB, buf, 65);
    }
    {
        const uint8_t sof[15] = {8, (uint8_t)(g.height >> 8), (uint8_t)g.height, (uint8_t)(g.width >> 8), (uint8_t)g.width, 3,
                                 1, (uint8_t)((g.h << 4) | g.v), 0, 2, 0x11, 1, 3, 0x11, 1};
        p = seg(p, 0xC0, sof, 15);
    }
    {
        const uint8_t *bits[4] = {kDcLumaBits, kAcLumaBits, kDcChromaBits, kAcChromaBits};
        const uint8_t *vals[4] = {kDcVals, kAcLumaVals, kDcVals, kAcChromaVals};
        const int nv[4] = {12, 162, 12, 162};
        const uint8_t id[4] = {0x00, 0x10, 0x01, 0x11};
        for (int t = 0; t < 4; t++) {
            buf[0] = id[t]; memcpy(buf + 1, bits[t], 16); memcpy(buf + 17, vals[t], (size_t)nv[t]);
            p = seg(p, 0xC4, buf, 17 + nv[t]);
        }
    }
    if (restart_interval > 0) {        // [ext] DRI, T.81 B.2.4.4
        const uint8_t dri[2] = {(uint8_t)(restart_interval >> 8), (uint8_t)restart_interval};
        p = seg(p, 0xDD, dri, 2);
    }
    { const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0}; p = seg(p, 0xDA, sos, 10); }
    return (int)(p - out);
}

}  // namespace b2j

using namespace b2j;

extern "C" const char *b2j_version(void) { return "b2j 0.1.0 (sm_100a)"; }

extern "C" size_t b2j_last_error(char *buf, size_t cap)
{
    size_t n = strlen(g_err);
    if (buf && cap) {
        size_t m = n < cap - 1 ? n : cap - 1;
        memcpy(buf, g_err, m);
        buf[m] = 0;
    }
    return n;
}

extern "C" int b2j_get_geometry(const b2j_image_desc *desc, b2j_geometry *geo)
{
    B2J_REQUIRE(geo, B2J_E_ARG, "b2j_get_geometry: null geometry");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    geo->h = g.h; geo->v = g.v; geo->mcus_x = g.mcus_x; geo->mcus_y = g.mcus_y;
    geo->blocks_per_mcu = g.bpm;
    geo->padded_width = g.mcus_x * g.mcu_w; geo->padded_height = g.mcus_y * g.mcu_h;
    geo->blocks_per_frame = g.blocks_per_frame;
    geo->header_bytes = kHeaderBytes;
    geo->reserved = 0;
    return B2J_OK;
}

extern "C" int b2j_fastpath_constants(const uint8_t qy_host[64], const uint8_t qc_host[64], float out[3][2][64])
{
    B2J_REQUIRE(qy_host && qc_host && out, B2J_E_ARG, "b2j_fastpath_constants: null pointer");
    for (int i = 0; i < 64; i++)
        B2J_REQUIRE(qy_host[i] >= 1 && qc_host[i] >= 1, B2J_E_ARG, "b2j_fastpath_constants: zero quantiser at %d", i);
    QuantF32 qf;
    const bool ok = make_quant_f32(qy_host, qc_host, &qf);
    memcpy(out[0], qf.r, sizeof qf.r);
    memcpy(out[1], qf.ha, sizeof qf.ha);
    memcpy(out[2], qf.hb, sizeof qf.hb);
    return ok ? 1 : 0;
}

extern "C" int b2j_quant_tables(int quality, uint8_t qy_host[64], uint8_t qc_host[64])
{
    B2J_REQUIRE(qy_host && qc_host, B2J_E_ARG, "b2j_quant_tables: null pointer");
    B2J_REQUIRE(quality >= 1 && quality <= 100, B2J_E_ARG, "b2j_quant_tables: quality %d not in 1..100", quality);
    const int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;    // IJG jcparam.c scaling
    for (int i = 0; i < 64; i++) {
        int a = (kBaseLuma[i] * s + 50) / 100, b = (kBaseChroma[i] * s + 50) / 100;
        qy_host[i] = (uint8_t)(a < 1 ? 1 : a > 255 ? 255 : a);
        qc_host[i] = (uint8_t)(b < 1 ? 1 : b > 255 ? 255 : b);
    }
    return B2J_OK;
}
