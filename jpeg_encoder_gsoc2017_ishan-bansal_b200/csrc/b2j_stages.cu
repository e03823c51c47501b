// K0: one un-fused kernel per reference stage operator, so that every
// test/*_tb.py-style check has a GPU target.  Throughput is not the point here
// (the fused path is in b2j_fused.cu); the arithmetic is shared with it through
// b2j_math.cuh.
#include "b2j_common.cuh"
#include "b2j_math.cuh"

namespace b2j {

__constant__ uint8_t c_zigzag[64] = B2J_ZIGZAG_INIT;

// ---- RGB2YCbCr (litejpeg/core/csc.py:41-160; model test/model/enc_frame.py:27-35)
__global__ void k_rgb2ycbcr(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ y,
                            uint8_t *__restrict__ cb, uint8_t *__restrict__ cr, long long n, int mode)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        uint32_t p = mode == B2J_CSC_JFIF ? rgb2ycbcr_px(r, g, b) : rgb2ycbcr_xapp930_px((int)r, (int)g, (int)b);
        y[i] = (uint8_t)p; cb[i] = (uint8_t)(p >> 8); cr[i] = (uint8_t)(p >> 16);
    }
}

// ---- YCbCr444to422 (litejpeg/core/crs.py:8-65)
__global__ void k_ycbcr444to422(const uint8_t *__restrict__ y, const uint8_t *__restrict__ cb,
                                const uint8_t *__restrict__ cr, uint8_t *__restrict__ yo,
                                uint8_t *__restrict__ cbcr, long long npairs)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < npairs; i += stride) {
        yo[2 * i] = y[2 * i]; yo[2 * i + 1] = y[2 * i + 1];
        cbcr[2 * i] = (uint8_t)(((int)cb[2 * i] + (int)cb[2 * i + 1]) >> 1);       // crs.py:39-54
        cbcr[2 * i + 1] = (uint8_t)(((int)cr[2 * i] + (int)cr[2 * i + 1]) >> 1);   // crs.py:57-65
    }
}

// ---- YCbCr2RGB (decode direction; litejpeg/core/csc.py:219-310 XAPP931 datapath with coef_w = 8:
// acoef 116, bcoef -8, ccoef -27, dcoef 98, products >> 6, saturate; mode JFIF is the float64
// inverse of test/common.py:172-178, int() truncation, NOT clamped - hence int16 outputs)
__global__ void k_ycbcr2rgb(const uint8_t *__restrict__ y, const uint8_t *__restrict__ cb, const uint8_t *__restrict__ cr,
                            int16_t *__restrict__ r, int16_t *__restrict__ g, int16_t *__restrict__ b, long long n, int mode)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int yy = y[i], cbm = (int)cb[i] - 128, crm = (int)cr[i] - 128;
        int rr, gg, bb;
        if (mode == B2J_CSC_XAPP930) {
            const int ym = yy - 16;
            rr = ym + ((crm * 116) >> 6);
            gg = ym + ((cbm * -8) >> 6) + ((crm * -27) >> 6);
            bb = ym + ((cbm * 98) >> 6);
            rr = min(max(rr, 0), 255); gg = min(max(gg, 0), 255); bb = min(max(bb, 0), 255);
        } else {
            const double yd = (double)yy, cbd = (double)cbm, crd = (double)crm;
            rr = __double2int_rz(__dadd_rn(yd, __dmul_rn(crd, 1.402)));
            gg = __double2int_rz(__dadd_rn(__dadd_rn(yd, __dmul_rn(cbd, -0.34414)), __dmul_rn(crd, -0.71414)));
            bb = __double2int_rz(__dadd_rn(yd, __dmul_rn(cbd, 1.772)));
        }
        r[i] = (int16_t)rr; g[i] = (int16_t)gg; b[i] = (int16_t)bb;
    }
}

// ---- YCbCr422to444 (decode direction; litejpeg/core/crs.py:88-128): sample repeat
__global__ void k_ycbcr422to444(const uint8_t *__restrict__ y, const uint8_t *__restrict__ cbcr, uint8_t *__restrict__ yo,
                                uint8_t *__restrict__ cb, uint8_t *__restrict__ cr, long long npairs)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < npairs; i += stride) {
        yo[2 * i] = y[2 * i]; yo[2 * i + 1] = y[2 * i + 1];
        cb[2 * i] = cb[2 * i + 1] = cbcr[2 * i];
        cr[2 * i] = cr[2 * i + 1] = cbcr[2 * i + 1];
    }
}

// ---- [ext] ingest: other pixel layouts -> the packed interleaved RGB the path reads.
// The reference does this on the host with PIL (enc_frame.py:7-10 takes arr[..., 0..2]; test/common.py:47-51
// unzips img.getdata()).  One thread per 4 output pixels (three 32-bit stores).
struct IngestParams {
    const uint8_t *src;
    uint8_t *dst;
    int width, height, layout;
    long long src_row_stride, src_plane_stride, src_frame_stride, n_frames;
};

__device__ __forceinline__ void ingest_px(const IngestParams &P, const uint8_t *frame, int y, int x, uint32_t &r, uint32_t &g, uint32_t &b)
{
    const uint8_t *row = frame + (long long)y * P.src_row_stride;
    switch (P.layout) {
    case B2J_LAYOUT_RGB:  r = row[3 * x]; g = row[3 * x + 1]; b = row[3 * x + 2]; break;
    case B2J_LAYOUT_BGR:  b = row[3 * x]; g = row[3 * x + 1]; r = row[3 * x + 2]; break;
    case B2J_LAYOUT_RGBA: { const uchar4 q = reinterpret_cast<const uchar4 *>(row)[x]; r = q.x; g = q.y; b = q.z; } break;
    case B2J_LAYOUT_BGRA: { const uchar4 q = reinterpret_cast<const uchar4 *>(row)[x]; b = q.x; g = q.y; r = q.z; } break;
    case B2J_LAYOUT_PLANAR_RGB:
        r = row[x]; g = row[P.src_plane_stride + x]; b = row[2 * P.src_plane_stride + x]; break;
    default: r = g = b = row[x]; break;                       // B2J_LAYOUT_GRAY
    }
}

__global__ void k_ingest(const __grid_constant__ IngestParams P)
{
    const long long groups_per_row = (P.width + 3) / 4;
    const long long n = P.n_frames * P.height * groups_per_row;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int gx = (int)(i % groups_per_row);
        const long long t = i / groups_per_row;
        const int y = (int)(t % P.height);
        const long long f = t / P.height;
        const uint8_t *frame = P.src + f * P.src_frame_stride;
        uint8_t *out = P.dst + ((f * P.height + y) * (long long)P.width + 4 * gx) * 3;
        const int npx = min(4, P.width - 4 * gx);
        uint32_t c[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t r = 0, g = 0, b = 0;
            if (k < npx) ingest_px(P, frame, y, 4 * gx + k, r, g, b);
            c[3 * k] = r; c[3 * k + 1] = g; c[3 * k + 2] = b;
        }
        if (npx == 4 && ((uintptr_t)out & 3) == 0) {
            uint32_t *o = reinterpret_cast<uint32_t *>(out);
            o[0] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
            o[1] = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
            o[2] = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
        } else {
            for (int k = 0; k < 3 * npx; k++) out[k] = (uint8_t)c[k];
        }
    }
}

// ---- [ext] plane subsampling: pair mean horizontally, then (4:2:0) vertically
__global__ void k_chroma_subsample(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                   int height, int width, int v2)
{
    const int ow = width / 2, oh = v2 ? height / 2 : height;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x, n = (long long)ow * oh;
    for (; i < n; i += stride) {
        int oy = (int)(i / ow), ox = (int)(i % ow);
        const uint8_t *p = in + (long long)(v2 ? 2 * oy : oy) * width + 2 * ox;
        int m = ((int)p[0] + (int)p[1]) >> 1;
        if (v2) m = (m + (((int)p[width] + (int)p[width + 1]) >> 1)) >> 1;
        out[i] = (uint8_t)m;
    }
}

// ---- DCT (operator litejpeg/core/new_dct6.py:400-525; arithmetic enc_frame.py:42-109)
__global__ void __launch_bounds__(128) k_dct(const int16_t *__restrict__ in, int16_t *__restrict__ out, long long nblk)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int16_t *p = in + b * 64;
    uint2 px[8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w0 |= ((uint32_t)p[x * 8 + k] & 0xFFu) << (8 * k);
            w1 |= ((uint32_t)p[x * 8 + 4 + k] & 0xFFu) << (8 * k);
        }
        px[x] = make_uint2(w0, w1);
    }
    int16_t *o = out + b * 64;
    dct_block(px, [&](int idx, uint32_t mag, bool neg) { o[idx] = (int16_t)(neg ? -(int)mag : (int)mag); });
}

// [ext] the gateware's fixed-point arithmetic (new_dct6.py:171-396, corrected; b2j_math.cuh)
__global__ void __launch_bounds__(128) k_dct_fixed(const int16_t *__restrict__ in, int16_t *__restrict__ out, long long nblk)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int16_t *p = in + b * 64;
    uint2 px[8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w0 |= ((uint32_t)p[x * 8 + k] & 0xFFu) << (8 * k);
            w1 |= ((uint32_t)p[x * 8 + 4 + k] & 0xFFu) << (8 * k);
        }
        px[x] = make_uint2(w0, w1);
    }
    int16_t *o = out + b * 64;
    dct_block_fixed(px, [&](int idx, uint32_t mag, bool neg) { o[idx] = (int16_t)(neg ? -(int)mag : (int)mag); });
}

// ---- Quantization (litejpeg/core/quantization.py:29-182; model enc_frame.py:111-129)
struct QTable { uint8_t q[64]; };
__global__ void k_quantize(const int16_t *__restrict__ in, int16_t *__restrict__ out, long long n, QTable t, int mode)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int v = in[i], d = t.q[i & 63];
        int r = mode == B2J_Q_ROUND ? quant_round(v, d) : mode == B2J_Q_TRUNC ? quant_trunc(v, d) : quant_reciprocal(v, d);
        out[i] = (int16_t)r;
    }
}

// ---- ZigZag (litejpeg/core/zigzag.py:37-212; model enc_frame.py:161-165)
__global__ void k_zigzag(const int16_t *__restrict__ in, int16_t *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[(i & ~63LL) + c_zigzag[i & 63]] = in[i];
}

// ---- RunLength / EntropyCoder / RLEMain (litejpeg/core/rle/*.py)
// One thread per block.  Slot k of the output corresponds to input index k.
// main_word = 0: 17-bit RunLength word  amp[0:12] run[12:16] valid[16]   (rlecore.py:139-146)
// main_word = 1: 21-bit RLEMain word    amp[0:12] size[12:16] run[16:20] valid[20] (rlemain.py:63-74)
__global__ void k_rle_words(const int16_t *__restrict__ in, uint32_t *__restrict__ words, long long nblk,
                            int prev_dc0, int main_word)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int16_t *c = in + b * 64;
    uint32_t *w = words + b * 64;
    const int prev = b == 0 ? prev_dc0 : (int)in[(b - 1) * 64];   // rlecore.py:86-98 prev_dc_0
    int zero_count = 0;
    for (int k = 0; k < 64; k++) {
        const int v = c[k];
        int valid = 0, run = 0, amp = 0;
        if (k == 0) { valid = 1; amp = v - prev; }
        else if (v == 0) {
            if (zero_count == 15) { valid = 1; run = 15; zero_count = 0; }    // :101-105
            else if (k == 63) { valid = 1; }                                   // :107-114
            else zero_count++;                                                 // :119-120
        } else { valid = 1; run = zero_count; amp = v; zero_count = 0; }      // :127-131
        uint32_t word;
        if (main_word) word = ((uint32_t)amp & 0xFFFu) | ((uint32_t)(bit_size(v) & 15) << 12) |
                              ((uint32_t)run << 16) | ((uint32_t)valid << 20);
        else word = ((uint32_t)amp & 0xFFFu) | ((uint32_t)run << 12) | ((uint32_t)valid << 16);
        w[k] = word;
    }
}

// The model's generic run-length code (enc_frame.py:175-176, groupby): one thread per block,
// (symbol, run) pairs flattened into pairs[b][0 .. counts[b]); at most 64 pairs = 128 values.
__global__ void k_rle_code(const int16_t *__restrict__ in, int16_t *__restrict__ pairs, int32_t *__restrict__ counts,
                           long long nblk)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int16_t *c = in + b * 64;
    int16_t *o = pairs + b * 128;
    int n = 0, sym = c[0], run = 1;
    for (int k = 1; k < 64; k++) {
        const int v = c[k];
        if (v == sym) { run++; continue; }
        o[n++] = (int16_t)sym;
        o[n++] = (int16_t)run;
        sym = v;
        run = 1;
    }
    o[n++] = (int16_t)sym;
    o[n++] = (int16_t)run;
    counts[b] = n;
}

__global__ void k_entropy_size(const int16_t *__restrict__ in, uint8_t *__restrict__ size, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) size[i] = (uint8_t)bit_size(in[i]);
}

static inline int grid_for(long long n, int block, int cap = 148 * 16)
{
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace b2j

using namespace b2j;

extern "C" int b2j_rgb2ycbcr(const uint8_t *rgb, uint8_t *y, uint8_t *cb, uint8_t *cr, int64_t n_pixels, int mode, void *stream)
{
    B2J_REQUIRE(rgb && y && cb && cr, B2J_E_ARG, "b2j_rgb2ycbcr: null pointer");
    B2J_REQUIRE(n_pixels >= 0, B2J_E_ARG, "b2j_rgb2ycbcr: negative size");
    B2J_REQUIRE(mode == B2J_CSC_JFIF || mode == B2J_CSC_XAPP930, B2J_E_ARG, "b2j_rgb2ycbcr: bad mode %d", mode);
    if (n_pixels == 0) return B2J_OK;
    k_rgb2ycbcr<<<grid_for(n_pixels, 256), 256, 0, as_stream(stream)>>>(rgb, y, cb, cr, n_pixels, mode);
    B2J_LAUNCH_CHECK("k_rgb2ycbcr");
    return B2J_OK;
}

extern "C" int b2j_ycbcr444to422(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *y_out,
                                 uint8_t *cb_cr_out, int64_t n_pixels, void *stream)
{
    B2J_REQUIRE(y && cb && cr && y_out && cb_cr_out, B2J_E_ARG, "b2j_ycbcr444to422: null pointer");
    B2J_REQUIRE(n_pixels >= 0 && n_pixels % 2 == 0, B2J_E_ARG, "b2j_ycbcr444to422: n_pixels must be even and >= 0");
    if (n_pixels == 0) return B2J_OK;
    k_ycbcr444to422<<<grid_for(n_pixels / 2, 256), 256, 0, as_stream(stream)>>>(y, cb, cr, y_out, cb_cr_out, n_pixels / 2);
    B2J_LAUNCH_CHECK("k_ycbcr444to422");
    return B2J_OK;
}

extern "C" int b2j_ycbcr2rgb(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, int16_t *r, int16_t *g, int16_t *b,
                             int64_t n_pixels, int mode, void *stream)
{
    B2J_REQUIRE(y && cb && cr && r && g && b, B2J_E_ARG, "b2j_ycbcr2rgb: null pointer");
    B2J_REQUIRE(n_pixels >= 0, B2J_E_ARG, "b2j_ycbcr2rgb: negative size");
    B2J_REQUIRE(mode == B2J_CSC_JFIF || mode == B2J_CSC_XAPP930, B2J_E_ARG, "b2j_ycbcr2rgb: bad mode %d", mode);
    if (n_pixels == 0) return B2J_OK;
    k_ycbcr2rgb<<<grid_for(n_pixels, 256), 256, 0, as_stream(stream)>>>(y, cb, cr, r, g, b, n_pixels, mode);
    B2J_LAUNCH_CHECK("k_ycbcr2rgb");
    return B2J_OK;
}

extern "C" int b2j_ycbcr422to444(const uint8_t *y, const uint8_t *cb_cr, uint8_t *y_out, uint8_t *cb_out, uint8_t *cr_out,
                                 int64_t n_pixels, void *stream)
{
    B2J_REQUIRE(y && cb_cr && y_out && cb_out && cr_out, B2J_E_ARG, "b2j_ycbcr422to444: null pointer");
    B2J_REQUIRE(n_pixels >= 0 && n_pixels % 2 == 0, B2J_E_ARG, "b2j_ycbcr422to444: n_pixels must be even and >= 0");
    if (n_pixels == 0) return B2J_OK;
    k_ycbcr422to444<<<grid_for(n_pixels / 2, 256), 256, 0, as_stream(stream)>>>(y, cb_cr, y_out, cb_out, cr_out, n_pixels / 2);
    B2J_LAUNCH_CHECK("k_ycbcr422to444");
    return B2J_OK;
}

extern "C" int b2j_chroma_subsample(const uint8_t *in, uint8_t *out, int32_t height, int32_t width, int32_t subsampling,
                                    void *stream)
{
    B2J_REQUIRE(in && out, B2J_E_ARG, "b2j_chroma_subsample: null pointer");
    B2J_REQUIRE(subsampling == B2J_SUB_422 || subsampling == B2J_SUB_420, B2J_E_ARG, "b2j_chroma_subsample: 422 or 420 only");
    B2J_REQUIRE(height > 0 && width > 0 && width % 2 == 0 && (subsampling == B2J_SUB_422 || height % 2 == 0), B2J_E_SHAPE,
                "b2j_chroma_subsample: sides must be positive and even along subsampled axes");
    const int v2 = subsampling == B2J_SUB_420;
    long long n = (long long)(width / 2) * (v2 ? height / 2 : height);
    k_chroma_subsample<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(in, out, height, width, v2);
    B2J_LAUNCH_CHECK("k_chroma_subsample");
    return B2J_OK;
}

extern "C" int b2j_ingest(const uint8_t *src, int layout, int32_t width, int32_t height, int32_t n_frames,
                          int64_t src_row_stride, int64_t src_plane_stride, int64_t src_frame_stride, uint8_t *rgb,
                          void *stream)
{
    B2J_REQUIRE(src && rgb, B2J_E_ARG, "b2j_ingest: null pointer");
    B2J_REQUIRE(layout >= B2J_LAYOUT_RGB && layout <= B2J_LAYOUT_GRAY, B2J_E_ARG, "b2j_ingest: bad layout %d", layout);
    B2J_REQUIRE(width >= 1 && height >= 1 && n_frames >= 1 && width <= 65535 && height <= 65535, B2J_E_SHAPE,
                "b2j_ingest: bad shape %d x %d x %d", n_frames, height, width);
    const int bpp = layout <= B2J_LAYOUT_BGR ? 3 : (layout <= B2J_LAYOUT_BGRA ? 4 : 1);
    if (src_row_stride == 0) src_row_stride = (int64_t)width * bpp;
    if (src_plane_stride == 0) src_plane_stride = src_row_stride * height;
    if (src_frame_stride == 0) src_frame_stride = layout == B2J_LAYOUT_PLANAR_RGB ? 3 * src_plane_stride : src_row_stride * height;
    B2J_REQUIRE(src_row_stride >= (int64_t)width * bpp, B2J_E_SHAPE, "b2j_ingest: row stride smaller than a row");
    if (layout == B2J_LAYOUT_RGBA || layout == B2J_LAYOUT_BGRA)
        B2J_REQUIRE(((uintptr_t)src & 3) == 0 && (src_row_stride & 3) == 0 && (src_frame_stride & 3) == 0, B2J_E_ALIGN,
                    "b2j_ingest: 4-byte pixels need 4-byte aligned rows");
    IngestParams P;
    P.src = src; P.dst = rgb; P.width = width; P.height = height; P.layout = layout;
    P.src_row_stride = src_row_stride; P.src_plane_stride = src_plane_stride; P.src_frame_stride = src_frame_stride;
    P.n_frames = n_frames;
    const long long n = (long long)n_frames * height * ((width + 3) / 4);
    k_ingest<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(P);
    B2J_LAUNCH_CHECK("k_ingest");
    return B2J_OK;
}

extern "C" int b2j_dct(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream)
{
    B2J_REQUIRE(in && out, B2J_E_ARG, "b2j_dct: null pointer");
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "b2j_dct: negative size");
    if (n_blocks == 0) return B2J_OK;
    k_dct<<<(unsigned)((n_blocks + 127) / 128), 128, 0, as_stream(stream)>>>(in, out, n_blocks);
    B2J_LAUNCH_CHECK("k_dct");
    return B2J_OK;
}

extern "C" int b2j_dct_fixed(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream)
{
    B2J_REQUIRE(in && out, B2J_E_ARG, "b2j_dct_fixed: null pointer");
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "b2j_dct_fixed: negative size");
    if (n_blocks == 0) return B2J_OK;
    k_dct_fixed<<<(unsigned)((n_blocks + 127) / 128), 128, 0, as_stream(stream)>>>(in, out, n_blocks);
    B2J_LAUNCH_CHECK("k_dct_fixed");
    return B2J_OK;
}

extern "C" int b2j_quantize(const int16_t *in, int16_t *out, int64_t n_blocks, const uint8_t table_host[64], int mode,
                            void *stream)
{
    B2J_REQUIRE(in && out && table_host, B2J_E_ARG, "b2j_quantize: null pointer");
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "b2j_quantize: negative size");
    B2J_REQUIRE(mode >= B2J_Q_ROUND && mode <= B2J_Q_RECIPROCAL, B2J_E_ARG, "b2j_quantize: bad mode %d", mode);
    QTable t;
    for (int i = 0; i < 64; i++) {
        B2J_REQUIRE(table_host[i] >= 1, B2J_E_ARG, "b2j_quantize: table entry %d is zero", i);
        t.q[i] = table_host[i];
    }
    if (n_blocks == 0) return B2J_OK;
    k_quantize<<<grid_for(n_blocks * 64, 256), 256, 0, as_stream(stream)>>>(in, out, n_blocks * 64, t, mode);
    B2J_LAUNCH_CHECK("k_quantize");
    return B2J_OK;
}

extern "C" int b2j_zigzag(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream)
{
    B2J_REQUIRE(in && out && in != out, B2J_E_ARG, "b2j_zigzag: null or aliased pointers");
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "b2j_zigzag: negative size");
    if (n_blocks == 0) return B2J_OK;
    k_zigzag<<<grid_for(n_blocks * 64, 256), 256, 0, as_stream(stream)>>>(in, out, n_blocks * 64);
    B2J_LAUNCH_CHECK("k_zigzag");
    return B2J_OK;
}

static int rle_words(const int16_t *in, uint32_t *words, int64_t n_blocks, int32_t prev_dc, int main_word, void *stream,
                     const char *who)
{
    B2J_REQUIRE(in && words, B2J_E_ARG, "%s: null pointer", who);
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "%s: negative size", who);
    if (n_blocks == 0) return B2J_OK;
    k_rle_words<<<(unsigned)((n_blocks + 127) / 128), 128, 0, as_stream(stream)>>>(in, words, n_blocks, prev_dc, main_word);
    B2J_LAUNCH_CHECK(who);
    return B2J_OK;
}

extern "C" int b2j_run_length(const int16_t *in, uint32_t *words, int64_t n_blocks, int32_t prev_dc, void *stream)
{
    return rle_words(in, words, n_blocks, prev_dc, 0, stream, "b2j_run_length");
}

extern "C" int b2j_rle_main(const int16_t *in, uint32_t *words, int64_t n_blocks, int32_t prev_dc, void *stream)
{
    return rle_words(in, words, n_blocks, prev_dc, 1, stream, "b2j_rle_main");
}

extern "C" int b2j_rle_code(const int16_t *in, int16_t *pairs, int32_t *counts, int64_t n_blocks, void *stream)
{
    B2J_REQUIRE(in && pairs && counts, B2J_E_ARG, "b2j_rle_code: null pointer");
    B2J_REQUIRE(n_blocks >= 0, B2J_E_ARG, "b2j_rle_code: negative size");
    if (n_blocks == 0) return B2J_OK;
    k_rle_code<<<(unsigned)((n_blocks + 127) / 128), 128, 0, as_stream(stream)>>>(in, pairs, counts, n_blocks);
    B2J_LAUNCH_CHECK("k_rle_code");
    return B2J_OK;
}

extern "C" int b2j_entropy_size(const int16_t *in, uint8_t *size, int64_t n, void *stream)
{
    B2J_REQUIRE(in && size, B2J_E_ARG, "b2j_entropy_size: null pointer");
    B2J_REQUIRE(n >= 0, B2J_E_ARG, "b2j_entropy_size: negative size");
    if (n == 0) return B2J_OK;
    k_entropy_size<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(in, size, n);
    B2J_LAUNCH_CHECK("k_entropy_size");
    return B2J_OK;
}
