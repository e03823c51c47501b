// The whole encode path on device buffers (b2j_encode) and the host-buffer
// encoder object used for end-to-end measurements (b2j_encoder_*).
#include <new>

#include "b2j_common.cuh"

using namespace b2j;

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
    size_t coefs, entropy_ws, total;
};

int ws_layout(const b2j_image_desc *desc, int64_t out_cap, int opts, WsLayout *w)
{
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    B2J_REQUIRE(out_cap >= kHeaderMax + 2, B2J_E_CAPACITY, "out_cap %lld smaller than the JFIF header", (long long)out_cap);
    B2J_REQUIRE(g.n_frames <= 65535, B2J_E_SHAPE, "at most 65535 frames per call");
    const size_t nb = (size_t)g.n_frames * (size_t)g.blocks_per_frame;
    size_t o = 0;
    w->coefs = o; o = align_up(o + nb * 64 * sizeof(int16_t), 256);
    w->entropy_ws = o; o = align_up(o + b2j_entropy_workspace_bytes_opts(desc, out_cap, opts & ~B2J_OPT_DCT_FIXED), 256);
    w->total = o;
    return B2J_OK;
}

}  // namespace

extern "C" int64_t b2j_default_out_cap(const b2j_image_desc *desc)
{
    Geo g;
    if (make_geo(desc, &g)) return 0;
    // 12 bits per padded pixel plus the header: ample for photographic content at any
    // quality; i.i.d. noise at quality > 95 can need more (then sizes[] reports the need).
    const long long px = (long long)g.mcus_x * g.mcu_w * g.mcus_y * g.mcu_h;
    return (px * 3 / 2 + 4096 + 15) & ~15LL;
}

extern "C" size_t b2j_encode_workspace_bytes_opts(const b2j_image_desc *desc, int64_t out_cap, int opts)
{
    WsLayout w;
    if (ws_layout(desc, out_cap, opts, &w)) return 0;
    return w.total;
}

extern "C" size_t b2j_encode_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap)
{
    WsLayout w;
    if (ws_layout(desc, out_cap, 0, &w)) return 0;
    return w.total;
}

extern "C" int b2j_encode_launch_count(const b2j_image_desc *desc)
{
    (void)desc;
    return 4;   // k_fused_tiled, k_entropy_fused, k_ff_scan, k_stuff_write (+ one small memset)
}

extern "C" int b2j_encode(const uint8_t *rgb, const b2j_image_desc *desc, int quality, const uint8_t qy_host[64],
                          const uint8_t qc_host[64], int zrl, void *workspace, size_t workspace_bytes, uint8_t *out,
                          int64_t out_cap, int64_t *sizes, void *stream)
{
    B2J_REQUIRE(rgb && desc && workspace && out && sizes, B2J_E_ARG, "b2j_encode: null pointer");
    B2J_REQUIRE(((uintptr_t)workspace & 255) == 0, B2J_E_ALIGN, "b2j_encode: workspace must be 256-byte aligned");
    uint8_t qy[64], qc[64];
    if (quality != 0) {
        int rc = b2j_quant_tables(quality, qy, qc);
        if (rc) return rc;
    } else {
        B2J_REQUIRE(qy_host && qc_host, B2J_E_ARG, "b2j_encode: quality == 0 needs explicit tables");
        memcpy(qy, qy_host, 64);
        memcpy(qc, qc_host, 64);
    }
    WsLayout w;
    int rc = ws_layout(desc, out_cap, zrl, &w);
    if (rc) return rc;
    B2J_REQUIRE(workspace_bytes >= w.total, B2J_E_WORKSPACE, "b2j_encode: workspace %zu < %zu bytes", workspace_bytes, w.total);
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    int16_t *coefs = reinterpret_cast<int16_t *>(ws + w.coefs);
    const int variant = B2J_FUSED_AUTO | ((zrl & B2J_OPT_DCT_FIXED) ? B2J_FUSED_DCT_FIXED : 0);
    if ((rc = b2j_fused_block(rgb, desc, qy, qc, coefs, variant, stream))) return rc;
    if ((rc = b2j_entropy_encode(coefs, desc, qy, qc, zrl & ~B2J_OPT_DCT_FIXED, ws + w.entropy_ws, out, out_cap, sizes, stream)))
        return rc;
    return B2J_OK;
}

// ---------------------------------------------------------------- host-buffer encoder
// Frames are processed in chunks; each chunk has its own stream, device input,
// device output, workspace and pinned staging, so that H2D of chunk i+1, the
// kernels of chunk i and D2H of chunk i-1 overlap (2 slots = double buffering).
struct b2j_encoder {
    b2j_image_desc desc;          // n_frames = frames per chunk
    int quality, zrl;
    int64_t out_cap;
    size_t frame_bytes, ws_bytes;
    static const int kSlots = 2;
    cudaStream_t stream[kSlots];
    uint8_t *d_rgb[kSlots], *d_out[kSlots], *d_ws[kSlots];
    int64_t *d_sizes[kSlots];
    uint8_t *h_rgb[kSlots], *h_out[kSlots];      // pinned staging
    int64_t *h_sizes[kSlots];
};

static void encoder_free(b2j_encoder *e)
{
    for (int s = 0; s < b2j_encoder::kSlots; s++) {
        if (e->d_rgb[s]) cudaFree(e->d_rgb[s]);
        if (e->d_out[s]) cudaFree(e->d_out[s]);
        if (e->d_ws[s]) cudaFree(e->d_ws[s]);
        if (e->d_sizes[s]) cudaFree(e->d_sizes[s]);
        if (e->h_rgb[s]) cudaFreeHost(e->h_rgb[s]);
        if (e->h_out[s]) cudaFreeHost(e->h_out[s]);
        if (e->h_sizes[s]) cudaFreeHost(e->h_sizes[s]);
        if (e->stream[s]) cudaStreamDestroy(e->stream[s]);
    }
    delete e;
}

extern "C" int b2j_encoder_create(b2j_encoder **enc, const b2j_image_desc *desc, int quality, int zrl, int64_t out_cap,
                                  int32_t frames_per_chunk)
{
    B2J_REQUIRE(enc && desc, B2J_E_ARG, "b2j_encoder_create: null pointer");
    B2J_REQUIRE(quality >= 1 && quality <= 100, B2J_E_ARG, "b2j_encoder_create: quality %d not in 1..100", quality);
    B2J_REQUIRE((zrl & 0xFF) == B2J_ZRL_EAGER || (zrl & 0xFF) == B2J_ZRL_STANDARD, B2J_E_ARG,
                "b2j_encoder_create: bad zrl policy");
    B2J_REQUIRE(frames_per_chunk >= 1, B2J_E_ARG, "b2j_encoder_create: frames_per_chunk must be >= 1");
    B2J_REQUIRE(desc->row_stride == 0 && desc->frame_stride == 0, B2J_E_ARG, "b2j_encoder_create: host frames must be packed");
    b2j_image_desc d = *desc;
    d.n_frames = frames_per_chunk;
    Geo g;
    int rc = make_geo(&d, &g);
    if (rc) return rc;
    if (out_cap <= 0) out_cap = b2j_default_out_cap(&d);
    b2j_encoder *e = new (std::nothrow) b2j_encoder();
    B2J_REQUIRE(e, B2J_E_ARG, "b2j_encoder_create: out of host memory");
    memset(e, 0, sizeof *e);
    e->desc = d; e->quality = quality; e->zrl = zrl; e->out_cap = out_cap;
    e->frame_bytes = (size_t)g.height * (size_t)g.width * 3;
    e->ws_bytes = b2j_encode_workspace_bytes_opts(&d, out_cap, zrl);
    if (!e->ws_bytes) { encoder_free(e); return B2J_E_CAPACITY; }
    for (int s = 0; s < b2j_encoder::kSlots; s++) {
        cudaError_t err;
        if ((err = cudaStreamCreateWithFlags(&e->stream[s], cudaStreamNonBlocking)) != cudaSuccess ||
            (err = cudaMalloc(&e->d_rgb[s], e->frame_bytes * frames_per_chunk)) != cudaSuccess ||
            (err = cudaMalloc(&e->d_out[s], (size_t)out_cap * frames_per_chunk)) != cudaSuccess ||
            (err = cudaMalloc(&e->d_ws[s], e->ws_bytes)) != cudaSuccess ||
            (err = cudaMalloc(&e->d_sizes[s], sizeof(int64_t) * frames_per_chunk)) != cudaSuccess ||
            (err = cudaMallocHost(&e->h_rgb[s], e->frame_bytes * frames_per_chunk)) != cudaSuccess ||
            (err = cudaMallocHost(&e->h_out[s], (size_t)out_cap * frames_per_chunk)) != cudaSuccess ||
            (err = cudaMallocHost(&e->h_sizes[s], sizeof(int64_t) * frames_per_chunk)) != cudaSuccess) {
            encoder_free(e);
            return cuda_fail(err, "b2j_encoder_create");
        }
    }
    *enc = e;
    return B2J_OK;
}

extern "C" int b2j_encoder_destroy(b2j_encoder *enc)
{
    if (enc) encoder_free(enc);
    return B2J_OK;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Finishes the chunk held by slot `s`: once its sizes are on the host, only the bytes every frame
// really has are copied back (a slot is several times larger than a typical file), straight into
// the caller's buffer when that is pinned.  Runs while the other slot's H2D copy and kernels are
// in flight, so the D2H traffic overlaps them.
static int drain_slot(b2j_encoder *e, int s, int first_frame, int n, uint8_t *out_host, int64_t *sizes_host, bool out_pinned,
                      int *bad)
{
    B2J_CUDA(cudaStreamSynchronize(e->stream[s]));                  // kernels done, sizes on the host
    for (int i = 0; i < n; i++) {
        const int64_t sz = e->h_sizes[s][i];
        sizes_host[first_frame + i] = sz;
        if (sz < 0) { *bad = 1; continue; }
        uint8_t *dst = out_pinned ? out_host + (size_t)(first_frame + i) * (size_t)e->out_cap : e->h_out[s] + (size_t)i * (size_t)e->out_cap;
        B2J_CUDA(cudaMemcpyAsync(dst, e->d_out[s] + (size_t)i * (size_t)e->out_cap, (size_t)sz, cudaMemcpyDeviceToHost, e->stream[s]));
    }
    B2J_CUDA(cudaStreamSynchronize(e->stream[s]));
    if (!out_pinned)
        for (int i = 0; i < n; i++) {
            const int64_t sz = e->h_sizes[s][i];
            if (sz > 0)
                memcpy(out_host + (size_t)(first_frame + i) * (size_t)e->out_cap, e->h_out[s] + (size_t)i * (size_t)e->out_cap, (size_t)sz);
        }
    return B2J_OK;
}

extern "C" int b2j_encoder_encode_host(b2j_encoder *enc, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host,
                                       int64_t *sizes_host)
{
    B2J_REQUIRE(enc && rgb_host && out_host && sizes_host, B2J_E_ARG, "b2j_encoder_encode_host: null pointer");
    B2J_REQUIRE(n_frames >= 1, B2J_E_ARG, "b2j_encoder_encode_host: n_frames must be >= 1");
    b2j_encoder *e = enc;
    const int fpc = e->desc.n_frames;
    const int n_chunks = (n_frames + fpc - 1) / fpc;
    int bad = 0;
    // pinned caller buffers are used directly; pageable ones go through the pinned staging slots
    const bool in_pinned = is_pinned(rgb_host), out_pinned = is_pinned(out_host);
    int pend_first[b2j_encoder::kSlots] = {0, 0}, pend_n[b2j_encoder::kSlots] = {0, 0};
    for (int c = 0; c < n_chunks; c++) {
        const int s = c % b2j_encoder::kSlots;
        if (pend_n[s]) {                     // slot still holds an older chunk: finish it first
            int rc = drain_slot(e, s, pend_first[s], pend_n[s], out_host, sizes_host, out_pinned, &bad);
            if (rc) return rc;
            pend_n[s] = 0;
        }
        const int first = c * fpc, n = n_frames - first < fpc ? n_frames - first : fpc;
        b2j_image_desc d = e->desc;
        d.n_frames = n;
        const uint8_t *src = rgb_host + (size_t)first * e->frame_bytes;
        if (!in_pinned) {                    // the pinned hop keeps the copy asynchronous for pageable callers
            memcpy(e->h_rgb[s], src, (size_t)n * e->frame_bytes);
            src = e->h_rgb[s];
        }
        B2J_CUDA(cudaMemcpyAsync(e->d_rgb[s], src, (size_t)n * e->frame_bytes, cudaMemcpyHostToDevice, e->stream[s]));
        int rc = b2j_encode(e->d_rgb[s], &d, e->quality, nullptr, nullptr, e->zrl, e->d_ws[s], e->ws_bytes, e->d_out[s],
                            e->out_cap, e->d_sizes[s], e->stream[s]);
        if (rc) return rc;
        B2J_CUDA(cudaMemcpyAsync(e->h_sizes[s], e->d_sizes[s], sizeof(int64_t) * n, cudaMemcpyDeviceToHost, e->stream[s]));
        // the file bytes follow in drain_slot(), once the sizes are known on the host
        pend_first[s] = first;
        pend_n[s] = n;
    }
    for (int s = 0; s < b2j_encoder::kSlots; s++)
        if (pend_n[s]) {
            int rc = drain_slot(e, s, pend_first[s], pend_n[s], out_host, sizes_host, out_pinned, &bad);
            if (rc) return rc;
        }
    if (bad) {
        set_error("b2j_encoder_encode_host: at least one frame did not fit its %lld-byte slot", (long long)e->out_cap);
        return B2J_E_CAPACITY;
    }
    return B2J_OK;
}
