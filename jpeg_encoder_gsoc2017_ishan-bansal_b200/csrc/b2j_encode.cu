// The whole encode path on device buffers (b2j_encode) and the host-buffer
// encoder object used for end-to-end measurements (b2j_encoder_*).
#include <condition_variable>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

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
    return 3;   // k_fused_pair (or _stream), k_entropy_fused, k_stuff_write (+ one small memset; + k_ff_scan with restart intervals)
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
    // the entropy workspace is cleared in front of K1: the three kernels then follow each other directly on the stream
    const bool cleared = entropy_try_clear(desc, out_cap, zrl & ~B2J_OPT_DCT_FIXED, ws + w.entropy_ws, reinterpret_cast<cudaStream_t>(stream));
    if ((rc = b2j_fused_block(rgb, desc, qy, qc, coefs, variant, stream))) return rc;
    if ((rc = entropy_encode_impl(coefs, desc, qy, qc, zrl & ~B2J_OPT_DCT_FIXED, ws + w.entropy_ws, out, out_cap, sizes, stream, cleared)))
        return rc;
    return B2J_OK;
}

// ---------------------------------------------------------------- host-buffer encoder
// Frames are processed in chunks; each chunk has its own stream, device input,
// device output, workspace and pinned staging, so that H2D of chunk i+1, the
// kernels of chunk i and D2H of chunk i-1 overlap (2 slots = double buffering).
struct b2j_encoder {
    int device;                   // the CUDA device everything below lives on
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

// Makes `device` current for the lifetime of the object (the encoder's buffers and streams
// belong to the device it was created on; callers may have another one current).
struct DeviceGuard {
    int saved = -1;
    bool switched = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&saved) == cudaSuccess && saved != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(saved); }
};

static void encoder_free(b2j_encoder *e)
{
    DeviceGuard guard(e->device);
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
    if (cudaGetDevice(&e->device) != cudaSuccess) { delete e; return cuda_fail(cudaGetLastError(), "b2j_encoder_create"); }
    e->desc = d; e->quality = quality; e->zrl = zrl; e->out_cap = out_cap;
    e->frame_bytes = (size_t)g.height * (size_t)g.width * (size_t)g.bpp;
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

static int encode_host_impl(b2j_encoder *e, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host, int64_t *sizes_host);

extern "C" int b2j_encoder_encode_host(b2j_encoder *enc, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host,
                                       int64_t *sizes_host)
{
    B2J_REQUIRE(enc && rgb_host && out_host && sizes_host, B2J_E_ARG, "b2j_encoder_encode_host: null pointer");
    B2J_REQUIRE(n_frames >= 1, B2J_E_ARG, "b2j_encoder_encode_host: n_frames must be >= 1");
    DeviceGuard guard(enc->device);
    for (int i = 0; i < n_frames; i++) sizes_host[i] = INT64_MIN;   // frames not reached after an error stay marked
    const int rc = encode_host_impl(enc, rgb_host, n_frames, out_host, sizes_host);
    if (rc != B2J_OK && rc != B2J_E_CAPACITY) {
        // error in the middle of the pipeline: queued copies may still read rgb_host / write out_host
        for (int s = 0; s < b2j_encoder::kSlots; s++) cudaStreamSynchronize(enc->stream[s]);
    }
    return rc;
}

static int encode_host_impl(b2j_encoder *e, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host, int64_t *sizes_host)
{
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

// ---------------------------------------------------------------- one call, several GPUs
// Frames [g * n / G, (g + 1) * n / G) of the host batch go to the g-th device of the list (the
// contiguous-shard rule of SURVEY 8e; the unit that shards is the per-image loop
// test/model/enc_frame.py:203-238).  One persistent host thread per device drives a b2j_encoder
// there, so the H2D copies, kernels and D2H copies of all devices run concurrently; there is no
// exchange between the devices.
struct b2j_multi_encoder {
    struct Worker {
        int device = 0;
        b2j_encoder *enc = nullptr;
        std::thread thread;
        // job (guarded by the owner's mutex)
        int job = 0;                  // 0 idle, 1 encode, 2 copy probe, 3 quit
        const uint8_t *rgb = nullptr;
        uint8_t *out = nullptr;
        int64_t *sizes = nullptr;
        int n = 0, repeats = 0;
        int rc = 0;
        char err[256] = "";
    };
    std::vector<Worker> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    int pending = 0;
    size_t frame_bytes = 0;
    int64_t out_cap = 0;
};

static void multi_worker_main(b2j_multi_encoder *m, int idx)
{
    b2j_multi_encoder::Worker &w = m->workers[idx];
    cudaSetDevice(w.device);
    for (;;) {
        int job;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_job.wait(lk, [&] { return w.job != 0; });
            job = w.job;
        }
        int rc = B2J_OK;
        if (job == 1) {
            rc = w.n > 0 ? b2j_encoder_encode_host(w.enc, w.rgb, w.n, w.out, w.sizes) : B2J_OK;
        } else if (job == 2) {
            // the H2D copies encode_host would issue, and nothing else
            b2j_encoder *e = w.enc;
            const int fpc = e->desc.n_frames;
            for (int r = 0; r < w.repeats && rc == B2J_OK; r++) {
                for (int first = 0, c = 0; first < w.n; first += fpc, c++) {
                    const int n = w.n - first < fpc ? w.n - first : fpc;
                    const int s = c % b2j_encoder::kSlots;
                    cudaError_t e2 = cudaMemcpyAsync(e->d_rgb[s], w.rgb + (size_t)first * e->frame_bytes, (size_t)n * e->frame_bytes,
                                                     cudaMemcpyHostToDevice, e->stream[s]);
                    if (e2 != cudaSuccess) { rc = cuda_fail(e2, "b2j_multi_encoder_h2d_probe"); break; }
                }
                for (int s = 0; s < b2j_encoder::kSlots; s++) cudaStreamSynchronize(e->stream[s]);
            }
        }
        if (rc) b2j_last_error(w.err, sizeof w.err);
        {
            std::lock_guard<std::mutex> lk(m->mu);
            w.rc = rc;
            w.job = 0;
            m->pending--;
        }
        m->cv_done.notify_all();
        if (job == 3) return;
    }
}

// Hands `job` to every worker and waits for all of them; returns the first non-zero return code.
static int multi_run(b2j_multi_encoder *m, int job)
{
    {
        std::lock_guard<std::mutex> lk(m->mu);
        for (auto &w : m->workers) w.job = job;
        m->pending = (int)m->workers.size();
    }
    m->cv_job.notify_all();
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->pending == 0; });
    int rc = B2J_OK;
    for (auto &w : m->workers)
        if (w.rc && (rc == B2J_OK || rc == B2J_E_CAPACITY)) {
            rc = w.rc;
            set_error("device %d: %s", w.device, w.err);
        }
    return rc;
}

static void multi_shard(b2j_multi_encoder *m, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host, int64_t *sizes_host)
{
    const long long G = (long long)m->workers.size();
    for (long long g = 0; g < G; g++) {
        const long long lo = g * n_frames / G, hi = (g + 1) * n_frames / G;
        b2j_multi_encoder::Worker &w = m->workers[(size_t)g];
        w.rgb = rgb_host + (size_t)lo * m->frame_bytes;
        w.out = out_host ? out_host + (size_t)lo * (size_t)m->out_cap : nullptr;
        w.sizes = sizes_host ? sizes_host + lo : nullptr;
        w.n = (int)(hi - lo);
    }
}

extern "C" int b2j_multi_encoder_destroy(b2j_multi_encoder *enc)
{
    if (!enc) return B2J_OK;
    bool threads = false;
    for (auto &w : enc->workers) threads = threads || w.thread.joinable();
    if (threads) {
        multi_run(enc, 3);
        for (auto &w : enc->workers)
            if (w.thread.joinable()) w.thread.join();
    }
    for (auto &w : enc->workers)
        if (w.enc) encoder_free(w.enc);
    delete enc;
    return B2J_OK;
}

extern "C" int b2j_multi_encoder_create(b2j_multi_encoder **enc, const int32_t *devices, int32_t n_devices,
                                        const b2j_image_desc *desc, int quality, int zrl, int64_t out_cap, int32_t frames_per_chunk)
{
    B2J_REQUIRE(enc && devices && desc, B2J_E_ARG, "b2j_multi_encoder_create: null pointer");
    B2J_REQUIRE(n_devices >= 1 && n_devices <= 64, B2J_E_ARG, "b2j_multi_encoder_create: n_devices %d not in 1..64", n_devices);
    int count = 0;
    B2J_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < n_devices; i++)
        B2J_REQUIRE(devices[i] >= 0 && devices[i] < count, B2J_E_ARG, "b2j_multi_encoder_create: device %d of %d", devices[i], count);
    b2j_multi_encoder *m = new (std::nothrow) b2j_multi_encoder();
    B2J_REQUIRE(m, B2J_E_ARG, "b2j_multi_encoder_create: out of host memory");
    m->workers.resize((size_t)n_devices);
    int saved = 0;
    cudaGetDevice(&saved);
    int rc = B2J_OK;
    for (int i = 0; i < n_devices && rc == B2J_OK; i++) {         // (a device may be listed more than once: two workers share it)
        m->workers[(size_t)i].device = devices[i];
        cudaError_t e = cudaSetDevice(devices[i]);
        rc = e != cudaSuccess ? cuda_fail(e, "cudaSetDevice") : b2j_encoder_create(&m->workers[(size_t)i].enc, desc, quality, zrl, out_cap, frames_per_chunk);
    }
    cudaSetDevice(saved);
    if (rc) { b2j_multi_encoder_destroy(m); return rc; }
    m->frame_bytes = m->workers[0].enc->frame_bytes;
    m->out_cap = m->workers[0].enc->out_cap;
    for (int i = 0; i < n_devices; i++) m->workers[(size_t)i].thread = std::thread(multi_worker_main, m, i);
    *enc = m;
    return B2J_OK;
}

extern "C" int b2j_multi_encoder_encode_host(b2j_multi_encoder *enc, const uint8_t *rgb_host, int32_t n_frames, uint8_t *out_host,
                                             int64_t *sizes_host)
{
    B2J_REQUIRE(enc && rgb_host && out_host && sizes_host, B2J_E_ARG, "b2j_multi_encoder_encode_host: null pointer");
    B2J_REQUIRE(n_frames >= 1, B2J_E_ARG, "b2j_multi_encoder_encode_host: n_frames must be >= 1");
    multi_shard(enc, rgb_host, n_frames, out_host, sizes_host);
    return multi_run(enc, 1);
}

extern "C" int b2j_multi_encoder_h2d_probe(b2j_multi_encoder *enc, const uint8_t *rgb_host, int32_t n_frames, int32_t repeats,
                                           double *seconds_per_pass)
{
    B2J_REQUIRE(enc && rgb_host && seconds_per_pass, B2J_E_ARG, "b2j_multi_encoder_h2d_probe: null pointer");
    B2J_REQUIRE(n_frames >= 1 && repeats >= 1, B2J_E_ARG, "b2j_multi_encoder_h2d_probe: n_frames and repeats must be >= 1");
    multi_shard(enc, rgb_host, n_frames, nullptr, nullptr);
    for (auto &w : enc->workers) w.repeats = repeats;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = multi_run(enc, 2);
    *seconds_per_pass = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / repeats;
    return rc;
}
