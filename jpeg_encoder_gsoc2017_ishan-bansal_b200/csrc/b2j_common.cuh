// Shared declarations of libb2j: error plumbing, geometry, constant tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b2j.h"

namespace b2j {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define B2J_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ::b2j::cuda_fail(e__, #call);          \
    } while (0)

#define B2J_REQUIRE(cond, code, ...)                                          \
    do {                                                                      \
        if (!(cond)) { ::b2j::set_error(__VA_ARGS__); return (code); }        \
    } while (0)

#define B2J_LAUNCH_CHECK(name)                                                \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) return ::b2j::cuda_fail(e__, name);           \
    } while (0)

// ---------------------------------------------------------------- launch geometry of persistent kernels
// Device attributes and occupancy are looked up once per device and kernel instantiation (the
// single-frame path is launch-latency bound: no attribute queries on the hot path).
struct LaunchInfo {
    int sms = 0, per_sm = 0;
};
static const int kMaxDevices = 64;

template <class Kernel>
static int launch_info(Kernel kernel, int threads, int smem, LaunchInfo *cache /*[kMaxDevices]*/, LaunchInfo *out)
{
    int dev = 0;
    B2J_CUDA(cudaGetDevice(&dev));
    const bool cached = dev >= 0 && dev < kMaxDevices;
    LaunchInfo li = cached ? cache[dev] : LaunchInfo();
    if (li.sms == 0) {
        B2J_CUDA(cudaDeviceGetAttribute(&li.sms, cudaDevAttrMultiProcessorCount, dev));
        B2J_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        B2J_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&li.per_sm, kernel, threads, smem));
        if (li.per_sm < 1) li.per_sm = 1;
        if (cached) cache[dev] = li;                               // benign race: every writer stores the same values
    }
    *out = li;
    return B2J_OK;
}

// ---------------------------------------------------------------- programmatic dependent launch
// The kernels of one encode follow each other on a stream.  A kernel launched with launch_pdl(..., true, ...) may become
// resident while its predecessor drains (that one calls pdl_launch_dependents() first thing) and runs its prologue; it must
// call pdl_wait() before it touches anything the predecessor writes.  Both instructions do nothing in a kernel launched the
// ordinary way.  One 1080p frame: the launch gaps between K1, the entropy kernel and the stuffing kernel are 10 % of the step.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                     Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// The two halves of b2j_entropy_encode for b2j_encode (b2j_entropy_fused.cu): the workspace head (ticket, status words, 0xFF
// counts) is cleared BEFORE K1 is launched, so that the entropy kernel follows K1 directly on the stream and can be launched
// as its programmatic dependent.  entropy_try_clear returns false (and does nothing) when the arguments are not valid: the
// entropy call itself then reports the error.
bool entropy_try_clear(const b2j_image_desc *desc, int64_t out_cap, int opts, void *workspace, cudaStream_t st);
int entropy_encode_impl(const int16_t *coefs, const b2j_image_desc *desc, const uint8_t qy_host[64], const uint8_t qc_host[64],
                        int opts, void *workspace, uint8_t *out, int64_t out_cap, int64_t *sizes, void *stream, bool after_k1);

// ---------------------------------------------------------------- geometry
struct Geo {
    int width, height;            // source pixels
    int h, v;                     // luma sampling factors
    int mcu_w, mcu_h;             // MCU size in pixels
    int mcus_x, mcus_y;
    int bpm;                      // blocks per MCU
    int n_frames;
    int layout, bpp;              // B2J_LAYOUT_RGB | BGR | RGBA | BGRA of the source frames; bytes per pixel
    long long row_stride, frame_stride;   // bytes
    long long blocks_per_frame;
};

int make_geo(const b2j_image_desc *d, Geo *g);

static const int kHeaderBytes = 623;   // SOI APP0 DQTx2 SOF0 DHTx4 SOS
static const int kHeaderMax = 640;     // + DRI (6 bytes) when a restart interval is set

// Quantisation multipliers for the fused kernel: q = (|c| * mul + 2^19) >> 20
// reproduces round-half-away(|c| / d) for every |c| <= 1024, d in 1..255
// (checked exhaustively, every |c| <= 1024 and d in 1..255, in tests/test_kernel_identities.py).
struct QuantMul {
    uint32_t mul[2][64];          // [luma|chroma][natural index]
};
void make_quant_mul(const uint8_t *qy, const uint8_t *qc, QuantMul *qm);

// Constants of the single-precision fast path of K1 (b2j_fused_pair.cuh), per natural index x * 8 + u, in units of
// the unscaled transform output `out` (v = out * s, s = sqrt2 for each of x, u in {3, 5}):
//   r  = s / (8 d)                       y = sign * (|out| - H) * r, q = rint(y)
//   ha = (h - e) * 8 / s, hb = (h + e) * 8 / s    the two ends of the error interval of |v / 8|
// h = 1/2 for odd d (int() truncates before the model rounds), 0 for even d, less 1/16 for the four coefficients
// that are exact multiples of 1/8 (e = 0 for them); e = kFastPathErr.  tests/test_fastpath_model.py derives the
// bound and restates the arithmetic.  make_quant_f32 returns false when a divisor is 1 (fast path unsupported).
static const double kFastPathErr = 1.0 / 1024;
struct QuantF32 {
    float r[2][64], ha[2][64], hb[2][64];
};
bool make_quant_f32(const uint8_t *qy, const uint8_t *qc, QuantF32 *qf, double err = kFastPathErr);

// ---------------------------------------------------------------- tables (host + device)
// Zig-zag position of natural index i (enc_frame.py:131-139 / zigzag.py:19-27).
#define B2J_ZIGZAG_INIT                                                         \
    { 0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42,                 \
      3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,             \
      10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,          \
      21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 }

extern const uint8_t kZigzag[64];          // natural index -> zig-zag position
extern const uint8_t kBaseLuma[64];        // Annex K.1
extern const uint8_t kBaseChroma[64];      // Annex K.2

// Annex K.3 typical Huffman tables (BITS, HUFFVAL)
extern const uint8_t kDcLumaBits[16], kDcChromaBits[16], kDcVals[12];
extern const uint8_t kAcLumaBits[16], kAcChromaBits[16], kAcLumaVals[162], kAcChromaVals[162];

// Device-side code tables: entry = (length << 16) | code.
struct HuffLut {
    uint32_t dc[2][16];           // [luma|chroma][size category]
    uint32_t ac[2][256];          // [luma|chroma][(run << 4) | size]
};
const HuffLut &host_huff_lut();

int jfif_header(const Geo &g, const uint8_t *qy, const uint8_t *qc, uint8_t *out /* >= kHeaderMax */, int restart_interval = 0);

static inline void *as_ptr(void *p) { return p; }
static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace b2j
