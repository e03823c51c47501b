// Shared pieces of the entropy kernels (K2..K4 and the fused single-pass coder).
#pragma once
#include "b2j_common.cuh"
#include "b2j_math.cuh"

namespace b2j {

struct EGeo {
    int hv, bpm;                  // luma blocks per MCU, blocks per MCU
    int n_frames;
    long long nblk;               // blocks per frame
};

static inline EGeo egeo_of(const Geo &g)
{
    EGeo e;
    e.hv = g.h * g.v; e.bpm = g.bpm; e.n_frames = g.n_frames; e.nblk = g.blocks_per_frame;
    return e;
}

struct LutParam { HuffLut lut; };

__device__ __forceinline__ void load_lut(const LutParam &L, uint32_t *s_dc /*[2][16]*/, uint32_t *s_ac /*[2][256]*/)
{
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s_dc[i] = L.lut.dc[i >> 4][i & 15];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_ac[i] = L.lut.ac[i >> 8][i & 255];
}

// DC of the previous block of the same component inside the frame (0 if none).
__device__ __forceinline__ int prev_dc(const int16_t *frame_coefs, long long b, const EGeo &e)
{
    const int k = (int)(b % e.bpm);
    long long p;
    if (k < e.hv) p = k > 0 ? b - 1 : b - e.bpm + e.hv - 1;
    else p = b - e.bpm;
    return p >= 0 ? (int)frame_coefs[p * 64] : 0;
}

// Walks one block in zig-zag order and hands every coded field to sink.put(value, nbits).
template <bool EAGER, class Sink>
__device__ __forceinline__ void walk_block(const int16_t *__restrict__ c, int diff, const uint32_t *__restrict__ lut_dc,
                                           const uint32_t *__restrict__ lut_ac, Sink &sink)
{
    {
        const int s = bit_size(diff);
        const uint32_t e = lut_dc[s];
        const uint32_t extra = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u);
        sink.put(((e & 0xFFFFu) << s) | extra, (int)(e >> 16) + s);
    }
    const uint32_t zrl = lut_ac[0xF0], eob = lut_ac[0x00];
    int run = 0;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(c);
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        const uint4 q = __ldg(c4 + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = i * 8 + j;
            if (k == 0) continue;
            const int v = (int)(int16_t)(w[j >> 1] >> (16 * (j & 1)));
            uint32_t val = 0u;
            int len = 0;
            if (v != 0) {
                if (!EAGER) {
                    while (run > 15) { sink.put(zrl & 0xFFFFu, (int)(zrl >> 16)); run -= 16; }
                }
                const int s = bit_size(v);
                const uint32_t e = lut_ac[(run << 4) | s];
                val = ((e & 0xFFFFu) << s) | ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << s) - 1u));
                len = (int)(e >> 16) + s;
                run = 0;
            } else if (EAGER && run == 15) {
                val = zrl & 0xFFFFu; len = (int)(zrl >> 16); run = 0;
            } else if (k == 63) {
                val = eob & 0xFFFFu; len = (int)(eob >> 16);
            } else {
                run++;
            }
            sink.put(val, len);
        }
    }
}

struct CountSink {
    int bits;
    __device__ __forceinline__ void put(uint32_t, int len) { bits += len; }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// exclusive scan over the CTA (256 threads); returns exclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t cta_excl_scan_256(uint32_t v, uint32_t *s_warp /*[8]*/, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}


static constexpr int kStuffChunk = 4096;         // unstuffed bytes per CTA of the stuffing kernels
struct HeaderParam { uint8_t bytes[kHeaderMax]; int len; };

static inline LutParam make_lut_param()
{
    LutParam L;
    L.lut = host_huff_lut();
    return L;
}

static inline long long ucap_of(int64_t out_cap) { return ((long long)out_cap + 15) & ~15LL; }
static inline int stuff_chunks(long long ucap) { return (int)((ucap + kStuffChunk - 1) / kStuffChunk); }

// 0xFF stuffing + JFIF framing of n_frames unstuffed streams (b2j_entropy.cu).
// error_flag may be null; when *error_flag != 0 every sizes[f] is set to INT64_MIN.
// counted: ffpart already holds the 0xFF count of every 4096-byte chunk (the fused entropy kernel
// counts while it composes the words), so the counting pass is skipped.
// nseg > 1 ([ext] restart intervals): every frame is nseg consecutive, independently packed
// streams ("segments": ucap, total_bits and ffpart are per segment, n_frames * nseg of them); they
// are joined with RSTm markers between them, m = (segment + rst_phase) mod 8.
int launch_stuffing(const uint8_t *unstuffed, const int64_t *total_bits, uint32_t *ffpart, int64_t *sizes, uint8_t *out,
                    long long ucap, long long out_cap, int n_frames, const HeaderParam &H, const int *error_flag,
                    bool counted, cudaStream_t st, int nseg = 1, int rst_phase = 0, bool pdl = false);

}  // namespace b2j
