// [ext] PNG ingest on the device: the step in front of the encode path.
// replaces: PIL's Image.open + array conversion on the host, test/model/enc_frame.py:7-10 and test/common.py:47-51.
//
// The host only walks the chunk list of every file (IHDR, IDAT payloads gathered into one zlib stream per file,
// RFC 1950) and copies those streams to the device.  Then
//   k_png_inflate  one warp per file: lane 0 decodes the DEFLATE stream (RFC 1951: stored, fixed and dynamic Huffman
//                  blocks; 10-bit / 8-bit first-level tables in shared memory, canonical walk for longer codes), the
//                  32 KB window lives in shared memory, matches are copied by the whole warp, and the window is
//                  written to the file's scanline buffer in coalesced 16 KB pieces;
//   k_png_unfilter one warp per file: the scanline filters (PNG 1.2 section 6: None, Sub, Up, Average, Paeth) are
//                  undone for 32 rows at a time in a skewed wavefront (lane = row, one pixel to the right per step; the
//                  pixel above arrives by a shuffle from the lane of the row above), and the packed RGB frame the
//                  encode path reads is written (alpha dropped as enc_frame.py:8-10 does, grey replicated).
// A file is a serial bit stream, so one file decodes at the speed of one thread (tens of milliseconds for a
// photographic 512 x 512 PNG); the throughput comes from the batch: a thousand files are in flight on one GPU.
// Supported: 8 bits per sample, colour types 0 (grey), 2 (RGB), 6 (RGBA), non-interlaced - what the reference's
// test images are.  Anything else is refused by b2j_png_info / b2j_png_decode with B2J_E_SHAPE (callers keep PIL).
#include <memory>
#include <new>
#include <vector>

#include "b2j_common.cuh"

namespace b2j {

constexpr int kWin = 32768;                       // DEFLATE window (bytes), a ring in shared memory
constexpr int kFlush = 16384;                     // the ring is written out in halves
constexpr int kFastLL = 10, kFastD = 8;           // first-level table widths (bits)
constexpr int kSoloMatch = 8;                     // matches up to this length are copied by lane 0 alone

enum PngStatus { kPngOk = 0, kPngBadHeader = 1, kPngBadCode = 2, kPngOutOverrun = 3, kPngInOverrun = 4, kPngSize = 5, kPngFilter = 6 };

struct PngParams {
    const uint8_t *zlib;          // [n][zcap]: zlib streams, 16-byte aligned, zero padded
    const int64_t *zbytes;        // [n]
    uint8_t *raw;                 // [n][raw_cap]: filtered scanlines, then reconstructed in place
    int32_t *status;              // [n]
    uint8_t *rgb;                 // [n][height][width][3]
    long long zcap, raw_cap;
    int width, height, channels;
    int n;
};

struct Huff {                     // canonical code (RFC 1951 3.2.2) of up to 288 symbols
    uint16_t count[16];           // codes per length
    uint16_t symbol[288];         // symbols ordered by (length, value)
};

constexpr int kInWords = 1024;                    // the compressed stream is staged by the warp through a 4 KB ring
constexpr int kInMargin = 256;                    // words lane 0 must have in front of it before a step (a dynamic block header
                                                  // reads at most ~75)
struct InflateSmem {
    uint8_t win[kWin];
    uint32_t in[kInWords];
    uint16_t fast_ll[1 << kFastLL];               // symbol << 4 | length; 0 = longer than the table
    uint16_t fast_d[1 << kFastD];
    Huff ll, d;
    uint8_t lengths[320];
};

struct BitReader {
    const uint32_t *ring;         // InflateSmem::in: words [loaded - kInWords, loaded) of the stream (little endian, zero padded)
    long long pos, nwords;        // next word to take; words of the stream (past them: zeros; consumed() tells)
    uint64_t buf;
    int cnt;
    __device__ __forceinline__ void refill()
    {
        if (cnt <= 32) {
            const uint32_t w = pos < nwords ? ring[pos & (kInWords - 1)] : 0u;
            pos++;
            buf |= (uint64_t)w << cnt;
            cnt += 32;
        }
    }
    __device__ __forceinline__ long long consumed() const { return pos * 32 - cnt; }   // bits
    __device__ __forceinline__ uint32_t bits(int n)                 // n <= 16
    {
        refill();
        const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        buf >>= n;
        cnt -= n;
        return v;
    }
};

// Builds the canonical tables from code lengths; returns < 0 for an over-subscribed set, > 0 for an incomplete one.
__device__ __noinline__ int huff_build(Huff &h, const uint8_t *length, int n, uint16_t *fast, int fast_bits)
{
    for (int i = 0; i < 16; i++) h.count[i] = 0;
    for (int i = 0; i < n; i++) h.count[length[i]]++;
    if (fast)
        for (int i = 0; i < (1 << fast_bits); i++) fast[i] = 0;
    if (h.count[0] == n) return 0;                                   // no codes at all (allowed for distances)
    int left = 1;
    for (int len = 1; len < 16; len++) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return left;
    }
    uint16_t offs[16], next[16];
    offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + h.count[len];
    {
        int code = 0;
        for (int len = 1; len < 16; len++) {
            code = (code + h.count[len - 1] * (len > 1)) << 1;
            next[len] = (uint16_t)code;
        }
    }
    for (int s = 0; s < n; s++) {
        const int len = length[s];
        if (!len) continue;
        h.symbol[offs[len]++] = (uint16_t)s;
        const int code = next[len]++;
        if (fast && len <= fast_bits) {                              // the stream delivers a code MSB first: index = reversed code
            const uint32_t rev = __brev((uint32_t)code) >> (32 - len);
            for (uint32_t k = rev; k < (1u << fast_bits); k += 1u << len) fast[k] = (uint16_t)((s << 4) | len);
        }
    }
    return left;
}

// The canonical walk, bit by bit (puff-style): codes longer than the first-level table, and the code-length code.  Out of
// line and rolled: it is rare, and inlined copies of it pushed the literal loop out of the instruction cache.
__device__ __noinline__ int huff_walk(uint32_t bits15, const Huff *h)     // returns symbol << 4 | length, or -1
{
    int code = 0, first = 0, index = 0;
#pragma unroll 1
    for (int len = 1; len <= 15; len++) {
        code |= (int)(bits15 & 1u);
        bits15 >>= 1;
        const int count = h->count[len];
        if (code - count < first) return (h->symbol[index + (code - first)] << 4) | len;
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// One symbol: first-level table (FAST_BITS > 0), else the canonical walk.
template <int FAST_BITS>
__device__ __forceinline__ int huff_decode(BitReader &br, const Huff &h, const uint16_t *fast)
{
    br.refill();
    int e = 0;
    if (FAST_BITS > 0) e = fast[(uint32_t)br.buf & ((1u << (FAST_BITS > 0 ? FAST_BITS : 1)) - 1u)];
    if (e == 0) {
        e = huff_walk((uint32_t)br.buf & 0x7FFFu, &h);
        if (e < 0) return -1;
    }
    br.buf >>= (e & 15);
    br.cnt -= (e & 15);
    return e >> 4;
}

__constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
                                       6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// What lane 0 hands to the warp after every step of the decoder.
enum { kOpLiteral = 0, kOpMatch = 1, kOpEnd = 3, kOpError = 4 };

__global__ void __launch_bounds__(32) k_png_inflate(const PngParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem &S = *reinterpret_cast<InflateSmem *>(smem_raw);
    const int lane = threadIdx.x;
    for (int f = blockIdx.x; f < P.n; f += gridDim.x) {
        uint8_t *out = P.raw + (long long)f * P.raw_cap;
        const long long need = (long long)P.height * (1 + (long long)P.width * P.channels);
        long long o = 0, flushed = 0;                                // bytes produced (lane 0 runs ahead between two hand-overs) / written to `out`
        int status = kPngOk;
        BitReader br;
        const uint32_t *zwords = reinterpret_cast<const uint32_t *>(P.zlib + (long long)f * P.zcap);
        br.ring = S.in;
        br.pos = 0;
        br.nwords = (P.zbytes[f] + 3) / 4;
        br.buf = 0;
        br.cnt = 0;
        long long loaded = 0;                                        // words of the stream staged so far (warp-uniform)
        auto stage_input = [&](long long pos) {                      // whole warp: words [loaded, pos + kInWords - 32) into the ring
            long long upto = pos + kInWords - 32;
            if (upto > br.nwords) upto = br.nwords;
            for (long long w = loaded + lane; w < upto; w += 32) S.in[w & (kInWords - 1)] = __ldg(zwords + w);
            if (upto > loaded) loaded = upto;
            __syncwarp();
        };
        stage_input(0);
        // decoder state of lane 0
        int in_block = 0, last = 0, block_type = 0;
        uint32_t stored_left = 0;
        const long long total_bits = P.zbytes[f] * 8;
        if (lane == 0) {
            const uint32_t cmf = br.bits(8), flg = br.bits(8);        // RFC 1950: deflate, window <= 32 KB, no preset dictionary
            if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) status = kPngBadHeader;
        }
        auto flush_to = [&](long long upto) {                         // whole warp: ring bytes [flushed, upto) -> out
            for (long long i = flushed + lane * 16; i < upto; i += 32 * 16) {
                if (i + 16 <= upto && i + 16 <= P.raw_cap)
                    *reinterpret_cast<uint4 *>(out + i) = *reinterpret_cast<const uint4 *>(S.win + (i & (kWin - 1)));
                else
                    for (long long j = i; j < upto && j < P.raw_cap; j++) out[j] = S.win[j & (kWin - 1)];
            }
            flushed = upto;
        };
        for (;;) {
            int op = kOpError, a = 0, b = 0;                          // literal: a = byte (-1: none); match: a = length, b = distance
            if (lane == 0) {
                // lane 0 decodes on its own until something needs the warp: a match to copy, the ring to flush, the end
                for (;;) {
                    op = kOpError;
                    a = b = 0;
                    if (status != kPngOk) {
                        op = kOpError;
                    } else if (!in_block) {
                        if (last) {
                            op = kOpEnd;
                        } else {
                            last = (int)br.bits(1);
                            block_type = (int)br.bits(2);
                            in_block = 1;
                            op = kOpLiteral;
                            a = -1;                                       // (no output this step)
                            if (block_type == 0) {
                                const int drop = br.cnt & 7;              // to the byte boundary
                                br.buf >>= drop;
                                br.cnt -= drop;
                                const uint32_t len = br.bits(16), nlen = br.bits(16);
                                if ((len ^ 0xFFFFu) != nlen) status = kPngBadCode;
                                stored_left = len;
                                if (len == 0) in_block = 0;
                            } else if (block_type == 1) {
                                for (int i = 0; i < 144; i++) S.lengths[i] = 8;
                                for (int i = 144; i < 256; i++) S.lengths[i] = 9;
                                for (int i = 256; i < 280; i++) S.lengths[i] = 7;
                                for (int i = 280; i < 288; i++) S.lengths[i] = 8;
                                huff_build(S.ll, S.lengths, 288, S.fast_ll, kFastLL);
                                for (int i = 0; i < 30; i++) S.lengths[i] = 5;
                                huff_build(S.d, S.lengths, 30, S.fast_d, kFastD);
                            } else if (block_type == 2) {
                                const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
                                if (nlen > 286 || ndist > 30) status = kPngBadCode;
                                for (int i = 0; i < 19; i++) S.lengths[i] = 0;
                                for (int i = 0; i < ncode; i++) S.lengths[kClOrder[i]] = (uint8_t)br.bits(3);
                                // the code-length code is decoded with the canonical walk only (S.d is free until the tables are built)
                                if (status == kPngOk && huff_build(S.d, S.lengths, 19, nullptr, 0) != 0) status = kPngBadCode;
                                int idx = 0;
                                while (status == kPngOk && idx < nlen + ndist) {
                                    const int sym = huff_decode<0>(br, S.d, nullptr);
                                    if (sym < 0) { status = kPngBadCode; break; }
                                    if (sym < 16) {
                                        S.lengths[idx++] = (uint8_t)sym;
                                    } else {
                                        int rep, val = 0;
                                        if (sym == 16) {
                                            if (idx == 0) { status = kPngBadCode; break; }
                                            val = S.lengths[idx - 1];
                                            rep = 3 + (int)br.bits(2);
                                        } else if (sym == 17) {
                                            rep = 3 + (int)br.bits(3);
                                        } else {
                                            rep = 11 + (int)br.bits(7);
                                        }
                                        if (idx + rep > nlen + ndist) { status = kPngBadCode; break; }
                                        while (rep--) S.lengths[idx++] = (uint8_t)val;
                                    }
                                }
                                if (status == kPngOk) {
                                    if (S.lengths[256] == 0) status = kPngBadCode;          // no end-of-block code
                                    // the distance lengths follow the literal/length ones: build from a copy of their own
                                    uint8_t dl[30];
                                    for (int i = 0; i < ndist; i++) dl[i] = S.lengths[nlen + i];
                                    const int e1 = huff_build(S.ll, S.lengths, nlen, S.fast_ll, kFastLL);
                                    if (e1 < 0 || (e1 > 0 && nlen - S.ll.count[0] != 1)) status = kPngBadCode;
                                    for (int i = 0; i < ndist; i++) S.lengths[i] = dl[i];
                                    const int e2 = huff_build(S.d, S.lengths, ndist, S.fast_d, kFastD);
                                    if (e2 < 0 || (e2 > 0 && ndist - S.d.count[0] != 1)) status = kPngBadCode;
                                }
                            } else {
                                status = kPngBadCode;
                            }
                            if (status != kPngOk) op = kOpError;
                        }
                    } else if (block_type == 0) {
                        // stored bytes (rare in PNG files: compression level 0): one per step through the bit reader
                        op = kOpLiteral;
                        a = (int)br.bits(8);
                        if (--stored_left == 0) in_block = 0;
                    } else {
                        // literals in a row (the bulk of photographic content), kept tight: until a length code, the end of
                        // the block, or the point where the warp has to flush the ring
                        long long room64 = (long long)(kWin - 512) - (o - flushed);
                        if (need + 1 - o < room64) room64 = need + 1 - o;
                        {                                             // (a literal takes at most 15 bits of input)
                            const long long in_room = (loaded - br.pos - 8) * 2;
                            if (loaded < br.nwords && in_room < room64) room64 = in_room;
                        }
                        int room = (int)room64, sym = -2;
                        uint32_t oi = (uint32_t)o & (kWin - 1);
                        while (room > 0) {
                            sym = huff_decode<kFastLL>(br, S.ll, S.fast_ll);
                            if ((unsigned)sym >= 256u) break;
                            S.win[oi] = (uint8_t)sym;
                            oi = (oi + 1u) & (kWin - 1);
                            o++;
                            room--;
                            sym = -2;
                        }
                        if (sym == -2) {                              // no room left: hand over (nothing to output)
                            op = kOpLiteral;
                            a = -1;
                        } else if (sym < 0) {
                            status = kPngBadCode;
                            op = kOpError;
                        } else if (sym < 256) {
                            op = kOpLiteral;
                            a = sym;
                        } else if (sym == 256) {
                            in_block = 0;
                            op = kOpLiteral;
                            a = -1;
                        } else if (sym > 285) {
                            status = kPngBadCode;
                            op = kOpError;
                        } else {
                            const int ls = sym - 257;
                            a = kLenBase[ls] + (int)br.bits(kLenExtra[ls]);
                            const int ds = huff_decode<kFastD>(br, S.d, S.fast_d);
                            if (ds < 0 || ds > 29) {
                                status = kPngBadCode;
                                op = kOpError;
                            } else {
                                b = kDistBase[ds] + (int)br.bits(kDistExtra[ds]);
                                op = kOpMatch;
                                if (a <= kSoloMatch && (long long)b <= o) {
                                    // short matches (most of those a photographic PNG has) cost less copied by lane 0 than
                                    // handed to the warp: sources first when they do not overlap the destination
                                    const uint32_t dst = (uint32_t)o & (kWin - 1), src = (uint32_t)(o - b) & (kWin - 1);
                                    if (b >= a) {
                                        uint8_t v[kSoloMatch];
#pragma unroll
                                        for (int i = 0; i < kSoloMatch; i++)
                                            if (i < a) v[i] = S.win[(src + i) & (kWin - 1)];
#pragma unroll
                                        for (int i = 0; i < kSoloMatch; i++)
                                            if (i < a) S.win[(dst + i) & (kWin - 1)] = v[i];
                                    } else {
                                        for (int i = 0; i < a; i++) S.win[(dst + i) & (kWin - 1)] = S.win[(src + i) & (kWin - 1)];
                                    }
                                    o += a;
                                    op = kOpLiteral;
                                    a = -1;
                                }
                            }
                        }
                    }
                    if (br.consumed() > total_bits) { status = kPngInOverrun; op = kOpError; }
                    if (op != kOpLiteral) break;
                    if (a >= 0) {
                        S.win[o & (kWin - 1)] = (uint8_t)a;
                        o++;
                    }
                    if (o > need || o - flushed >= kWin - 512) break;
                    if (loaded < br.nwords && loaded - br.pos < kInMargin) break;     // the warp stages more input
                }
            }
            op = __shfl_sync(0xffffffffu, op, 0);
            o = __shfl_sync(0xffffffffu, o, 0);
            if (op == kOpEnd || op == kOpError) break;
            stage_input(__shfl_sync(0xffffffffu, br.pos, 0));
            if (op == kOpMatch) {
                a = __shfl_sync(0xffffffffu, a, 0);
                b = __shfl_sync(0xffffffffu, b, 0);
                if ((long long)b > o) {
                    if (lane == 0) status = kPngBadCode;              // distance reaches in front of the stream
                    break;
                }
                __syncwarp();
                if (b >= 32) {                                        // chunks of 32 bytes: every source byte exists before its chunk
                    for (int base = 0; base < a; base += 32) {
                        const int i = base + lane;
                        if (i < a) S.win[(o + i) & (kWin - 1)] = S.win[(o + i - b) & (kWin - 1)];
                        __syncwarp();
                    }
                } else {                                              // overlapping copy: the pattern of `b` bytes repeats
                    uint8_t v[9];
                    int k = 0;
                    for (int i = lane; i < a; i += 32) v[k++] = S.win[(o - b + (i % b)) & (kWin - 1)];
                    __syncwarp();
                    k = 0;
                    for (int i = lane; i < a; i += 32) S.win[(o + i) & (kWin - 1)] = v[k++];
                }
                o += a;
            }
            if (o > need) {
                if (lane == 0) status = kPngOutOverrun;
                break;
            }
            __syncwarp();
            if (o - flushed >= kWin - 512) flush_to(flushed + kFlush);  // keep more than a maximal match of room in the ring
        }
        __syncwarp();
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status == kPngOk) {
            flush_to(o);
            if (o != need) status = kPngSize;
        }
        if (lane == 0) P.status[f] = status;
        __syncwarp();
    }
}

__device__ __forceinline__ int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Undoes the scanline filters and writes the RGB frame.  Lane = row of a group of 32, skewed by one pixel per
// row; a pixel is up to four samples packed into one word (x = left, b = above, c = above left).
__global__ void __launch_bounds__(32) k_png_unfilter(const PngParams P)
{
    const int lane = threadIdx.x;
    const int ch = P.channels, W = P.width, H = P.height;
    const long long stride = 1 + (long long)W * ch;
    for (int f = blockIdx.x; f < P.n; f += gridDim.x) {
        if (P.status[f] != kPngOk) continue;
        uint8_t *raw = P.raw + (long long)f * P.raw_cap;
        uint8_t *rgb = P.rgb + (long long)f * H * W * 3;
        bool bad = false;
        for (int r0 = 0; r0 < H; r0 += 32) {
            const int r = r0 + lane;
            const bool row_ok = r < H;
            uint8_t *line = raw + (long long)(row_ok ? r : 0) * stride;
            const int ft = row_ok ? line[0] : 0;
            if (ft > 4) bad = true;
            const uint8_t *above = raw + (long long)(r0 - 1) * stride + 1;     // lane 0: the last row of the previous group
            uint32_t left = 0u, up_left = 0u, mine_prev = 0u;
            // the filtered samples of a pixel (and, for lane 0, the reconstructed pixel above it) are fetched two steps ahead:
            // bytes right of the current pixel are still untouched, so the loads do not wait for this step's stores
            auto fetch = [&](const uint8_t *base, int x, bool ok) {
                uint32_t v = 0u;
                if (ok && x >= 0 && x < W)
                    for (int c = 0; c < ch; c++) v |= (uint32_t)base[(long long)x * ch + c] << (8 * c);
                return v;
            };
            const bool top = lane == 0 && r0 > 0;
            uint32_t f0 = fetch(line + 1, -lane, row_ok), f1 = fetch(line + 1, 1 - lane, row_ok);
            uint32_t a0 = fetch(above, 0, top), a1 = fetch(above, 1, top);
            for (int t = 0; t < W + 31; t++) {
                // the pixel above: what the lane of the row above produced in the previous step
                uint32_t up = __shfl_up_sync(0xffffffffu, mine_prev, 1);
                const int x = t - lane;
                const bool on = row_ok && x >= 0 && x < W;
                const uint32_t filt = f0;
                f0 = f1;
                f1 = fetch(line + 1, x + 2, row_ok);
                if (lane == 0) {
                    up = a0;
                    a0 = a1;
                    a1 = fetch(above, x + 2, top);
                }
                if (r == 0) up = 0u;
                uint32_t cur = 0u;
                if (on) {
                    uint8_t *px = line + 1 + (long long)x * ch;
                    for (int c = 0; c < ch; c++) {
                        const int a = (left >> (8 * c)) & 255, b = (up >> (8 * c)) & 255, cc = (up_left >> (8 * c)) & 255;
                        int pred = 0;
                        if (ft == 1) pred = a;
                        else if (ft == 2) pred = b;
                        else if (ft == 3) pred = (a + b) >> 1;
                        else if (ft == 4) pred = paeth(a, b, cc);
                        const int v = (int)(((filt >> (8 * c)) & 255u) + (uint32_t)pred) & 255;
                        cur |= (uint32_t)v << (8 * c);
                    }
                    if (lane == 31)                                       // reconstructed in place: the next group reads its row above
                        for (int c = 0; c < ch; c++) px[c] = (uint8_t)(cur >> (8 * c));
                    uint8_t *o = rgb + ((long long)r * W + x) * 3;
                    if (ch == 1) {
                        o[0] = o[1] = o[2] = (uint8_t)cur;
                    } else {
                        o[0] = (uint8_t)cur;
                        o[1] = (uint8_t)(cur >> 8);
                        o[2] = (uint8_t)(cur >> 16);
                    }
                    left = cur;
                    up_left = up;
                }
                mine_prev = cur;
            }
            __syncwarp();                                                 // the group's last row is complete before the next group reads it
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) P.status[f] = kPngFilter;
    }
}

// ---------------------------------------------------------------- host: chunk walk
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngFile {
    int width, height, channels;
    int64_t zlib_bytes;
};

// Walks the chunks; gather != nullptr collects the IDAT payloads (gather must hold zlib_bytes).
static int png_walk(const uint8_t *png, int64_t n, PngFile *info, uint8_t *gather)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    B2J_REQUIRE(png && n >= 8 + 25 && memcmp(png, sig, 8) == 0, B2J_E_ARG, "b2j_png: not a PNG file");
    int64_t pos = 8, z = 0;
    bool have_ihdr = false, have_iend = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(png + pos);
        const uint8_t *type = png + pos + 4, *data = png + pos + 8;
        B2J_REQUIRE(pos + 12 + (int64_t)len <= n, B2J_E_ARG, "b2j_png: truncated chunk");
        if (!memcmp(type, "IHDR", 4)) {
            B2J_REQUIRE(len == 13, B2J_E_ARG, "b2j_png: bad IHDR");
            info->width = (int)be32(data);
            info->height = (int)be32(data + 4);
            const int depth = data[8], ctype = data[9], interlace = data[12];
            B2J_REQUIRE(info->width >= 1 && info->width <= 65535 && info->height >= 1 && info->height <= 65535, B2J_E_SHAPE,
                        "b2j_png: image side out of range 1..65535");
            B2J_REQUIRE(depth == 8 && (ctype == 0 || ctype == 2 || ctype == 6) && interlace == 0 && data[10] == 0 && data[11] == 0,
                        B2J_E_SHAPE, "b2j_png: only 8-bit grey / RGB / RGBA, non-interlaced files are decoded on the device "
                        "(bit depth %d, colour type %d, interlace %d)", depth, ctype, interlace);
            info->channels = ctype == 0 ? 1 : (ctype == 2 ? 3 : 4);
            have_ihdr = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            B2J_REQUIRE(have_ihdr, B2J_E_ARG, "b2j_png: IDAT before IHDR");
            if (gather) memcpy(gather + z, data, len);
            z += len;
        } else if (!memcmp(type, "IEND", 4)) {
            have_iend = true;
            break;
        }
        pos += 12 + (int64_t)len;
    }
    B2J_REQUIRE(have_ihdr && have_iend && z >= 6, B2J_E_ARG, "b2j_png: missing IHDR, IDAT or IEND");
    info->zlib_bytes = z;
    return B2J_OK;
}

struct PngWs { size_t zlib, zbytes, raw, status, total; long long zcap, raw_cap; };

static PngWs png_ws(int width, int height, int channels, int n, int64_t max_zlib)
{
    PngWs w;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.zcap = ((long long)max_zlib + 16 + 15) & ~15LL;
    w.raw_cap = ((long long)height * (1 + (long long)width * channels) + 15) & ~15LL;
    w.zlib = 0;
    w.zbytes = up((size_t)n * w.zcap);
    w.raw = up(w.zbytes + (size_t)n * sizeof(int64_t));
    w.status = up(w.raw + (size_t)n * w.raw_cap);
    w.total = up(w.status + (size_t)n * sizeof(int32_t));
    return w;
}

}  // namespace b2j

using namespace b2j;

extern "C" int b2j_png_info(const uint8_t *png_host, int64_t png_bytes, int32_t *width, int32_t *height, int32_t *channels,
                            int64_t *zlib_bytes)
{
    B2J_REQUIRE(width && height && channels && zlib_bytes, B2J_E_ARG, "b2j_png_info: null pointer");
    PngFile info;
    const int rc = png_walk(png_host, png_bytes, &info, nullptr);
    if (rc) return rc;
    *width = info.width;
    *height = info.height;
    *channels = info.channels;
    *zlib_bytes = info.zlib_bytes;
    return B2J_OK;
}

extern "C" size_t b2j_png_workspace_bytes(int32_t width, int32_t height, int32_t channels, int32_t n_files, int64_t max_zlib_bytes)
{
    if (width < 1 || width > 65535 || height < 1 || height > 65535 || (channels != 1 && channels != 3 && channels != 4) || n_files < 1 ||
        max_zlib_bytes < 0)
        return 0;
    return png_ws(width, height, channels, n_files, max_zlib_bytes).total;
}

extern "C" int b2j_png_decode(const uint8_t *const *png_host, const int64_t *png_bytes, int32_t n_files, int32_t width, int32_t height,
                              int32_t channels, void *workspace, size_t workspace_bytes, uint8_t *rgb, int32_t *status_host, void *stream)
{
    B2J_REQUIRE(png_host && png_bytes && workspace && rgb && n_files >= 1, B2J_E_ARG, "b2j_png_decode: null pointer or no files");
    B2J_REQUIRE(((uintptr_t)workspace & 255) == 0, B2J_E_ALIGN, "b2j_png_decode: workspace must be 256-byte aligned");
    std::vector<PngFile> info((size_t)n_files);
    int64_t max_z = 0;
    for (int i = 0; i < n_files; i++) {
        const int rc = png_walk(png_host[i], png_bytes[i], &info[(size_t)i], nullptr);
        if (rc) return rc;
        B2J_REQUIRE(info[(size_t)i].width == width && info[(size_t)i].height == height && info[(size_t)i].channels == channels, B2J_E_SHAPE,
                    "b2j_png_decode: file %d is %d x %d x %d, the batch is %d x %d x %d", i, info[(size_t)i].width, info[(size_t)i].height,
                    info[(size_t)i].channels, width, height, channels);
        if (info[(size_t)i].zlib_bytes > max_z) max_z = info[(size_t)i].zlib_bytes;
    }
    const PngWs W = png_ws(width, height, channels, n_files, max_z);
    B2J_REQUIRE(workspace_bytes >= W.total, B2J_E_WORKSPACE, "b2j_png_decode: workspace of %zu bytes, %zu needed", workspace_bytes, W.total);
    // gather the zlib streams (zero padded) and their lengths: two copies to the device
    const size_t staging_bytes = (size_t)n_files * (size_t)W.zcap;
    std::unique_ptr<uint8_t[]> staging(new (std::nothrow) uint8_t[staging_bytes]);      // (not zero-filled: only the tails are)
    B2J_REQUIRE(staging != nullptr, B2J_E_WORKSPACE, "b2j_png_decode: no host memory for %zu bytes of zlib streams", staging_bytes);
    std::vector<int64_t> zb((size_t)n_files);
    for (int i = 0; i < n_files; i++) {
        PngFile tmp;
        uint8_t *dst = staging.get() + (size_t)i * (size_t)W.zcap;
        const int rc = png_walk(png_host[i], png_bytes[i], &tmp, dst);
        if (rc) return rc;
        zb[(size_t)i] = tmp.zlib_bytes;
        memset(dst + tmp.zlib_bytes, 0, (size_t)(W.zcap - tmp.zlib_bytes));
    }
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B2J_CUDA(cudaMemcpyAsync(ws + W.zlib, staging.get(), staging_bytes, cudaMemcpyHostToDevice, st));
    B2J_CUDA(cudaMemcpyAsync(ws + W.zbytes, zb.data(), zb.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    B2J_CUDA(cudaStreamSynchronize(st));                              // the staging vectors go out of scope at return
    PngParams P;
    P.zlib = ws + W.zlib;
    P.zbytes = reinterpret_cast<const int64_t *>(ws + W.zbytes);
    P.raw = ws + W.raw;
    P.status = reinterpret_cast<int32_t *>(ws + W.status);
    P.rgb = rgb;
    P.zcap = W.zcap;
    P.raw_cap = W.raw_cap;
    P.width = width;
    P.height = height;
    P.channels = channels;
    P.n = n_files;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    B2J_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        B2J_CUDA(cudaFuncSetAttribute(k_png_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InflateSmem)));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    k_png_inflate<<<n_files, 32, sizeof(InflateSmem), st>>>(P);
    B2J_LAUNCH_CHECK("k_png_inflate");
    k_png_unfilter<<<n_files, 32, 0, st>>>(P);
    B2J_LAUNCH_CHECK("k_png_unfilter");
    if (status_host) {
        B2J_CUDA(cudaMemcpyAsync(status_host, P.status, (size_t)n_files * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        B2J_CUDA(cudaStreamSynchronize(st));
    }
    return B2J_OK;
}
