// K2..K4: run-length + Huffman symbol sizes per block, per-frame exclusive scan
// of bit offsets, parallel bit packing, 0xFF byte stuffing, JFIF framing.
//
// Symbol rules (the only JPEG-style RLE in the reference is gateware):
//   DC: diff against the previous block of the same component, predictor 0 at
//       the start of a frame, never reset (test/model/enc_frame.py:202,219-230;
//       litejpeg/core/rle/rlecore.py:86-98);
//   AC: (run, amplitude) pairs; B2J_ZRL_EAGER emits (15,0) on every 16th
//       consecutive zero and (0,0) when index 63 is a zero that did not just
//       complete a ZRL (rlecore.py:101-114); B2J_ZRL_STANDARD is T.81 F.1.2.2;
//   size = bit length of |x| (litejpeg/core/rle/entropycoder.py:55-65).
// Everything after the symbols is [ext] (the reference's huffman() is an
// identity stub, enc_frame.py:179-180): Annex-K codes, additional bits
// (x > 0 ? x : x - 1), MSB-first packing, 1-padding, 0xFF -> 0xFF00, header, EOI.
//
// The same block walker produces the sizes (K2) and the bits (K4), so scanned
// offsets and emitted lengths cannot disagree.
#include "b2j_entropy.cuh"

namespace b2j {

// ---------------------------------------------------------------- K2
template <bool EAGER>
__global__ void __launch_bounds__(128) k_block_bits(const int16_t *__restrict__ coefs, int32_t *__restrict__ bits, EGeo e,
                                                    const __grid_constant__ LutParam L)
{
    __shared__ uint32_t s_dc[32], s_ac[512];
    load_lut(L, s_dc, s_ac);
    __syncthreads();
    const long long total = e.nblk * e.n_frames;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const long long f = g / e.nblk, b = g % e.nblk;
    const int16_t *fc = coefs + f * e.nblk * 64;
    const int16_t *c = fc + b * 64;
    const int chroma = (int)(b % e.bpm) >= e.hv;
    CountSink sink{0};
    walk_block<EAGER>(c, (int)c[0] - prev_dc(fc, b, e), s_dc + 16 * chroma, s_ac + 256 * chroma, sink);
    bits[g] = sink.bits;
}

// ---------------------------------------------------------------- K3 (three small kernels)
constexpr int kScanChunk = 1024;          // blocks per scan chunk (256 threads x 4)

__global__ void __launch_bounds__(256) k_chunk_sums(const int32_t *__restrict__ bits, uint32_t *__restrict__ partial,
                                                    long long nblk, int chunks)
{
    __shared__ uint32_t s_warp[8];
    const int f = blockIdx.y, ch = blockIdx.x;
    const int32_t *p = bits + (long long)f * nblk;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
        long long b = (long long)ch * kScanChunk + i * 256 + threadIdx.x;
        if (b < nblk) v += (uint32_t)p[b];
    }
    uint32_t tot;
    cta_excl_scan_256(v, s_warp, &tot);
    if (threadIdx.x == 0) partial[(long long)f * chunks + ch] = tot;
}

// one CTA per frame: partial[] becomes the exclusive prefix; total_bits[f] the frame sum
__global__ void __launch_bounds__(256) k_scan_partials(uint32_t *__restrict__ partial, int64_t *__restrict__ total_bits, int chunks)
{
    __shared__ uint32_t s_warp[8];
    uint32_t *p = partial + (long long)blockIdx.x * chunks;
    unsigned long long carry = 0;
    for (int base = 0; base < chunks; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < chunks ? p[i] : 0u;
        uint32_t tot;
        const uint32_t ex = cta_excl_scan_256(v, s_warp, &tot);
        if (i < chunks) p[i] = (uint32_t)carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) total_bits[blockIdx.x] = (int64_t)carry;
}

__global__ void __launch_bounds__(256) k_scan_apply(const int32_t *__restrict__ bits, const uint32_t *__restrict__ partial,
                                                    uint32_t *__restrict__ offsets, long long nblk, int chunks)
{
    __shared__ uint32_t s_warp[8];
    const int f = blockIdx.y, ch = blockIdx.x;
    const int32_t *p = bits + (long long)f * nblk;
    uint32_t *o = offsets + (long long)f * nblk;
    // thread t owns 4 consecutive blocks so that the scan order is the block order
    const long long b0 = (long long)ch * kScanChunk + threadIdx.x * 4;
    uint32_t v[4], sum = 0;
    for (int i = 0; i < 4; i++) { v[i] = b0 + i < nblk ? (uint32_t)p[b0 + i] : 0u; sum += v[i]; }
    uint32_t tot;
    uint32_t ex = cta_excl_scan_256(sum, s_warp, &tot) + partial[(long long)f * chunks + ch];
    for (int i = 0; i < 4; i++) {
        if (b0 + i < nblk) o[b0 + i] = ex;
        ex += v[i];
    }
}

// ---------------------------------------------------------------- K4a: bit packing (unstuffed)
constexpr int kPackBlocks = 128;                    // blocks per CTA
constexpr int kLook = 8;                            // lookahead blocks (>= 31 bits / 4 bits per smallest block)
constexpr int kStageWords = ((kPackBlocks + kLook) * 1728 + 31) / 32 + 4;

struct PackSink {
    uint32_t *stage;              // shared staging words, zeroed
    unsigned long long acc;
    int nbits;                    // valid low bits of acc
    int widx;                     // next staging word
    __device__ __forceinline__ void begin(uint32_t *s, uint32_t bitpos)
    {
        stage = s; acc = 0ull; nbits = (int)(bitpos & 31u); widx = (int)(bitpos >> 5);
    }
    __device__ __forceinline__ void put(uint32_t val, int len)
    {
        acc = (acc << len) | val;
        nbits += len;
        if (nbits >= 32) {
            atomicOr(stage + widx, (uint32_t)(acc >> (nbits - 32)));
            widx++;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nbits > 0) atomicOr(stage + widx, (uint32_t)(acc << (32 - nbits)));
    }
};

// grid: (chunks_per_frame, n_frames).  unstuffed: n_frames slots of ucap bytes
// (ucap multiple of 4), big-endian bit order, padded with 1s to a whole byte.
template <bool EAGER>
__global__ void __launch_bounds__(kPackBlocks) k_bitpack(const int16_t *__restrict__ coefs, const uint32_t *__restrict__ offsets,
                                                         const int64_t *__restrict__ total_bits, uint8_t *__restrict__ unstuffed,
                                                         long long ucap, EGeo e, const __grid_constant__ LutParam L)
{
    __shared__ uint32_t s_dc[32], s_ac[512];
    __shared__ uint32_t s_stage[kStageWords];
    load_lut(L, s_dc, s_ac);
    const int f = blockIdx.y;
    const long long b0 = (long long)blockIdx.x * kPackBlocks;
    const long long b1 = min(b0 + kPackBlocks, e.nblk);
    const int16_t *fc = coefs + (long long)f * e.nblk * 64;
    const uint32_t *off = offsets + (long long)f * e.nblk;
    const uint32_t tbits = (uint32_t)total_bits[f];
    const uint32_t S = off[b0];
    const uint32_t E = b1 < e.nblk ? off[b1] : tbits;
    const bool last = b1 >= e.nblk;
    const uint32_t wbase = S >> 5;                                   // first staging word
    const uint32_t W0 = (S + 31u) >> 5;                              // first owned word
    // padded end of the frame's bit stream (whole bytes)
    const uint32_t Epad = last ? ((E + 7u) & ~7u) : E;
    const uint32_t W1 = (Epad + 31u) >> 5;                           // one past the last owned word
    const int nstage = (int)(W1 - wbase) + 1;
    for (int i = threadIdx.x; i < nstage && i < kStageWords; i += kPackBlocks) s_stage[i] = 0u;
    __syncthreads();

    // own blocks, then the few following blocks that reach into the last owned word
    for (int pass = 0; pass < 2; pass++) {
        const long long b = pass == 0 ? b0 + threadIdx.x : b1 + threadIdx.x;
        bool active = pass == 0 ? b < b1 : (threadIdx.x < kLook && b < e.nblk);
        if (active && pass == 1) active = off[b] < (W1 << 5);
        if (active) {
            const int chroma = (int)(b % e.bpm) >= e.hv;
            const int16_t *c = fc + b * 64;
            PackSink sink;
            sink.begin(s_stage, off[b] - (wbase << 5));
            if (EAGER) walk_block<true>(c, (int)c[0] - prev_dc(fc, b, e), s_dc + 16 * chroma, s_ac + 256 * chroma, sink);
            else walk_block<false>(c, (int)c[0] - prev_dc(fc, b, e), s_dc + 16 * chroma, s_ac + 256 * chroma, sink);
            sink.finish();
        }
    }
    if (last && threadIdx.x == 0 && Epad > E) {                      // pad the final byte with 1s
        const uint32_t pos = E - (wbase << 5), n = Epad - E;
        atomicOr(s_stage + (pos >> 5), ((1u << n) - 1u) << (32u - (pos & 31u) - n));
    }
    __syncthreads();
    uint32_t *dst = reinterpret_cast<uint32_t *>(unstuffed + (long long)f * ucap);
    const uint32_t wcap = (uint32_t)(ucap >> 2);
    for (uint32_t w = W0 + threadIdx.x; w < W1; w += kPackBlocks)
        if (w < wcap) dst[w] = __byte_perm(s_stage[w - wbase], 0u, 0x0123);     // big-endian bytes
}

// ---------------------------------------------------------------- K4b: 0xFF stuffing + framing

__device__ __forceinline__ uint32_t ff_mask16(const uint4 &q, int nvalid)
{
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (i < nvalid && ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) == 0xFFu) m |= 1u << i;
    return m;
}

__global__ void __launch_bounds__(256) k_ff_count(const uint8_t *__restrict__ unstuffed, const int64_t *__restrict__ total_bits,
                                                  uint32_t *__restrict__ ffpart, long long ucap, int chunks)
{
    __shared__ uint32_t s_warp[8];
    const int f = blockIdx.y, ch = blockIdx.x;
    const long long nbytes = min((long long)((total_bits[f] + 7) >> 3), ucap);
    if ((long long)ch * kStuffChunk >= nbytes) {                    // chunk beyond the stream: nothing to count
        if (threadIdx.x == 0) ffpart[(long long)f * chunks + ch] = 0u;
        return;
    }
    const long long pos = (long long)ch * kStuffChunk + threadIdx.x * 16;
    uint32_t cnt = 0;
    if (pos < nbytes) {
        const uint4 q = *reinterpret_cast<const uint4 *>(unstuffed + (long long)f * ucap + pos);
        cnt = __popc(ff_mask16(q, (int)min(16LL, nbytes - pos)));
    }
    uint32_t tot;
    cta_excl_scan_256(cnt, s_warp, &tot);
    if (threadIdx.x == 0) ffpart[(long long)f * chunks + ch] = tot;
}

// One CTA per frame.  A frame is nseg independently packed streams ("segments": 1 without restart
// intervals); ffpart, total_bits and segbase are indexed by frame * nseg + segment.
//   1. per segment (one warp each): exclusive scan of the 0xFF counts of its 4096-byte chunks, in
//      place; segment size = bytes + 0xFF count + 2 (the RSTm / EOI marker that follows it);
//   2. exclusive scan of the segment sizes -> segbase; final file size / capacity verdict.
__global__ void __launch_bounds__(256) k_ff_scan(uint32_t *__restrict__ ffpart, const int64_t *__restrict__ total_bits,
                                                 uint32_t *__restrict__ segbase, int64_t *__restrict__ sizes, long long ucap,
                                                 long long out_cap, int chunks, int nseg, int header_bytes,
                                                 const int *__restrict__ error_flag)
{
    __shared__ uint32_t s_warp[8];
    __shared__ int s_over;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long f = blockIdx.x;
    uint32_t *fp = ffpart + f * nseg * chunks;
    const int64_t *tb = total_bits + f * nseg;
    uint32_t *sb = segbase + f * nseg;
    if (threadIdx.x == 0) s_over = 0;
    __syncthreads();
    if (nseg == 1) {                                                // one long segment: the whole CTA scans its chunks
        unsigned long long c0 = 0;
        for (int base = 0; base < chunks; base += 256) {
            const int i = base + threadIdx.x;
            const uint32_t v = i < chunks ? fp[i] : 0u;
            uint32_t tot;
            const uint32_t ex = cta_excl_scan_256(v, s_warp, &tot);
            if (i < chunks) fp[i] = (uint32_t)c0 + ex;
            c0 += tot;
        }
        if (threadIdx.x == 0) {
            const long long nb = (tb[0] + 7) >> 3;
            if (nb > ucap) s_over = 1;
            sb[0] = (uint32_t)c0 + (uint32_t)nb + 2u;
        }
    } else {
        for (int k = warp; k < nseg; k += 8) {                      // many short segments: one warp each
            uint32_t *p = fp + (long long)k * chunks;
            uint32_t carry = 0;
            for (int base = 0; base < chunks; base += 32) {
                const int i = base + lane;
                const uint32_t v = i < chunks ? p[i] : 0u;
                const uint32_t inc = warp_incl_scan(v);
                if (i < chunks) p[i] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) {
                const long long nb = (tb[k] + 7) >> 3;
                if (nb > ucap) s_over = 1;
                sb[k] = carry + (uint32_t)nb + 2u;
            }
        }
    }
    __syncthreads();
    unsigned long long carry = 0;
    for (int base = 0; base < nseg; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nseg ? sb[i] : 0u;
        uint32_t tot;
        const uint32_t ex = cta_excl_scan_256(v, s_warp, &tot);
        if (i < nseg) sb[i] = (uint32_t)carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if (error_flag && *error_flag) sizes[f] = (int64_t)0x8000000000000000LL;            // internal error
        else if (s_over) sizes[f] = -(header_bytes + 2 * (long long)carry);                 // upper bound of the need
        else {
            const long long need = header_bytes + (long long)carry;
            sizes[f] = need <= out_cap ? need : -need;
        }
    }
}


// grid (segments of all frames, y): the CTAs of a segment stride over the chunks its stream really
// has (the slot capacity is several times the typical stream: empty CTAs cost more than the copy).
// The first segment's first CTA writes the header; the CTA that writes a segment's last chunk
// appends the marker that follows it: RSTm between segments, EOI after the last one.
// OWN_SCAN (frames of one segment, i.e. no restart intervals): the CTA scans the 0xFF counts of its frame's chunks itself
// (a few hundred values) and derives the file size and the capacity verdict - no k_ff_scan launch in front; ffpart then
// holds the raw counts and segbase is not used.
#ifndef B2J_STUFF_MIN_CTAS
#define B2J_STUFF_MIN_CTAS 6
#endif
template <bool OWN_SCAN>
__global__ void __launch_bounds__(256, B2J_STUFF_MIN_CTAS) k_stuff_write(const uint8_t *__restrict__ unstuffed, const int64_t *__restrict__ total_bits,
                                                     const uint32_t *__restrict__ ffpart, const uint32_t *__restrict__ segbase,
                                                     int64_t *__restrict__ sizes, uint8_t *__restrict__ out, long long ucap,
                                                     long long out_cap, int chunks, int nseg, int rst_phase,
                                                     const int *__restrict__ error_flag, const __grid_constant__ HeaderParam H)
{
    extern __shared__ uint32_t s_prefix[];                          // OWN_SCAN: exclusive prefix of the frame's 0xFF counts
    __shared__ uint32_t s_warp[8];
    __shared__ __align__(16) uint8_t s_out[2 * kStuffChunk + 64];
    __shared__ __align__(16) uint4 s_in[256];                        // the chunk's pieces, for the byte-wise pass
    __shared__ uint32_t s_head[257];                                // first three stuffed bytes of every piece
    __shared__ uint16_t s_o[256], s_list[256];                      // staging offset of every piece; pieces of the byte-wise pass
    __shared__ uint32_t s_nlist;
    __shared__ uint32_t s_warp2[2][8];
    const long long vf = blockIdx.x;
    const long long f = vf / nseg;
    const int seg = (int)(vf % nseg);
    pdl_wait();                                                     // the packed streams and their sizes are complete
    const long long nbytes = (total_bits[vf] + 7) >> 3;
    if (OWN_SCAN) {
        const int nused = (int)min((long long)chunks, (nbytes + kStuffChunk - 1) / kStuffChunk);
        unsigned long long c0 = 0;
        for (int base = 0; base < nused; base += 256) {
            const int i = base + threadIdx.x;
            const uint32_t v = i < nused ? ffpart[vf * chunks + i] : 0u;
            uint32_t tot;
            const uint32_t ex = cta_excl_scan_256(v, s_warp, &tot);
            if (i < nused) s_prefix[i] = (uint32_t)c0 + ex;
            c0 += tot;
        }
        __syncthreads();
        // the verdict of k_ff_scan, derived by every CTA of the frame alike
        long long sz;
        if (error_flag && *error_flag) sz = (long long)0x8000000000000000LL;                    // internal error
        else if (nbytes > ucap) sz = -((long long)H.len + 2 * ((long long)c0 + nbytes + 2));    // upper bound of the need
        else {
            const long long need = (long long)H.len + (long long)c0 + nbytes + 2;
            sz = need <= out_cap ? need : -need;
        }
        if (blockIdx.y == 0 && threadIdx.x == 0) sizes[f] = sz;
        if (sz < 0) return;                                         // slot too small: leave it untouched
    } else if (sizes[f] < 0) {
        return;
    }
    uint8_t *dst_frame = out + f * out_cap;
    if (seg == 0 && blockIdx.y == 0)
        for (int i = threadIdx.x; i < H.len; i += 256) dst_frame[i] = H.bytes[i];
    uint8_t *dst_seg = dst_frame + H.len + (OWN_SCAN ? 0u : segbase[vf]);
    // A chunk is 256 pieces of 16 bytes, one per thread.  Staging offset of piece t = (dst & 15) + 16 t + (0xFF bytes in
    // front of it): shared and global memory then agree modulo 16 and the chunk leaves as 16-byte vectors.  Every aligned
    // word of the staging area has ONE writer, the piece in which its first byte lies (that piece appends the first
    // bytes of its successor's stuffed stream, s_head, to fill its last word).  Pieces without a 0xFF (94 %) store four
    // words; the others are listed and written byte by byte by the first threads of the CTA in a second pass, so the
    // byte loop runs in one warp per chunk instead of in every warp that holds such a piece (86 % of them).
    const int t = threadIdx.x;
    auto fetch = [&](int ch_) {                                      // (a chunk's piece is loaded one iteration ahead)
        const long long pos_ = (long long)ch_ * kStuffChunk + t * 16;
        return pos_ < nbytes ? __ldg(reinterpret_cast<const uint4 *>(unstuffed + vf * ucap + pos_)) : make_uint4(0, 0, 0, 0);
    };
    uint4 q_next = fetch(blockIdx.y);
    int parity = 0;
    for (int ch = blockIdx.y; (long long)ch * kStuffChunk < nbytes; ch += gridDim.y, parity ^= 1) {
        const long long pos = (long long)ch * kStuffChunk + t * 16;
        uint4 q = q_next;
        q_next = fetch(ch + (int)gridDim.y);
        if (pos < nbytes) {
            const int nvalid = (int)min(16LL, nbytes - pos);
            if (nvalid < 16) {                                      // the stream's last piece: bytes past its end are not data
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int keep = min(max(nvalid - 4 * i, 0), 4);
                    w[i] = keep == 4 ? w[i] : (keep == 0 ? 0u : w[i] & ((1u << (8 * keep)) - 1u));
                }
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        // 0x80 in every 0xFF byte
        auto ff_bits = [](uint32_t w) { return ~((~w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) & w & 0x80808080u; };
        const uint32_t f0 = ff_bits(q.x), f1 = ff_bits(q.y), f2 = ff_bits(q.z), f3 = ff_bits(q.w);
        const uint32_t cnt = (uint32_t)(__popc(f0) + __popc(f1) + __popc(f2) + __popc(f3));
        // first three bytes of this piece's stuffed stream: [FF 00 b1], [b0 FF 00] or [b0 b1 b2]
        s_head[t] = (f0 & 0x80u) ? (0xFFu | ((q.x & 0xFF00u) << 8)) : ((f0 & 0x8000u) ? (q.x & 0xFFFFu) : (q.x & 0xFFFFFFu));
        if (t == 0) { s_head[256] = 0u; s_nlist = 0u; }
        // exclusive scan over the CTA; the warp totals alternate between two rows, so one barrier is enough
        // (s_head and s_nlist are visible after it)
        uint32_t tot = 0u, ex;
        {
            const uint32_t inc = warp_incl_scan(cnt);
            uint32_t *row = s_warp2[parity];
            if ((t & 31) == 31) row[t >> 5] = inc;
            __syncthreads();
            uint32_t base = 0u;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t x = row[w];
                if (w < (t >> 5)) base += x;
                tot += x;
            }
            ex = base + inc - cnt;
        }
        const long long chunk_in = min((long long)kStuffChunk, nbytes - (long long)ch * kStuffChunk);
        const int n_out = (int)chunk_in + (int)tot;
        uint8_t *dst = dst_seg + (long long)ch * kStuffChunk + (OWN_SCAN ? s_prefix[ch] : ffpart[vf * chunks + ch]);
        const uint32_t shift = (uint32_t)((uintptr_t)dst & 15u);
        const uint32_t o = shift + (uint32_t)t * 16u + ex;
        if (cnt == 0u && t != 0) {
            const uint32_t skip = (0u - o) & 3u, sh = skip * 8u;     // bytes in front of this piece's first word: its predecessor's
            const uint32_t nx = s_head[t + 1];
            uint32_t *w = reinterpret_cast<uint32_t *>(s_out + o + skip);
            w[0] = __funnelshift_r(q.x, q.y, sh);
            w[1] = __funnelshift_r(q.y, q.z, sh);
            w[2] = __funnelshift_r(q.z, q.w, sh);
            w[3] = __funnelshift_r(q.w, nx, sh);
        } else {                                                    // (piece 0 has no predecessor to write its first bytes)
            s_in[t] = q;
            s_o[t] = (uint16_t)o;
            s_list[atomicAdd(&s_nlist, 1u)] = (uint16_t)t;
        }
        __syncthreads();
        if (t < (int)s_nlist) {
            const int p = (int)s_list[t];
            const uint4 qq = s_in[p];
            const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
            uint32_t oo = s_o[p];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint8_t b = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
                s_out[oo++] = b;
                if (b == 0xFF) s_out[oo++] = 0;
            }
            uint32_t nx = s_head[p + 1];
            while (oo & 3u) {                                       // fill the last word with the successor's first bytes
                s_out[oo++] = (uint8_t)nx;
                nx >>= 8;
            }
        }
        __syncthreads();
        // out: head bytes up to the first 16-byte boundary, whole vectors, tail bytes
        {
            const uint32_t end = shift + (uint32_t)n_out;           // staged bytes [shift, end) <-> dst[0, n_out)
            const uint32_t v0 = (shift + 15u) >> 4, v1 = end >> 4;
            uint8_t *gbase = dst - shift;
            for (uint32_t v = v0 + (uint32_t)t; v < v1; v += 256u)
                *reinterpret_cast<uint4 *>(gbase + 16u * v) = *reinterpret_cast<const uint4 *>(s_out + 16u * v);
            if (t < 16) {
                const uint32_t head_end = min(16u * v0, end);
                if (shift + (uint32_t)t < head_end) gbase[shift + t] = s_out[shift + t];
                if (v1 >= v0) {
                    const uint32_t i = 16u * v1 + (uint32_t)t;
                    if (i < end) gbase[i] = s_out[i];
                }
            }
        }
        // the marker after the segment's last chunk: RSTm (T.81 B.2.4.4 / E.1.4), or EOI after the last segment
        if ((long long)(ch + 1) * kStuffChunk >= nbytes && threadIdx.x == 0) {
            dst[n_out] = 0xFF;
            dst[n_out + 1] = seg == nseg - 1 ? 0xD9 : (uint8_t)(0xD0 + ((seg + rst_phase) & 7));
        }
        __syncthreads();                                            // s_out and s_warp are reused by the next chunk
    }
}

static int scan_chunks(long long nblk) { return (int)((nblk + kScanChunk - 1) / kScanChunk); }

int launch_stuffing(const uint8_t *unstuffed, const int64_t *total_bits, uint32_t *ffpart, int64_t *sizes, uint8_t *out,
                    long long ucap, long long out_cap, int n_frames, const HeaderParam &H, const int *error_flag,
                    bool counted, cudaStream_t st, int nseg, int rst_phase, bool pdl)
{
    const int schunks = stuff_chunks(ucap);
    const long long nvf = (long long)n_frames * nseg;
    uint32_t *segbase = ffpart + nvf * schunks;                     // [n_frames * nseg], after the chunk counts
    if (!counted) {
        dim3 sgrid((unsigned)schunks, (unsigned)n_frames);
        k_ff_count<<<sgrid, 256, 0, st>>>(unstuffed, total_bits, ffpart, ucap, schunks);
        B2J_LAUNCH_CHECK("k_ff_count");
    }
    // enough CTAs per segment to fill the GPU a few times over, never more than the slot has chunks
    long long per_seg = (148 * 16 + nvf - 1) / nvf;
    if (per_seg > schunks) per_seg = schunks;
    if (per_seg < 1) per_seg = 1;
    if (nseg == 1 && (size_t)schunks * sizeof(uint32_t) <= 32768) {
        // (pdl: launched straight behind the fused entropy kernel, as its programmatic dependent)
        const cudaError_t le = launch_pdl(k_stuff_write<true>, dim3((unsigned)nvf, (unsigned)per_seg), dim3(256),
                                          (size_t)schunks * sizeof(uint32_t), st, pdl && counted, unstuffed, total_bits, ffpart, segbase,
                                          sizes, out, ucap, out_cap, schunks, nseg, rst_phase, error_flag, H);
        if (le != cudaSuccess) return cuda_fail(le, "k_stuff_write");
    } else {
        k_ff_scan<<<(unsigned)n_frames, 256, 0, st>>>(ffpart, total_bits, segbase, sizes, ucap, out_cap, schunks, nseg, H.len,
                                                      error_flag);
        B2J_LAUNCH_CHECK("k_ff_scan");
        k_stuff_write<false><<<dim3((unsigned)nvf, (unsigned)per_seg), 256, 0, st>>>(unstuffed, total_bits, ffpart, segbase, sizes, out, ucap,
                                                                                     out_cap, schunks, nseg, rst_phase, error_flag, H);
    }
    B2J_LAUNCH_CHECK("k_stuff_write");
    return B2J_OK;
}

}  // namespace b2j

using namespace b2j;

extern "C" int b2j_block_bits(const int16_t *coefs, const b2j_image_desc *desc, int zrl, int32_t *bits, void *stream)
{
    B2J_REQUIRE(coefs && bits, B2J_E_ARG, "b2j_block_bits: null pointer");
    B2J_REQUIRE(zrl == B2J_ZRL_EAGER || zrl == B2J_ZRL_STANDARD, B2J_E_ARG, "b2j_block_bits: bad zrl policy %d", zrl);
    B2J_REQUIRE(((uintptr_t)coefs & 15) == 0, B2J_E_ALIGN, "b2j_block_bits: coefs must be 16-byte aligned");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    const EGeo e = egeo_of(g);
    const long long total = e.nblk * e.n_frames;
    const LutParam L = make_lut_param();
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (zrl == B2J_ZRL_EAGER) k_block_bits<true><<<grid, 128, 0, as_stream(stream)>>>(coefs, bits, e, L);
    else k_block_bits<false><<<grid, 128, 0, as_stream(stream)>>>(coefs, bits, e, L);
    B2J_LAUNCH_CHECK("k_block_bits");
    return B2J_OK;
}

extern "C" size_t b2j_scan_workspace_bytes(const b2j_image_desc *desc)
{
    Geo g;
    if (make_geo(desc, &g)) return 0;
    return (size_t)g.n_frames * scan_chunks(g.blocks_per_frame) * sizeof(uint32_t) + 256;
}

extern "C" int b2j_scan(const int32_t *bits, const b2j_image_desc *desc, uint32_t *offsets, int64_t *total_bits,
                        void *workspace, void *stream)
{
    B2J_REQUIRE(bits && offsets && total_bits && workspace, B2J_E_ARG, "b2j_scan: null pointer");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    const int chunks = scan_chunks(g.blocks_per_frame);
    uint32_t *partial = reinterpret_cast<uint32_t *>(workspace);
    cudaStream_t st = as_stream(stream);
    dim3 grid((unsigned)chunks, (unsigned)g.n_frames);
    k_chunk_sums<<<grid, 256, 0, st>>>(bits, partial, g.blocks_per_frame, chunks);
    B2J_LAUNCH_CHECK("k_chunk_sums");
    k_scan_partials<<<(unsigned)g.n_frames, 256, 0, st>>>(partial, total_bits, chunks);
    B2J_LAUNCH_CHECK("k_scan_partials");
    k_scan_apply<<<grid, 256, 0, st>>>(bits, partial, offsets, g.blocks_per_frame, chunks);
    B2J_LAUNCH_CHECK("k_scan_apply");
    return B2J_OK;
}


extern "C" size_t b2j_bitpack_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap)
{
    Geo g;
    if (make_geo(desc, &g) || out_cap < kHeaderBytes + 2) return 0;
    const long long ucap = ucap_of(out_cap);
    return (size_t)g.n_frames * (size_t)ucap + (size_t)g.n_frames * (stuff_chunks(ucap) + 1) * sizeof(uint32_t) + 512;
}

extern "C" int b2j_bitpack(const int16_t *coefs, const uint32_t *offsets, const int64_t *total_bits,
                           const b2j_image_desc *desc, const uint8_t qy_host[64], const uint8_t qc_host[64], int zrl,
                           void *workspace, uint8_t *out, int64_t out_cap, int64_t *sizes, void *stream)
{
    B2J_REQUIRE(coefs && offsets && total_bits && qy_host && qc_host && workspace && out && sizes, B2J_E_ARG,
                "b2j_bitpack: null pointer");
    B2J_REQUIRE(zrl == B2J_ZRL_EAGER || zrl == B2J_ZRL_STANDARD, B2J_E_ARG, "b2j_bitpack: bad zrl policy %d", zrl);
    B2J_REQUIRE(out_cap >= kHeaderBytes + 2, B2J_E_CAPACITY, "b2j_bitpack: out_cap %lld smaller than the header", (long long)out_cap);
    B2J_REQUIRE(((uintptr_t)coefs & 15) == 0 && ((uintptr_t)workspace & 15) == 0, B2J_E_ALIGN,
                "b2j_bitpack: coefs and workspace must be 16-byte aligned");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    const EGeo e = egeo_of(g);
    const long long ucap = ucap_of(out_cap);
    uint8_t *unstuffed = reinterpret_cast<uint8_t *>(workspace);
    uint32_t *ffpart = reinterpret_cast<uint32_t *>(unstuffed + (size_t)g.n_frames * (size_t)ucap);
    HeaderParam H;
    memset(&H, 0, sizeof H);
    H.len = jfif_header(g, qy_host, qc_host, H.bytes);
    if (H.len != kHeaderBytes) {
        set_error("b2j_bitpack: internal header length mismatch");
        return B2J_E_ARG;
    }
    const LutParam L = make_lut_param();
    cudaStream_t st = as_stream(stream);
    dim3 pgrid((unsigned)((e.nblk + kPackBlocks - 1) / kPackBlocks), (unsigned)g.n_frames);
    if (zrl == B2J_ZRL_EAGER) k_bitpack<true><<<pgrid, kPackBlocks, 0, st>>>(coefs, offsets, total_bits, unstuffed, ucap, e, L);
    else k_bitpack<false><<<pgrid, kPackBlocks, 0, st>>>(coefs, offsets, total_bits, unstuffed, ucap, e, L);
    B2J_LAUNCH_CHECK("k_bitpack");
    return launch_stuffing(unstuffed, total_bits, ffpart, sizes, out, ucap, out_cap, g.n_frames, H, nullptr, false, st);
}
