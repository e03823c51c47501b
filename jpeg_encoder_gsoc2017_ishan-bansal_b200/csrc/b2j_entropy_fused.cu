// Fused single-pass entropy coder (K2 + K3 + K4 packing in one kernel).
//
// Persistent CTAs take chunks of 256 consecutive blocks of a stream by ticket (a stream is a frame,
// or with [ext] restart intervals one segment of a frame; tickets interleave the streams):
//   1. the chunk's coefficients (+ 8 lookahead blocks) are loaded once, coalesced, and kept in
//      shared memory as saturated BYTES (17-word pitch per block: conflict free for the per-thread
//      reads; the DC values separately as int16); the loading thread also records which of its
//      eight coefficients are non-zero, so every block gets its 64-bit non-zero mask.  A byte at
//      either end of its range sends the walk to the int16 in global memory (L2);
//   2. every thread takes one block and walks ONLY the set bits of the mask (symbols as in
//      b2j_entropy.cuh: gateware run-length rules, Annex-K codes), packing the block's bits
//      MSB-first into a private column of shared memory and counting them.  Which block a thread
//      takes follows the blocks' non-zero counts (warp 7 the long ones ... warp 0 the short ones:
//      "permutation" below), so that the lanes of a warp finish together;
//   3. exclusive scan of the bit counts in block order (warp 0); the chunk total is published and
//      the chunk's bit offset inside the frame is obtained by a decoupled look-back over the
//      chunk totals of its predecessors (one 64-bit status word per chunk: flag | value);
//   4. every thread composes the 32-bit words of the frame's stream whose first bit lies in
//      its block: its own private bits shifted to the final alignment, completed with the
//      first bits of the following blocks (the 8 lookahead blocks cover the chunk's tail).
//      Every output word has exactly one writer: plain stores, no atomics, no zero-init;
//   5. whichever thread composes the stream's final word pads its last byte with 1s (that thread can
//      sit in the previous chunk when the tail is a run of 4-6 bit blocks); total_bits[stream] is
//      written; 0xFF bytes are counted per 4096-byte chunk of the packed stream on the fly.
// 0xFF stuffing, RSTm / EOI markers and JFIF framing then run on the packed stream (launch_stuffing).
#include "b2j_entropy.cuh"

namespace b2j {

#ifndef B2J_ENT_CHUNK
#define B2J_ENT_CHUNK 256
#endif
constexpr int kChunk = B2J_ENT_CHUNK;             // blocks per CTA == threads per CTA
constexpr int kEntWarps = kChunk / 32;
constexpr int kLookahead = 8;                     // >= 31 bits / (4 bits per smallest block)
constexpr int kCols = kChunk + kLookahead;        // blocks held in shared memory
constexpr int kPitch = 17;                        // words per block in shared memory (64 coefficient BYTES + 1 pad)
constexpr int kPrivWords = 55;                    // >= (1728 + 7 pad bits) / 32 + 1
#ifndef B2J_ENT_CAP
#define B2J_ENT_CAP 5
#endif
constexpr int kCap = B2J_ENT_CAP;                 // private words kept in shared memory (160 bits); the rest of a
                                                  // (rare) larger block goes to an L2-resident scratch column
constexpr int kOvfWords = kPrivWords - kCap;
constexpr int kMaxSlots = 2048;                   // persistent CTAs that may own a scratch slot
constexpr unsigned long long kFlagA = 1ull << 62, kFlagP = 2ull << 62, kValMask = (1ull << 62) - 1;
constexpr int kSpinLimit = 1 << 22;
#ifndef B2J_ENT_SMEM_PAD
#define B2J_ENT_SMEM_PAD 0
#endif
constexpr int kLutWords = 32 + 512 + 4;            // dc[2][16], ac[2][256] (walk layout), zrl_eob[2][2]: gen_huff_lut.py
// Blocks are handed to the walking threads by non-zero count (see "permutation" in the kernel); 0 = thread t walks block t
#ifndef B2J_ENT_SORT
#define B2J_ENT_SORT 1
#endif
#ifndef B2J_ENT_EARLY_LB
#define B2J_ENT_EARLY_LB 1
#endif
#if !B2J_ENT_SORT                                  // (it lives in the permuted kernel's walk-to-scan section)
#undef B2J_ENT_EARLY_LB
#define B2J_ENT_EARLY_LB 0
#endif
constexpr int kSortWords = B2J_ENT_SORT ? kChunk + kChunk / 2 : 0;   // exclusive bit offsets [kChunk] + permutation [kChunk] (16 bit)
constexpr int kFusedSmem = (kCols * kPitch + kCap * kCols + kCols + kCols * 2 + kCols / 2 + kLutWords + 32 + kSortWords) * 4 + B2J_ENT_SMEM_PAD;

// The Annex-K code tables in the layout of HuffLut (dc[2][16], ac[2][256]) as a device constant: a CTA copies them with
// coalesced loads (as a kernel parameter every lane read another word of the constant bank: 11 % of a single frame's time).
static __device__ const uint32_t kHuffLutDev[kLutWords] = {
#include "b2j_huff_lut.inc"
};
static const uint32_t kHuffLutHost[kLutWords] = {
#include "b2j_huff_lut.inc"
};

// The same words derived from the host's code tables (the check before the first launch).
static bool fused_lut_matches_host()
{
    const HuffLut &h = host_huff_lut();
    uint32_t w[kLutWords];
    memcpy(w, h.dc, 32 * sizeof(uint32_t));
    for (int t = 0; t < 2; t++) {
        for (int i = 0; i < 256; i++) {
            const int size = (i & 15) + 1;
            const uint32_t e = size <= 10 ? h.ac[t][(i & 0xF0) | size] : 0u;
            const uint32_t len = e >> 16, code = e & 0xFFFFu;
            w[32 + t * 256 + i] = (len ? code << (32 - len) : 0u) | ((len + size) << 8) | len;
        }
        w[544 + 2 * t] = h.ac[t][0xF0];
        w[545 + 2 * t] = h.ac[t][0x00];
    }
    return memcmp(w, kHuffLutHost, sizeof w) == 0;
}

struct FusedEntropyParams {
    const int16_t *coefs;
    uint8_t *unstuffed;
    int64_t *total_bits;
    unsigned long long *status;   // [n_frames * chunks]
    unsigned int *ticket;
    int *error_flag;
    uint32_t *overflow;           // [gridDim.x][kOvfWords][kCols]
    uint32_t *ffpart;             // [n_frames][schunks]: 0xFF bytes per 4096-byte chunk of the packed stream (zeroed)
    int schunks;
    long long ucap;
    int chunks;                   // chunks per segment
    int nseg;                     // [ext] restart intervals: segments per frame (1 = none); a segment is an independent,
    long long frame_blocks;       // byte-aligned stream with its own DC predictors.  e.nblk = blocks per full segment,
    EGeo e;                       // e.n_frames = frames * nseg; frame_blocks = blocks of a whole frame
};

__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t sm_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_s8(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Packs fields MSB-first into a private column of shared memory: word k of column c lives at
// col[k * kCols] (word-interleaved, so any mix of k across the lanes of a warp is conflict free).  The column is walked
// with a 32-bit shared address; words kCap.. continue, at the same stride, in a global scratch column.
constexpr uint32_t kColStride = kCols * 4u;       // bytes between consecutive words of a column
struct PrivSink {
    uint32_t p0, p;               // shared address of the column's first word, and of the word being filled (virtual past kCap words)
    uint32_t *ovf;                // global scratch column (words kCap..)
    uint32_t acc;                 // pending bits, left aligned
    int nb;                       // number of pending bits (< 32)
    __device__ __forceinline__ void begin(uint32_t *c, uint32_t *o) { p0 = p = sm_addr(c); ovf = o; acc = 0u; nb = 0; }
    __device__ __forceinline__ void store()
    {
        if (p < p0 + kCap * kColStride) sts_u32(p, acc);
        else if (p < p0 + kPrivWords * kColStride)                   // (coefficients outside the JPEG range: drop, never write past the column)
            *reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(ovf) + (p - p0 - kCap * kColStride)) = acc;
    }
    __device__ __forceinline__ uint32_t load(int i) const { return i < kCap ? lds_u32(p0 + (uint32_t)i * kColStride) : ovf[(i - kCap) * kCols]; }
    __device__ __forceinline__ void put_left(uint32_t x, int len)   // x: the field left aligned, 0 <= len <= 32
    {
        acc |= x >> nb;
        const uint32_t spill = __funnelshift_r(0u, x, nb);           // the bits that do not fit: x << (32 - nb), 0 for nb = 0
        nb += len;
        if (nb >= 32) {
            store();
            p += kColStride;
            acc = spill;
            nb -= 32;
        }
    }
    __device__ __forceinline__ void put(uint32_t val, int len)      // 1 <= len <= 32, val < 2^len
    {
        put_left(val << (32 - len), len);
    }
    __device__ __forceinline__ int words() const { return (int)((p - p0) / kColStride); }
    __device__ __forceinline__ int bits() const { return words() * 32 + nb; }
    __device__ __forceinline__ void flush() { store(); }            // zero padded partial word
};

// Sparse walk of one block: only the non-zero AC positions (set bits of the mask) are visited.
// blk: the block's 64 coefficients as saturated bytes in shared memory; a byte at either end of the range (+127 / -128)
// stands for "read the int16 from the coefficient array" (g16: the block in global memory; rare at any quality).
// lut_ac: the walk layout of gen_huff_lut.py (index run << 4 | size - 1; code left aligned, field length in byte 1,
// code length in the low bits); zrl_eob: the two codes without amplitude bits as (length << 16) | code.
// With PARTIAL the walk stops once 32 bits are out (lookahead blocks only have to complete the previous word).
// Shared memory through 32-bit addresses; ONE loop for all lanes of a warp.  (A short first loop for the common case - no
// ZRL, words still inside the shared column - followed by this general one was 0.90 instead of 0.93 ms on the headline
// batch, but lanes of a warp in different loops run one after the other: 8K noise 2.5 -> 4.0 ms.  Picking the variant per
// warp from the symbol counts cost more in code size than it gave: profiles/r02/experiments.txt.)
template <bool EAGER, bool PARTIAL>
__device__ __forceinline__ void encode_block(uint32_t mlo, uint32_t mhi, int diff, const uint32_t *blk, const int16_t *__restrict__ g16,
                                             const uint32_t *__restrict__ lut_dc, const uint32_t *__restrict__ lut_ac,
                                             const uint32_t *__restrict__ zrl_eob, int *error_flag, PrivSink &sink)
{
    {
        const int s = bit_size(diff);
        const uint32_t e = lut_dc[s];
        const uint32_t extra = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u);
        sink.put(((e & 0xFFFFu) << s) | extra, (int)(e >> 16) + s);
    }
    const uint32_t blk_s = sm_addr(blk), lut_s = sm_addr(lut_ac);
    const uint32_t zrl = zrl_eob[0], eob = zrl_eob[1];
    int prev = 0;
    uint32_t m = mlo & ~1u, mnext = mhi;                            // position 0 is the DC
    int base = 0;
    for (;;) {
#pragma unroll 1
        while (m) {
            if (PARTIAL && sink.p != sink.p0) return;
            const int z = base + __ffs((int)m) - 1;
            m &= m - 1u;
            int v = lds_s8(blk_s + (uint32_t)z);
            if ((unsigned)(v + 127) >= 254u) {                      // saturated byte: the real value
                v = (int)__ldg(g16 + z);
                if (abs(v) > 1023) atomicExch(error_flag, 1);       // no baseline table codes it -> every sizes[] becomes INT64_MIN
            }
            int run = z - prev - 1;
            prev = z;
            if (run >= 16) {                                        // ZRLs in front of a non-zero: same for both policies
#pragma unroll 1
                for (int n = run >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));
                run &= 15;
            }
            // amplitude bits (v - 1 for negative values, T.81 F.1.2.2) left aligned by the position of the top bit of |v|
            uint32_t top;
            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(abs(v)));     // size - 1
            const uint32_t e = lds_u32(lut_s + (uint32_t)run * 64u + top * 4u);
            const uint32_t amp = (uint32_t)(v + (v >> 31)) << (31u - top);
            // (the shift uses the low five bits of e, the code length; byte 1 is the length of the field)
            sink.put_left((e & 0xFFFF0000u) | __funnelshift_r(amp, 0u, e), (int)__byte_perm(e, 0u, 0x4441));
        }
        if (base) break;
        m = mnext;
        mnext = 0u;
        base = 32;
    }
    if (PARTIAL && sink.p != sink.p0) return;
    const int t = 63 - prev;                                        // trailing zeros
    if (EAGER) {                                                    // rlecore.py:101-114
#pragma unroll 1
        for (int n = t >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));
        if (t & 15) sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    } else if (t > 0) {
        sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    }
}

#ifndef B2J_ENT_MIN_CTAS
#define B2J_ENT_MIN_CTAS 5
#endif
template <bool EAGER, bool SEG>
__global__ void __launch_bounds__(kChunk, B2J_ENT_MIN_CTAS) k_entropy_fused(const __grid_constant__ FusedEntropyParams P)
{
    extern __shared__ __align__(16) uint32_t fsm[];
    uint32_t *s_coef = fsm;                                  // [kCols][kPitch]: coefficients as saturated bytes
    uint32_t *s_priv = s_coef + kCols * kPitch;              // [kCap][kCols]
    uint32_t *s_nbits = s_priv + kCap * kCols;               // [kCols]
    uint32_t *s_mask = s_nbits + kCols;                      // [kCols][2]: eight mask bytes per block
    int16_t *s_dc16 = reinterpret_cast<int16_t *>(s_mask + kCols * 2);   // [kCols]: the DC coefficients, unsaturated
    uint32_t *s_dc = s_mask + kCols * 2 + kCols / 2;         // [2][16]
    uint32_t *s_ac = s_dc + 32;                              // [2][256], the walk's layout
    uint32_t *s_ze = s_ac + 512;                             // [2][2]: ZRL and EOB
    uint32_t *s_misc = s_ze + 4;                             // warp sums [0..15], ticket [16], prefix lo/hi [17..18]
#if B2J_ENT_SORT
    uint32_t *s_excl = s_misc + 32;                          // [kChunk]: bit offset of every column inside the chunk
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(s_excl + kChunk);   // [kChunk]: the column thread t walks and composes
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *my_ovf = P.overflow + (size_t)blockIdx.x * kOvfWords * kCols;
    const unsigned int total_chunks = (unsigned)P.chunks * (unsigned)P.e.n_frames;
    pdl_launch_dependents();                                                  // (the stuffing kernel's CTAs may queue up)
    for (int i = tid; i < kLutWords; i += kChunk) s_dc[i] = kHuffLutDev[i];   // s_ac and s_ze follow s_dc
    pdl_wait();                                                               // K1's coefficients are complete and visible
    const unsigned bpm_recip = 65536u / (unsigned)P.e.bpm + 1u;                // n / bpm for n < 4096
  for (;;) {                                                          // persistent: one chunk per ticket
    __syncthreads();                                                  // shared memory of the previous chunk is free
    if (tid == 0) s_misc[16] = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const unsigned int id = s_misc[16];
    if (id >= total_chunks) break;
    // tickets interleave the streams: consecutive CTAs work on different streams, so the predecessor
    // chunk of the same stream was started n_frames tickets earlier and has usually published its
    // inclusive prefix by the time it is needed (short look-back, no serial chain inside a wave).
    // A stream is a frame or, with [ext] restart intervals (SEG), one segment of a frame: segment
    // f % nseg of frame f / nseg; the last segment of a frame may be short.
    const int f = (int)(id % (unsigned)P.e.n_frames), ch = (int)(id / (unsigned)P.e.n_frames);
    const EGeo &e = P.e;
    long long nblk = e.nblk;
    const int16_t *fc;
    if (SEG) {
        const long long seg0 = (long long)(f % P.nseg) * e.nblk;
        nblk = min(e.nblk, P.frame_blocks - seg0);
        fc = P.coefs + ((long long)(f / P.nseg) * P.frame_blocks + seg0) * 64;
    } else {
        fc = P.coefs + (long long)f * e.nblk * 64;
    }
    const long long b0 = (long long)ch * kChunk;
    if (SEG && b0 >= nblk) continue;                                 // chunk beyond a short last segment
    const long long b1 = min(b0 + kChunk, nblk);
    const long long bl = min(b1 + kLookahead, nblk);                // end of the lookahead blocks
    const int nown = (int)(b1 - b0), nall = (int)(bl - b0);
    const bool last = b1 >= nblk;

    // ---- 1. coalesced load (16-byte pieces) into the padded layout + non-zero byte masks
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(fc + b0 * 64);
        uint8_t *mask8 = reinterpret_cast<uint8_t *>(s_mask);
        const int npieces = nall * 8;
#ifndef B2J_ENT_LOAD_BATCH
#define B2J_ENT_LOAD_BATCH 2
#endif
        // the loads of a batch are issued together: the phase is bound by memory latency (9 dependent round trips
        // per chunk otherwise)
        auto stage_piece = [&](int i, const uint4 q) {
            const int b = i >> 3, c = i & 7;
            // shared memory holds bytes: both halves of a word clamped to -128..127 (packed 16-bit min / max), the low
            // bytes of the four words gathered by two permutes.  Halving the staging area is what lets 13 instead of 8
            // CTAs share an SM, and the kernel's speed follows its occupancy (profiles/r02/experiments.txt)
            uint32_t t0, t1, t2, t3;
            asm("min.s16x2 %0, %1, %2;" : "=r"(t0) : "r"(q.x), "r"(0x007F007Fu));
            asm("min.s16x2 %0, %1, %2;" : "=r"(t1) : "r"(q.y), "r"(0x007F007Fu));
            asm("min.s16x2 %0, %1, %2;" : "=r"(t2) : "r"(q.z), "r"(0x007F007Fu));
            asm("min.s16x2 %0, %1, %2;" : "=r"(t3) : "r"(q.w), "r"(0x007F007Fu));
            asm("max.s16x2 %0, %1, %2;" : "=r"(t0) : "r"(t0), "r"(0xFF80FF80u));
            asm("max.s16x2 %0, %1, %2;" : "=r"(t1) : "r"(t1), "r"(0xFF80FF80u));
            asm("max.s16x2 %0, %1, %2;" : "=r"(t2) : "r"(t2), "r"(0xFF80FF80u));
            asm("max.s16x2 %0, %1, %2;" : "=r"(t3) : "r"(t3), "r"(0xFF80FF80u));
            uint32_t *d = s_coef + b * kPitch + c * 2;
            d[0] = __byte_perm(t0, t1, 0x6420);
            d[1] = __byte_perm(t2, t3, 0x6420);
            if (c == 0) s_dc16[b] = (int16_t)(q.x & 0xFFFFu);
            // non-zero flags of the eight halves: min(half, 1) per half (one packed instruction per word), the four
            // words interleaved by multiply-adds: low halves at bits 0, 2, 4, 6, high halves at bits 16, 18, 20, 22
            uint32_t n0, n1, n2, n3;
            asm("min.u16x2 %0, %1, %2;" : "=r"(n0) : "r"(q.x), "r"(0x00010001u));
            asm("min.u16x2 %0, %1, %2;" : "=r"(n1) : "r"(q.y), "r"(0x00010001u));
            asm("min.u16x2 %0, %1, %2;" : "=r"(n2) : "r"(q.z), "r"(0x00010001u));
            asm("min.u16x2 %0, %1, %2;" : "=r"(n3) : "r"(q.w), "r"(0x00010001u));
            const uint32_t mm = n0 + 4u * n1 + 16u * n2 + 64u * n3;
            const uint32_t m = (mm | (mm >> 15)) & 0xFFu;
            mask8[b * 8 + c] = (uint8_t)m;
        };
        for (int i0 = tid; i0 < npieces; i0 += B2J_ENT_LOAD_BATCH * kChunk) {
            uint4 qb[B2J_ENT_LOAD_BATCH];
#pragma unroll
            for (int j = 0; j < B2J_ENT_LOAD_BATCH; j++)
                if (i0 + j * kChunk < npieces) qb[j] = __ldg(src + i0 + j * kChunk);
#pragma unroll
            for (int j = 0; j < B2J_ENT_LOAD_BATCH; j++) {
                const int i = i0 + j * kChunk;
                if (i >= npieces) break;
                stage_piece(i, qb[j]);
            }
        }
    }
    // k0: position of column 0 inside its MCU
    const int k0 = (int)((unsigned)b0 % (unsigned)e.bpm);
    auto mcu_pos = [&](int col) { const unsigned a = (unsigned)(k0 + col); return (int)(a - ((a * bpm_recip) >> 16) * (unsigned)e.bpm); };
#if B2J_ENT_SORT
    // ---- permutation: a warp's time in the walk is that of its longest block, and blocks of very different lengths sit next
    // to each other (19 of 32 lanes busy when thread t walks block t).  The pieces a warp has just loaded belong to blocks
    // 32 a + 4 w + k (a = 0..7, k = 0..3): for every k the warp ranks those eight blocks by non-zero count (seven
    // shuffles, ties by lane), and the block of rank r (0 = longest) goes to thread 32 (7 - r) + 4 w + k: warp 7 walks the
    // longest block of every group, warp 0 the shortest.  Warp 0 is the one with work after the walk (the lookahead
    // blocks, the scan, the look-back) while the other warps wait for it, so it gets the light end: with the long blocks
    // on warp 0 the permutation was 0.87 ms on the headline batch, with the short ones 0.81 (consecutive blocks: 0.89).
    // The walking thread's lane is still the block's index modulo 32, so the shared-memory bank pattern of the walk is
    // unchanged.  No barrier of its own (the masks were written by the warp itself) and no CTA-wide sort; columns, bit
    // counts, the scan and the composition stay in block order.
    {
        static_assert(kEntWarps == 8 || kEntWarps == 4, "the ranking is written for eight or four warps of 32 blocks");
        __syncwarp();
        // (with four warps a warp's pieces belong to blocks 16 a + 4 w + k: the four blocks of one residue modulo 32 are
        // those of every other a, and they are ranked among themselves)
        constexpr int kStep = 32 / kEntWarps;                        // lane distance between two blocks of one group
        const int myblk = (lane >> 2) * (kChunk / 8) + warp * 4 + (lane & 3);
        int key = 0;                                                 // absent blocks of a short chunk rank last
        if (myblk < nown) key = 1 + __popc(s_mask[myblk * 2] & ~1u) + __popc(s_mask[myblk * 2 + 1]);
        key = key * 32 + (31 - lane);                                // ties: the lower lane first
        int r = 0;
#pragma unroll
        for (int d = 1; d < kEntWarps; d++) r += __shfl_xor_sync(0xFFFFFFFFu, key, kStep * d) > key ? 1 : 0;
        s_perm[(kEntWarps - 1 - r) * 32 + (myblk & 31)] = (uint16_t)myblk;
    }
#endif
    __syncthreads();

    // ---- 2. one walk per block: private packed bits + bit count
    auto dc_of = [&](long long b) -> int {          // DC of block b of this frame (b < bl)
        if (b >= b0) return (int)s_dc16[(int)(b - b0)];
        return b >= 0 ? (int)fc[b * 64] : 0;
    };
    auto prev_index = [&](long long b, int k) -> long long {        // k: position of block b inside its MCU
        if (k < e.hv) return k > 0 ? b - 1 : b - e.bpm + e.hv - 1;
        return b - e.bpm;
    };
#if B2J_ENT_SORT
    int col = (int)s_perm[tid];                                      // the column this thread walks
#else
    const int col = tid;                                             // the column this thread walks and composes
#endif
    PrivSink sink;
    sink.begin(s_priv + col, my_ovf + col);
    int mybits = 0;
    if (col < nown) {
        const long long b = b0 + col;
        const uint32_t *blk = s_coef + col * kPitch;
        const int k = mcu_pos(col), chroma = k >= e.hv;
        const int diff = (int)s_dc16[col] - dc_of(prev_index(b, k));
        encode_block<EAGER, false>(s_mask[col * 2], s_mask[col * 2 + 1], diff, blk, fc + b * 64, s_dc + 16 * chroma, s_ac + 256 * chroma,
                                   s_ze + 2 * chroma, P.error_flag, sink);
        sink.flush();
        int bits = sink.bits();
        if (bits > (kPrivWords - 1) * 32) {                         // only with coefficients no baseline table can code
            atomicExch(P.error_flag, 1);                            // -> every sizes[] becomes INT64_MIN
            bits = (kPrivWords - 1) * 32;
        }
        s_nbits[col] = (uint32_t)bits;
        mybits = bits;
    }
#if B2J_ENT_SORT
    else s_nbits[col] = 0u;                                          // (the scan reads every column)
#endif
    // the lookahead blocks only have to supply the bits that complete this chunk's last word
    if (tid < kLookahead && kChunk + tid < nall && !last) {
        const int colidx = kChunk + tid;
        const long long b = b0 + colidx;
        const uint32_t *blk = s_coef + colidx * kPitch;
        const int k = mcu_pos(colidx), chroma = k >= e.hv;
        const int diff = (int)s_dc16[colidx] - dc_of(prev_index(b, k));
        PrivSink sk;
        sk.begin(s_priv + colidx, nullptr);
        encode_block<EAGER, true>(s_mask[colidx * 2], s_mask[colidx * 2 + 1], diff, blk, fc + b * 64, s_dc + 16 * chroma, s_ac + 256 * chroma,
                                  s_ze + 2 * chroma, P.error_flag, sk);
        sk.flush();
        s_nbits[colidx] = (uint32_t)sk.bits();                      // >= 32 when the walk was cut short
    }

    // ---- 3. CTA exclusive scan of the bit counts
#if B2J_ENT_SORT
#if B2J_ENT_EARLY_LB
    // warp 0 is through with its (short) blocks long before the others: it reads the predecessor chunk's status word now,
    // so that the L2 round trip runs while it waits at the barrier.  In a batch the predecessor was started n_streams
    // tickets earlier and has published its inclusive prefix: the look-back behind the barrier then has nothing to wait for
    // (256 x 1080p: 0.803 -> 0.791 ms).  With few streams the predecessor is still at work and the read would be wasted (one
    // frame: +0.9 us), so it is only issued when the streams are at least a quarter of the CTAs in flight
    unsigned long long early = 0ull;
    if (warp == 0 && ch > 0 && (unsigned)P.e.n_frames * 4u >= gridDim.x) early = ld_volatile(P.status + (long long)f * P.chunks + ch - 1);
#endif
    // in block order, by warp 0 alone (it also does the look-back): lane l sums columns 8 l .. 8 l + 7
    __syncthreads();
    uint32_t T = 0;
    if (warp == 0) {
        constexpr int kPer = kChunk / 32;
        static_assert(kPer % 4 == 0 && 32 % kEntWarps == 0 && (kCap * kCols) % 4 == 0 && (kCols * kPitch) % 4 == 0 && kCols % 4 == 0,
                      "chunk size / column count do not fit the vector accesses of the scan");
        uint4 v[kPer / 4];
#pragma unroll
        for (int k = 0; k < kPer / 4; k++) v[k] = reinterpret_cast<const uint4 *>(s_nbits + lane * kPer)[k];
        uint32_t sum = 0u;
#pragma unroll
        for (int k = 0; k < kPer / 4; k++) {                         // v becomes the exclusive prefix inside the lane
            uint32_t x;
            x = v[k].x; v[k].x = sum; sum += x;
            x = v[k].y; v[k].y = sum; sum += x;
            x = v[k].z; v[k].z = sum; sum += x;
            x = v[k].w; v[k].w = sum; sum += x;
        }
        const uint32_t inc = warp_incl_scan(sum);
        T = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const uint32_t base = inc - sum;
#pragma unroll
        for (int k = 0; k < kPer / 4; k++)
            reinterpret_cast<uint4 *>(s_excl + lane * kPer)[k] = make_uint4(v[k].x + base, v[k].y + base, v[k].z + base, v[k].w + base);
    }
#else
    const uint32_t inc = warp_incl_scan((uint32_t)mybits);
    if (lane == 31) s_misc[warp] = inc;
    __syncthreads();
    uint32_t wbase_bits = 0, T = 0;
#pragma unroll
    for (int w = 0; w < kEntWarps; w++) {
        const uint32_t x = s_misc[w];
        if (w < warp) wbase_bits += x;
        T += x;
    }
    const uint32_t excl = wbase_bits + inc - (uint32_t)mybits;
#endif

    // ---- decoupled look-back over the chunk totals of this frame (warp 0)
    if (warp == 0) {
        unsigned long long *st = P.status + (long long)f * P.chunks;
        unsigned long long E = 0;
#if B2J_ENT_SORT && B2J_ENT_EARLY_LB
        if ((early >> 62) == 2ull) E = early & kValMask;
        else
#endif
        if (ch > 0) {
            if (lane == 0) st_volatile(st + ch, kFlagA | (unsigned long long)T);
            int i = ch - 1;
            bool done = false;
            while (!done) {
                const int idx = i - lane;
                unsigned long long v = idx >= 0 ? ld_volatile(st + idx) : kFlagP;
                int spins = 0;
                while (__any_sync(0xffffffffu, (v >> 62) == 0ull)) {
                    if (++spins > kSpinLimit) { if (lane == 0) atomicExch(P.error_flag, 1); v = kFlagP; break; }
                    __nanosleep(40);
                    if (idx >= 0) v = ld_volatile(st + idx);
                }
                const unsigned pmask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                const int first_p = pmask ? __ffs((int)pmask) - 1 : 32;
                unsigned long long val = lane <= first_p ? (v & kValMask) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                E += val;
                done = pmask != 0u;
                i -= 32;
            }
        }
        if (lane == 0) {
            st_volatile(st + ch, kFlagP | (E + (unsigned long long)T));
            s_misc[17] = (uint32_t)E;
            s_misc[18] = (uint32_t)(E >> 32);
        }
    }
    __syncthreads();
    const unsigned long long E = (unsigned long long)s_misc[17] | ((unsigned long long)s_misc[18] << 32);

    // ---- 4. compose the owned words.  Bit g of the frame's stream lives in word g >> 5.
    // (bit offsets inside a frame fit 32 bits: make_geo() refuses larger frames)
    // The frame's final byte is padded with 1s by whichever thread composes the frame's final word:
    // that is not always the last block's thread, nor even a thread of the last chunk (a short
    // tail of 4-6 bit blocks can sit entirely inside a word owned by the previous chunk, which
    // sees those blocks as its lookahead columns).
    const bool reaches_end = bl >= nblk;                             // columns 0..nall-1 run to the stream's end
#if B2J_ENT_SORT
    // in block order again: the words a warp stores are neighbours in the stream
    col = tid;
    if (col < nown) {
        mybits = (int)s_nbits[col];
        sink.begin(s_priv + col, my_ovf + col);
    }
#endif
    if (col < nown) {
#if B2J_ENT_SORT
        const uint32_t excl = s_excl[col];
#endif
        const uint32_t total = (uint32_t)mybits;
        if (last && col == nown - 1) P.total_bits[f] = (int64_t)((uint32_t)E + excl + total);
        const uint32_t G = (uint32_t)E + excl;
        uint32_t *dst = reinterpret_cast<uint32_t *>(P.unstuffed + (long long)f * P.ucap);
        const uint32_t wcap = (uint32_t)(P.ucap >> 2);
        uint32_t *ffp = P.ffpart + (long long)f * P.schunks;
        const int nsucc = last ? nown : nall;                        // columns that may supply following bits
        uint32_t w = (G + 31u) >> 5;
        const uint32_t wend = (G + total + 31u) >> 5;
        int rel = (int)((w << 5) - G);                               // first bit of word w inside my block
        uint32_t a = sink.load(rel >> 5);
        uint32_t n_ff = 0u, ff_chunk = w >> 10;
#pragma unroll 1
        for (; w < wend; w++, rel += 32) {
            const int k = rel >> 5, sh = rel & 31;
            const int have0 = (int)total - rel;                      // my bits available from `rel`
            const uint32_t bnext = have0 > 32 - sh ? sink.load(k + 1) : 0u;
            uint32_t v = __funnelshift_l(bnext, a, sh);
            a = bnext;
            if (have0 < 32) {                                        // complete the word with the following blocks' first bits
                int have = have0;
#pragma unroll 1
                for (int nb = col + 1; have < 32 && nb < nsucc; nb++) {
                    v |= s_priv[nb] >> have;
                    have += (int)s_nbits[nb];
                }
                if (have < 32 && reaches_end) {                      // ran out of blocks: this is the frame's final word
                    const int pad = (8 - (have & 7)) & 7;
                    v |= ((1u << pad) - 1u) << (32 - have - pad);
                }
            }
            if (w < wcap) {
                dst[w] = __byte_perm(v, 0u, 0x0123);                 // big-endian bytes
                const uint32_t z = ~v;                               // a zero byte of z is a 0xFF byte of the stream
                if ((z - 0x01010101u) & ~z & 0x80808080u) {          // (some byte is zero; which ones, exactly:)
                    const uint32_t zb = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);
                    if ((w >> 10) != ff_chunk) {                     // 1024 words per stuffing chunk: a block can straddle two
                        if (n_ff) atomicAdd(ffp + ff_chunk, n_ff);
                        n_ff = 0u;
                        ff_chunk = w >> 10;
                    }
                    n_ff += (uint32_t)__popc(zb);
                }
            }
        }
        if (n_ff) atomicAdd(ffp + ff_chunk, n_ff);                   // one atomic per block with 0xFF bytes, not one per word
    }
  }
}

}  // namespace b2j

using namespace b2j;

// [ext] restart intervals: the option word's bits 16..30 give the interval in MCUs.  A frame then
// is nseg independent streams ("segments") of `seg_blocks` blocks (the last one may be shorter).
struct SegPlan {
    int ri, nseg, chunks;         // interval (MCUs), segments per frame, 128-block chunks per segment
    long long seg_blocks, ucap;   // blocks per full segment, unstuffed capacity per segment (bytes)
};

static SegPlan seg_plan(const Geo &g, int64_t out_cap, int opts)
{
    SegPlan sp;
    const long long n_mcu = (long long)g.mcus_x * g.mcus_y;
    sp.ri = (opts >> 16) & 0x7FFF;
    sp.nseg = (sp.ri > 0 && sp.ri < n_mcu) ? (int)((n_mcu + sp.ri - 1) / sp.ri) : 1;
    sp.seg_blocks = sp.nseg > 1 ? (long long)sp.ri * g.bpm : g.blocks_per_frame;
    sp.ucap = ucap_of(out_cap);
    if (sp.nseg > 1) {            // a block never needs more than 208 bytes (11+11 + 63 x (16+10) bits)
        const long long worst = (sp.seg_blocks * 256 + 64 + 15) & ~15LL;
        if (worst < sp.ucap) sp.ucap = worst;
    }
    sp.chunks = (int)((sp.seg_blocks + kChunk - 1) / kChunk);
    return sp;
}

struct FusedWs { size_t status, ffpart, zero_bytes, total_bits, unstuffed, overflow, total; };

static FusedWs fused_ws(const Geo &g, const SegPlan &sp)
{
    FusedWs w;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t nvf = (size_t)g.n_frames * sp.nseg;
    w.status = 256;                                                     // [0,256): ticket, error flag
    w.ffpart = up(w.status + nvf * sp.chunks * sizeof(unsigned long long));
    // 0xFF counts per 4096-byte chunk, then one segment base each; [0, zero_bytes) is cleared per call
    w.zero_bytes = up(w.ffpart + nvf * (stuff_chunks(sp.ucap) + 1) * sizeof(uint32_t));
    w.total_bits = w.zero_bytes;
    w.unstuffed = up(w.total_bits + nvf * sizeof(int64_t));
    w.overflow = up(w.unstuffed + nvf * (size_t)sp.ucap);
    // one scratch slot per persistent CTA: never more CTAs than chunks (a small frame needs kilobytes, not 56 MB)
    const size_t total_chunks = nvf * (size_t)sp.chunks;
    w.total = w.overflow + (total_chunks < (size_t)kMaxSlots ? total_chunks : (size_t)kMaxSlots) * kOvfWords * kCols * sizeof(uint32_t);
    return w;
}

extern "C" size_t b2j_entropy_workspace_bytes_opts(const b2j_image_desc *desc, int64_t out_cap, int opts)
{
    Geo g;
    if (make_geo(desc, &g) || out_cap < kHeaderMax + 2) return 0;
    return fused_ws(g, seg_plan(g, out_cap, opts)).total;
}

extern "C" size_t b2j_entropy_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap)
{
    return b2j_entropy_workspace_bytes_opts(desc, out_cap, 0);
}

bool b2j::entropy_try_clear(const b2j_image_desc *desc, int64_t out_cap, int opts, void *workspace, cudaStream_t st)
{
    Geo g;
    if (!desc || !workspace || ((uintptr_t)workspace & 255) != 0 || out_cap < kHeaderMax + 2) return false;
    if ((opts & ~(0xFF | B2J_OPT_RST_PHASE(7) | B2J_OPT_RESTART(0x7FFF))) != 0 || make_geo(desc, &g)) return false;
    const FusedWs W = fused_ws(g, seg_plan(g, out_cap, opts));
    if (cudaMemsetAsync(workspace, 0, W.zero_bytes, st) != cudaSuccess) {
        (void)cudaGetLastError();                                        // the entropy call clears (and reports) itself
        return false;
    }
    return true;
}

// K2..K4 in one pass over the coefficients: coefs -> JFIF files.
extern "C" int b2j_entropy_encode(const int16_t *coefs, const b2j_image_desc *desc, const uint8_t qy_host[64],
                                  const uint8_t qc_host[64], int opts, void *workspace, uint8_t *out, int64_t out_cap,
                                  int64_t *sizes, void *stream)
{
    return entropy_encode_impl(coefs, desc, qy_host, qc_host, opts, workspace, out, out_cap, sizes, stream, false);
}

// after_k1: the workspace head was cleared by entropy_try_clear before K1 (the launch in front of this one on the stream).
int b2j::entropy_encode_impl(const int16_t *coefs, const b2j_image_desc *desc, const uint8_t qy_host[64],
                             const uint8_t qc_host[64], int opts, void *workspace, uint8_t *out, int64_t out_cap,
                             int64_t *sizes, void *stream, bool after_k1)
{
    const int zrl = opts & 0xFF;
    B2J_REQUIRE(coefs && qy_host && qc_host && workspace && out && sizes, B2J_E_ARG, "b2j_entropy_encode: null pointer");
    B2J_REQUIRE(zrl == B2J_ZRL_EAGER || zrl == B2J_ZRL_STANDARD, B2J_E_ARG, "b2j_entropy_encode: bad zrl policy %d", zrl);
    B2J_REQUIRE((opts & ~(0xFF | B2J_OPT_RST_PHASE(7) | B2J_OPT_RESTART(0x7FFF))) == 0, B2J_E_ARG,
                "b2j_entropy_encode: unknown option bits 0x%x", opts);
    B2J_REQUIRE(out_cap >= kHeaderMax + 2, B2J_E_CAPACITY, "b2j_entropy_encode: out_cap %lld smaller than the header",
                (long long)out_cap);
    B2J_REQUIRE(((uintptr_t)coefs & 15) == 0 && ((uintptr_t)workspace & 255) == 0, B2J_E_ALIGN,
                "b2j_entropy_encode: coefs must be 16-byte and workspace 256-byte aligned");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    const SegPlan sp = seg_plan(g, out_cap, opts);
    const long long nvf = (long long)g.n_frames * sp.nseg;
    B2J_REQUIRE(nvf * sp.chunks < 0xFFFFFFFFLL && nvf < 0x7FFFFFFFLL, B2J_E_SHAPE,
                "b2j_entropy_encode: %lld restart segments x %d chunks exceed the ticket range", nvf, sp.chunks);
    const long long ucap = sp.ucap;
    const int chunks = sp.chunks;
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    const FusedWs W = fused_ws(g, sp);
    FusedEntropyParams P;
    P.coefs = coefs;
    P.ticket = reinterpret_cast<unsigned int *>(ws);
    P.error_flag = reinterpret_cast<int *>(ws + 64);
    P.status = reinterpret_cast<unsigned long long *>(ws + W.status);
    P.ffpart = reinterpret_cast<uint32_t *>(ws + W.ffpart);
    P.schunks = stuff_chunks(ucap);
    P.total_bits = reinterpret_cast<int64_t *>(ws + W.total_bits);
    P.unstuffed = ws + W.unstuffed;
    P.overflow = reinterpret_cast<uint32_t *>(ws + W.overflow);
    uint32_t *ffpart = P.ffpart;
    const size_t status_bytes = W.zero_bytes;
    P.ucap = ucap;
    P.chunks = chunks;
    P.nseg = sp.nseg;
    P.frame_blocks = g.blocks_per_frame;
    P.e = egeo_of(g);
    P.e.nblk = sp.seg_blocks;
    P.e.n_frames = (int)nvf;
    HeaderParam H;
    memset(&H, 0, sizeof H);
    H.len = jfif_header(g, qy_host, qc_host, H.bytes, sp.ri);
    if (H.len != kHeaderBytes + (sp.ri > 0 ? 6 : 0)) {
        set_error("b2j_entropy_encode: internal header length mismatch");
        return B2J_E_ARG;
    }
    static const bool lut_ok = fused_lut_matches_host();
    B2J_REQUIRE(lut_ok, B2J_E_ARG, "b2j_entropy_encode: b2j_huff_lut.inc is not the table the host code derives");
    cudaStream_t st = as_stream(stream);
    if (!after_k1) B2J_CUDA(cudaMemsetAsync(ws, 0, status_bytes, st));   // ticket, error flag, chunk status words
    const long long total = (long long)chunks * nvf;
    using Kernel = void (*)(const FusedEntropyParams);
    const bool seg = sp.nseg > 1;
    const int which = (zrl == B2J_ZRL_EAGER ? 0 : 2) + (seg ? 1 : 0);
    static const Kernel kernels[4] = {k_entropy_fused<true, false>, k_entropy_fused<true, true>, k_entropy_fused<false, false>,
                                      k_entropy_fused<false, true>};
    const Kernel kernel = kernels[which];
    static LaunchInfo cache[4][kMaxDevices];                          // no attribute queries on the hot path
    LaunchInfo li;
    if ((rc = launch_info(kernel, kChunk, kFusedSmem, cache[which], &li))) return rc;
    long long slots = (long long)li.sms * li.per_sm;
    if (slots > kMaxSlots) slots = kMaxSlots;
    const unsigned grid = (unsigned)(total < slots ? total : slots);   // persistent CTAs, one scratch slot each
    {
        const cudaError_t le = launch_pdl(kernel, dim3(grid), dim3(kChunk), (size_t)kFusedSmem, st, after_k1, P);
        if (le != cudaSuccess) return cuda_fail(le, "k_entropy_fused");
    }
    return launch_stuffing(P.unstuffed, P.total_bits, ffpart, sizes, out, ucap, out_cap, g.n_frames, H, P.error_flag, true, st,
                           sp.nseg, (opts >> 12) & 7, true);
}
