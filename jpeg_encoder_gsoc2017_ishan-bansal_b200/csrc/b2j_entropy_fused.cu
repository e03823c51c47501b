// Fused single-pass entropy coder (K2 + K3 + K4 packing in one kernel).
//
// One CTA per chunk of 128 consecutive blocks of a frame, in launch order (ticket):
//   1. the chunk's coefficients (+ 8 lookahead blocks) are loaded once, coalesced, into
//      shared memory (33-word pitch per block: conflict free both for the cooperative
//      stores and for the per-thread reads); the loading thread also records which of its
//      eight coefficients are non-zero, so every block gets its 64-bit non-zero mask;
//   2. every thread takes one block and walks ONLY the set bits of the mask (symbols as in
//      b2j_entropy.cuh: gateware run-length rules, Annex-K codes), packing the block's bits
//      MSB-first into a thread-private column of shared memory and counting them;
//   3. warp-shuffle + CTA exclusive scan of the bit counts; the chunk total is published and
//      the chunk's bit offset inside the frame is obtained by a decoupled look-back over the
//      chunk totals of its predecessors (one 64-bit status word per chunk: flag | value);
//   4. every thread composes the 32-bit words of the frame's stream whose first bit lies in
//      its block: its own private bits shifted to the final alignment, completed with the
//      first bits of the following blocks (the 8 lookahead blocks cover the chunk's tail).
//      Every output word has exactly one writer: plain stores, no atomics, no zero-init;
//   5. the last block of a frame pads the final byte with 1s; total_bits[frame] is written.
// 0xFF stuffing and JFIF framing then run on the packed stream (launch_stuffing).
#include "b2j_entropy.cuh"
#include <cstdlib>

namespace b2j {

constexpr int kChunk = 128;                       // blocks per CTA == threads per CTA
constexpr int kLookahead = 8;                     // >= 31 bits / (4 bits per smallest block)
constexpr int kCols = kChunk + kLookahead;        // blocks held in shared memory
constexpr int kPitch = 33;                        // words per block in shared memory (32 + 1 pad)
constexpr int kPrivWords = 55;                    // >= (1728 + 7 pad bits) / 32 + 1
constexpr int kCap = 8;                           // private words kept in shared memory (256 bits); the rest of a
                                                  // (rare) larger block goes to an L2-resident scratch column
constexpr int kOvfWords = kPrivWords - kCap;
constexpr int kMaxSlots = 2048;                   // persistent CTAs that may own a scratch slot
constexpr unsigned long long kFlagA = 1ull << 62, kFlagP = 2ull << 62, kValMask = (1ull << 62) - 1;
constexpr int kSpinLimit = 1 << 22;
constexpr int kFusedSmem = (kCols * kPitch + kCap * kCols + kCols + kCols * 2 + 32 + 512 + 16) * 4;

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
    int chunks;
    EGeo e;
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

// Packs fields MSB-first into a private column of shared memory: word k of column c lives at
// col[k * kCols] (word-interleaved, so any mix of k across the lanes of a warp is conflict free).
struct PrivSink {
    uint32_t *col;                // shared column (first kCap words)
    uint32_t *ovf;                // global scratch column (words kCap..)
    uint32_t acc;                 // pending bits, left aligned
    int nb;                       // number of pending bits (< 32)
    int k;                        // words already stored
    __device__ __forceinline__ void begin(uint32_t *c, uint32_t *o) { col = c; ovf = o; acc = 0u; nb = 0; k = 0; }
    __device__ __forceinline__ void store(uint32_t w)
    {
        if (k < kCap) col[k * kCols] = w;
        else ovf[(k - kCap) * kCols] = w;
    }
    __device__ __forceinline__ uint32_t load(int i) const { return i < kCap ? col[i * kCols] : ovf[(i - kCap) * kCols]; }
    __device__ __forceinline__ void put(uint32_t val, int len)      // 1 <= len <= 32, val < 2^len
    {
        const uint32_t x = val << (32 - len);                        // left aligned field
        acc |= x >> nb;
        const int n = nb + len;
        if (n >= 32) {
            store(acc);
            k++;
            acc = nb ? x << (32 - nb) : 0u;
            nb = n - 32;
        } else {
            nb = n;
        }
    }
    __device__ __forceinline__ int bits() const { return k * 32 + nb; }
    __device__ __forceinline__ void flush() { store(acc); }         // zero padded partial word
};

// Sparse walk of one block: only the non-zero AC positions (set bits of the mask) are visited.
// blk: the block's 32 words (two coefficients per word) in shared memory.  With PARTIAL the walk
// stops once 32 bits are out (lookahead blocks only have to complete the previous word).
template <bool EAGER, bool PARTIAL>
__device__ __forceinline__ void encode_block(uint32_t mlo, uint32_t mhi, int diff, const uint32_t *blk,
                                             const uint32_t *__restrict__ lut_dc, const uint32_t *__restrict__ lut_ac,
                                             PrivSink &sink)
{
    {
        const int s = bit_size(diff);
        const uint32_t e = lut_dc[s];
        const uint32_t extra = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u);
        sink.put(((e & 0xFFFFu) << s) | extra, (int)(e >> 16) + s);
    }
    const uint32_t zrl = lut_ac[0xF0], eob = lut_ac[0x00];
    const int16_t *c16 = reinterpret_cast<const int16_t *>(blk);
    int prev = 0;
    uint32_t m = mlo & ~1u;                                         // position 0 is the DC
    int base = 0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        while (m) {
            if (PARTIAL && sink.k > 0) return;
            const int z = base + __ffs((int)m) - 1;
            m &= m - 1u;
            const int v = (int)c16[z];
            const int run = z - prev - 1;
            prev = z;
            if (run >= 16) {                                        // ZRLs in front of a non-zero: same for both policies
#pragma unroll 1
                for (int n = run >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));
            }
            const int s = bit_size(v);
            const uint32_t e = lut_ac[((run & 15) << 4) | s];
            sink.put(((e & 0xFFFFu) << s) | ((uint32_t)(v + (v >> 31)) & ((1u << s) - 1u)), (int)(e >> 16) + s);
        }
        m = mhi;
        base = 32;
    }
    if (PARTIAL && sink.k > 0) return;
    const int t = 63 - prev;                                        // trailing zeros
    if (EAGER) {                                                    // rlecore.py:101-114
#pragma unroll 1
        for (int n = t >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));
        if (t & 15) sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    } else if (t > 0) {
        sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    }
}

template <bool EAGER>
__global__ void __launch_bounds__(kChunk) k_entropy_fused(const __grid_constant__ FusedEntropyParams P,
                                                          const __grid_constant__ LutParam L)
{
    extern __shared__ __align__(16) uint32_t fsm[];
    uint32_t *s_coef = fsm;                                  // [kCols][kPitch]
    uint32_t *s_priv = s_coef + kCols * kPitch;              // [kCap][kCols]
    uint32_t *s_nbits = s_priv + kCap * kCols;               // [kCols]
    uint32_t *s_mask = s_nbits + kCols;                      // [kCols][2]: eight mask bytes per block
    uint32_t *s_dc = s_mask + kCols * 2;                     // [2][16]
    uint32_t *s_ac = s_dc + 32;                              // [2][256]
    uint32_t *s_misc = s_ac + 512;                           // warp sums [4], ticket, prefix lo/hi
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *my_ovf = P.overflow + (size_t)blockIdx.x * kOvfWords * kCols;
    const unsigned int total_chunks = (unsigned)P.chunks * (unsigned)P.e.n_frames;
    load_lut(L, s_dc, s_ac);
  for (;;) {                                                          // persistent: one chunk per ticket
    __syncthreads();                                                  // shared memory of the previous chunk is free
    if (tid == 0) s_misc[4] = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const unsigned int id = s_misc[4];
    if (id >= total_chunks) break;
    // tickets interleave the frames: consecutive CTAs work on different frames, so the predecessor
    // chunk of the same frame was started n_frames tickets earlier and has usually published its
    // inclusive prefix by the time it is needed (short look-back, no serial chain inside a wave)
    const int f = (int)(id % (unsigned)P.e.n_frames), ch = (int)(id / (unsigned)P.e.n_frames);
    const EGeo &e = P.e;
    const long long b0 = (long long)ch * kChunk;
    const long long b1 = min(b0 + kChunk, e.nblk);
    const long long bl = min(b1 + kLookahead, e.nblk);              // end of the lookahead blocks
    const int nown = (int)(b1 - b0), nall = (int)(bl - b0);
    const int16_t *fc = P.coefs + (long long)f * e.nblk * 64;
    const bool last = b1 >= e.nblk;

    // ---- 1. coalesced load (16-byte pieces) into the padded layout + non-zero byte masks
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(fc + b0 * 64);
        uint8_t *mask8 = reinterpret_cast<uint8_t *>(s_mask);
        const int npieces = nall * 8;
        for (int i = tid; i < npieces; i += kChunk) {
            const int b = i >> 3, c = i & 7;
            const uint4 q = __ldg(src + i);
            uint32_t *d = s_coef + b * kPitch + c * 4;
            d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            uint32_t m = 0u;
            if (q.x & 0xFFFFu) m |= 1u;
            if (q.x >> 16) m |= 2u;
            if (q.y & 0xFFFFu) m |= 4u;
            if (q.y >> 16) m |= 8u;
            if (q.z & 0xFFFFu) m |= 16u;
            if (q.z >> 16) m |= 32u;
            if (q.w & 0xFFFFu) m |= 64u;
            if (q.w >> 16) m |= 128u;
            mask8[b * 8 + c] = (uint8_t)m;
        }
    }
    __syncthreads();

    // ---- 2. one walk per block: private packed bits + bit count
    auto dc_of = [&](long long b) -> int {          // DC of block b of this frame (b < bl)
        if (b >= b0) return (int)(int16_t)(s_coef[(int)(b - b0) * kPitch] & 0xFFFFu);
        return b >= 0 ? (int)fc[b * 64] : 0;
    };
    auto prev_index = [&](long long b) -> long long {
        const int k = (int)(b % e.bpm);
        if (k < e.hv) return k > 0 ? b - 1 : b - e.bpm + e.hv - 1;
        return b - e.bpm;
    };
    PrivSink sink;
    sink.begin(s_priv + tid, my_ovf + tid);
    int mybits = 0;
    if (tid < nown) {
        const long long b = b0 + tid;
        const uint32_t *blk = s_coef + tid * kPitch;
        const int chroma = (int)(b % e.bpm) >= e.hv;
        const int diff = (int)(int16_t)(blk[0] & 0xFFFFu) - dc_of(prev_index(b));
        encode_block<EAGER, false>(s_mask[tid * 2], s_mask[tid * 2 + 1], diff, blk, s_dc + 16 * chroma, s_ac + 256 * chroma, sink);
        sink.flush();
        mybits = sink.bits();
        s_nbits[tid] = (uint32_t)mybits;
    }
    // the lookahead blocks only have to supply the bits that complete this chunk's last word
    if (tid < kLookahead && kChunk + tid < nall && !last) {
        const int colidx = kChunk + tid;
        const long long b = b0 + colidx;
        const uint32_t *blk = s_coef + colidx * kPitch;
        const int chroma = (int)(b % e.bpm) >= e.hv;
        const int diff = (int)(int16_t)(blk[0] & 0xFFFFu) - dc_of(prev_index(b));
        PrivSink sk;
        sk.begin(s_priv + colidx, nullptr);
        encode_block<EAGER, true>(s_mask[colidx * 2], s_mask[colidx * 2 + 1], diff, blk, s_dc + 16 * chroma, s_ac + 256 * chroma, sk);
        sk.flush();
        s_nbits[colidx] = (uint32_t)sk.bits();                      // >= 32 when the walk was cut short
    }

    // ---- 3. CTA exclusive scan of the bit counts
    const uint32_t inc = warp_incl_scan((uint32_t)mybits);
    if (lane == 31) s_misc[warp] = inc;
    __syncthreads();
    uint32_t wbase_bits = 0, T = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t x = s_misc[w];
        if (w < warp) wbase_bits += x;
        T += x;
    }
    const uint32_t excl = wbase_bits + inc - (uint32_t)mybits;

    // ---- decoupled look-back over the chunk totals of this frame (warp 0)
    if (warp == 0) {
        unsigned long long *st = P.status + (long long)f * P.chunks;
        unsigned long long E = 0;
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
            s_misc[5] = (uint32_t)E;
            s_misc[6] = (uint32_t)(E >> 32);
        }
    }
    __syncthreads();
    const unsigned long long E = (unsigned long long)s_misc[5] | ((unsigned long long)s_misc[6] << 32);

    // ---- 4. compose the owned words.  Bit g of the frame's stream lives in word g >> 5.
    // (bit offsets inside a frame fit 32 bits: make_geo() refuses larger frames)
    // The frame's final byte is padded with 1s by whichever thread composes the frame's final word:
    // that is not always the last block's thread, nor even a thread of the last chunk (a short
    // tail of 4-6 bit blocks can sit entirely inside a word owned by the previous chunk, which
    // sees those blocks as its lookahead columns).
    const uint32_t total = (uint32_t)mybits;
    const bool reaches_end = bl >= e.nblk;                           // columns 0..nall-1 run to the frame's end
    if (last && tid == nown - 1) P.total_bits[f] = (int64_t)((uint32_t)E + excl + total);
    if (tid < nown) {
        const uint32_t G = (uint32_t)E + excl;
        uint32_t *dst = reinterpret_cast<uint32_t *>(P.unstuffed + (long long)f * P.ucap);
        const uint32_t wcap = (uint32_t)(P.ucap >> 2);
        uint32_t *ffp = P.ffpart + (long long)f * P.schunks;
        const int nsucc = last ? nown : nall;                        // columns that may supply following bits
        uint32_t w = (G + 31u) >> 5;
        const uint32_t wend = (G + total + 31u) >> 5;
        int rel = (int)((w << 5) - G);                               // first bit of word w inside my block
        uint32_t a = sink.load(rel >> 5);
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
                for (int nb = tid + 1; have < 32 && nb < nsucc; nb++) {
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
                if ((z - 0x01010101u) & ~z & 0x80808080u) {
                    const uint32_t n = ((z & 0xFFu) == 0u) + ((z & 0xFF00u) == 0u) + ((z & 0xFF0000u) == 0u) + ((z >> 24) == 0u);
                    atomicAdd(ffp + (w >> 10), n);                   // 1024 words per stuffing chunk
                }
            }
        }
    }
  }
}

}  // namespace b2j

using namespace b2j;

static size_t fused_status_bytes(const Geo &g)
{
    const size_t chunks = (size_t)((g.blocks_per_frame + kChunk - 1) / kChunk);
    return 256 + (size_t)g.n_frames * chunks * sizeof(unsigned long long);
}

struct FusedWs { size_t status, ffpart, zero_bytes, total_bits, unstuffed, overflow, total; };

static FusedWs fused_ws(const Geo &g, long long ucap)
{
    FusedWs w;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.status = 256;                                                     // [0,256): ticket, error flag
    w.ffpart = up(w.status + fused_status_bytes(g) - 256);
    w.zero_bytes = up(w.ffpart + (size_t)g.n_frames * stuff_chunks(ucap) * sizeof(uint32_t));   // [0, zero_bytes) is cleared per call
    w.total_bits = w.zero_bytes;
    w.unstuffed = up(w.total_bits + (size_t)g.n_frames * sizeof(int64_t));
    w.overflow = up(w.unstuffed + (size_t)g.n_frames * (size_t)ucap);
    w.total = w.overflow + (size_t)kMaxSlots * kOvfWords * kCols * sizeof(uint32_t);
    return w;
}

extern "C" size_t b2j_entropy_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap)
{
    Geo g;
    if (make_geo(desc, &g) || out_cap < kHeaderBytes + 2) return 0;
    return fused_ws(g, ucap_of(out_cap)).total;
}

// K2..K4 in one pass over the coefficients: coefs -> JFIF files.
extern "C" int b2j_entropy_encode(const int16_t *coefs, const b2j_image_desc *desc, const uint8_t qy_host[64],
                                  const uint8_t qc_host[64], int zrl, void *workspace, uint8_t *out, int64_t out_cap,
                                  int64_t *sizes, void *stream)
{
    B2J_REQUIRE(coefs && qy_host && qc_host && workspace && out && sizes, B2J_E_ARG, "b2j_entropy_encode: null pointer");
    B2J_REQUIRE(zrl == B2J_ZRL_EAGER || zrl == B2J_ZRL_STANDARD, B2J_E_ARG, "b2j_entropy_encode: bad zrl policy %d", zrl);
    B2J_REQUIRE(out_cap >= kHeaderBytes + 2, B2J_E_CAPACITY, "b2j_entropy_encode: out_cap %lld smaller than the header",
                (long long)out_cap);
    B2J_REQUIRE(((uintptr_t)coefs & 15) == 0 && ((uintptr_t)workspace & 255) == 0, B2J_E_ALIGN,
                "b2j_entropy_encode: coefs must be 16-byte and workspace 256-byte aligned");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    const long long ucap = ucap_of(out_cap);
    const int chunks = (int)((g.blocks_per_frame + kChunk - 1) / kChunk);
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    const FusedWs W = fused_ws(g, ucap);
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
    P.e = egeo_of(g);
    HeaderParam H;
    memset(&H, 0, sizeof H);
    if (jfif_header(g, qy_host, qc_host, H.bytes) != kHeaderBytes) {
        set_error("b2j_entropy_encode: internal header length mismatch");
        return B2J_E_ARG;
    }
    const LutParam L = make_lut_param();
    cudaStream_t st = as_stream(stream);
    B2J_CUDA(cudaMemsetAsync(ws, 0, status_bytes, st));              // ticket, error flag, chunk status words
    const long long total = (long long)chunks * g.n_frames;
    int dev = 0, sms = 148, per_sm = 1;
    B2J_CUDA(cudaGetDevice(&dev));
    B2J_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (zrl == B2J_ZRL_EAGER) {
        B2J_CUDA(cudaFuncSetAttribute(k_entropy_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
        B2J_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_entropy_fused<true>, kChunk, kFusedSmem));
    } else {
        B2J_CUDA(cudaFuncSetAttribute(k_entropy_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
        B2J_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_entropy_fused<false>, kChunk, kFusedSmem));
    }
    if (const char *cap = getenv("B2J_ENT_CTAS_PER_SM")) {              // tuning knob: leave room for a concurrent kernel
        const int c = atoi(cap);
        if (c >= 1 && c < per_sm) per_sm = c;
    }
    long long slots = (long long)sms * (per_sm < 1 ? 1 : per_sm);
    if (slots > kMaxSlots) slots = kMaxSlots;
    const unsigned grid = (unsigned)(total < slots ? total : slots);   // persistent CTAs, one scratch slot each
    if (zrl == B2J_ZRL_EAGER) k_entropy_fused<true><<<grid, kChunk, kFusedSmem, st>>>(P, L);
    else k_entropy_fused<false><<<grid, kChunk, kFusedSmem, st>>>(P, L);
    B2J_LAUNCH_CHECK("k_entropy_fused");
    return launch_stuffing(P.unstuffed, P.total_bits, ffpart, sizes, out, ucap, out_cap, g.n_frames, H, P.error_flag, true, st);
}
