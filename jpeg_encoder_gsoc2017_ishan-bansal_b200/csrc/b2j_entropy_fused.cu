// Fused single-pass entropy coder (K2 + K3 + K4 packing in one kernel).
//
// One CTA per chunk of 128 consecutive blocks of a frame, in launch order (ticket):
//   1. the chunk's coefficients (+ 8 lookahead blocks) are loaded once, coalesced, into
//      shared memory (16-byte pieces, XOR-swizzled so that the per-thread reads are
//      conflict free);
//   2. every thread takes one block: 64-bit non-zero mask, DC difference, and a SPARSE walk
//      over the set bits only (symbols as in b2j_entropy.cuh: gateware run-length rules,
//      Annex-K code lengths) that counts the block's bits;
//   3. warp-shuffle + CTA exclusive scan of the bit counts; the chunk total is published and
//      the chunk's bit offset inside the frame is obtained by a decoupled look-back over the
//      chunk totals of its predecessors (one 64-bit status word per chunk: flag | value);
//   4. a second sparse walk emits the bits at their final positions into a shared-memory
//      window (32-bit words, MSB first); the chunk owns every 32-bit word of the stream whose
//      first bit it contains and completes its last owned word by also emitting the first
//      bits of the next few blocks (lookahead), so all global stores are plain, coalesced,
//      full-word stores - no global atomics, no zero-initialised output;
//   5. the last chunk of a frame pads the final byte with 1s and writes total_bits[frame].
// 0xFF stuffing and JFIF framing then run on the packed stream (launch_stuffing).
#include "b2j_entropy.cuh"

namespace b2j {

constexpr int kChunk = 128;                       // blocks per CTA == threads per CTA
constexpr int kLookahead = 8;                     // >= 31 bits / (4 bits per smallest block)
constexpr int kWinWords = 2048;                   // staging window (65536 bits)
constexpr int kSlackWords = 64;                   // >= one maximal block (1728 bits)
constexpr unsigned long long kFlagA = 1ull << 62, kFlagP = 2ull << 62, kValMask = (1ull << 62) - 1;
constexpr int kSpinLimit = 1 << 22;

struct FusedEntropyParams {
    const int16_t *coefs;
    uint8_t *unstuffed;
    int64_t *total_bits;
    unsigned long long *status;   // [n_frames * chunks]
    unsigned int *ticket;
    int *error_flag;
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

// coefficient z (zig-zag position) of a block stored as 8 swizzled 16-byte pieces
__device__ __forceinline__ int coef_at(const uint8_t *blk, int swz, int z)
{
    return (int)*reinterpret_cast<const int16_t *>(blk + ((((z >> 3) ^ swz) & 7) << 4) + ((z & 7) << 1));
}

__device__ __forceinline__ void block_mask(const uint8_t *blk, int swz, uint32_t &mlo, uint32_t &mhi)
{
    mlo = 0u; mhi = 0u;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 q = *reinterpret_cast<const uint4 *>(blk + (((c ^ swz) & 7) << 4));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t m = 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (w[j] & 0xFFFFu) m |= 1u << (2 * j);
            if (w[j] >> 16) m |= 2u << (2 * j);
        }
        if (c < 4) mlo |= m << (8 * c);
        else mhi |= m << (8 * (c - 4));
    }
}

// Sparse walk: only the non-zero AC positions (bits of the mask) are visited.
template <bool EAGER, class Sink>
__device__ __forceinline__ void walk_sparse(uint32_t mlo, uint32_t mhi, int diff, const uint8_t *blk, int swz,
                                            const uint32_t *__restrict__ lut_dc, const uint32_t *__restrict__ lut_ac, Sink &sink)
{
    {
        const int s = bit_size(diff);
        const uint32_t e = lut_dc[s];
        const uint32_t extra = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u);
        sink.put(((e & 0xFFFFu) << s) | extra, (int)(e >> 16) + s);
    }
    const uint32_t zrl = lut_ac[0xF0], eob = lut_ac[0x00];
    int prev = 0;
    uint32_t m = mlo & ~1u;                         // position 0 is the DC
    int base = 0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        while (m) {
            const int z = base + __ffs((int)m) - 1;
            m &= m - 1u;
            const int v = coef_at(blk, swz, z);
            const int run = z - prev - 1;
            prev = z;
            for (int n = run >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));   // same for both ZRL policies
            const int s = bit_size(v);
            const uint32_t e = lut_ac[((run & 15) << 4) | s];
            sink.put(((e & 0xFFFFu) << s) | ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << s) - 1u)), (int)(e >> 16) + s);
        }
        m = mhi;
        base = 32;
    }
    const int t = 63 - prev;                        // trailing zeros
    if (EAGER) {                                    // rlecore.py:101-114
        for (int n = t >> 4; n > 0; n--) sink.put(zrl & 0xFFFFu, (int)(zrl >> 16));
        if (t & 15) sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    } else if (t > 0) {
        sink.put(eob & 0xFFFFu, (int)(eob >> 16));
    }
}

struct CountSink2 {
    int bits;
    __device__ __forceinline__ void put(uint32_t, int len) { bits += len; }
};

// Emits into the shared staging window; words at or beyond `limit` are dropped (lookahead
// blocks only contribute the bits that complete the chunk's last owned word).
struct StageSink {
    uint32_t *stage;
    unsigned long long acc;
    int nbits, widx, limit;
    __device__ __forceinline__ void begin(uint32_t *s, uint32_t bitpos, int lim)
    {
        stage = s; acc = 0ull; nbits = (int)(bitpos & 31u); widx = (int)(bitpos >> 5); limit = lim;
    }
    __device__ __forceinline__ void put(uint32_t val, int len)
    {
        acc = (acc << len) | val;
        nbits += len;
        if (nbits >= 32) {
            if (widx < limit) atomicOr(stage + widx, (uint32_t)(acc >> (nbits - 32)));
            widx++;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nbits > 0 && widx < limit) atomicOr(stage + widx, (uint32_t)(acc << (32 - nbits)));
    }
};

template <bool EAGER>
__global__ void __launch_bounds__(kChunk) k_entropy_fused(const __grid_constant__ FusedEntropyParams P,
                                                          const __grid_constant__ LutParam L)
{
    __shared__ __align__(16) uint8_t s_coef[(kChunk + kLookahead) * 128];
    __shared__ uint32_t s_stage[kWinWords + kSlackWords];
    __shared__ uint32_t s_dc[32], s_ac[512];
    __shared__ uint32_t s_warp[4];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_id;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_id = atomicAdd(P.ticket, 1u);
    load_lut(L, s_dc, s_ac);
    __syncthreads();
    const unsigned int id = s_id;
    const int f = (int)(id / (unsigned)P.chunks), ch = (int)(id % (unsigned)P.chunks);
    const EGeo &e = P.e;
    const long long b0 = (long long)ch * kChunk;
    const long long b1 = min(b0 + kChunk, e.nblk);
    const long long bl = min(b1 + kLookahead, e.nblk);              // end of the lookahead blocks
    const int16_t *fc = P.coefs + (long long)f * e.nblk * 64;
    const bool last = b1 >= e.nblk;

    // ---- 1. coalesced load of the chunk (+ lookahead) into swizzled shared memory
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(fc + b0 * 64);
        const int npieces = (int)(bl - b0) * 8;
        for (int i = tid; i < npieces; i += kChunk) {
            const int b = i >> 3, c = i & 7;
            *reinterpret_cast<uint4 *>(s_coef + b * 128 + (((c ^ b) & 7) << 4)) = __ldg(src + i);
        }
    }
    __syncthreads();

    // ---- 2. per block: mask, DC difference, bit count
    auto dc_of = [&](long long b) -> int {          // DC of block b of this frame (b < bl)
        if (b >= b0) return coef_at(s_coef + (int)(b - b0) * 128, (int)(b - b0) & 7, 0);
        return b >= 0 ? (int)fc[b * 64] : 0;
    };
    auto prev_index = [&](long long b) -> long long {
        const int k = (int)(b % e.bpm);
        if (k < e.hv) return k > 0 ? b - 1 : b - e.bpm + e.hv - 1;
        return b - e.bpm;
    };
    const long long myb = b0 + tid;
    const bool own = myb < b1;
    uint32_t mlo = 0u, mhi = 0u;
    int diff = 0, chroma = 0, mybits = 0;
    const uint8_t *myblk = s_coef + tid * 128;
    if (own) {
        block_mask(myblk, tid & 7, mlo, mhi);
        chroma = (int)(myb % e.bpm) >= e.hv;
        diff = coef_at(myblk, tid & 7, 0) - dc_of(prev_index(myb));
        CountSink2 cs{0};
        walk_sparse<EAGER>(mlo, mhi, diff, myblk, tid & 7, s_dc + 16 * chroma, s_ac + 256 * chroma, cs);
        mybits = cs.bits;
    }

    // ---- 3. CTA exclusive scan of the bit counts
    uint32_t inc = warp_incl_scan((uint32_t)mybits);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase_bits = 0, T = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t x = s_warp[w];
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
            s_prefix = E;
        }
    }
    __syncthreads();
    const unsigned long long E = s_prefix;                           // bits of the frame before this chunk

    // ---- 4. emit.  Global bit g of the frame lives in word g >> 5 of the frame's stream.
    const unsigned long long Eend = E + T;
    const unsigned long long Epad = last ? ((Eend + 7ull) & ~7ull) : Eend;
    const unsigned long long W0 = (E + 31ull) >> 5, W1 = (Epad + 31ull) >> 5;     // owned words [W0, W1)
    const unsigned long long B0 = (E >> 5) << 5;                                   // window origin (bit)
    const unsigned long long myG = E + excl;                                       // first bit of my block
    constexpr unsigned long long WB = (unsigned long long)kWinWords * 32ull;
    const int n_rounds = (int)((Epad - B0 + WB - 1ull) / WB) > 0 ? (int)((Epad - B0 + WB - 1ull) / WB) : 1;
    uint32_t *dst = reinterpret_cast<uint32_t *>(P.unstuffed + (long long)f * P.ucap);
    const unsigned long long wcap = (unsigned long long)(P.ucap >> 2);

    for (int r = 0; r < n_rounds; r++) {
        const unsigned long long wb = B0 + (unsigned long long)r * WB;            // first bit of this window
        if (r == 0) {
            for (int i = tid; i < kWinWords + kSlackWords; i += kChunk) s_stage[i] = 0u;
        } else {
            // bits that spilled past the previous window are the first words of this one
            __syncthreads();
            const uint32_t carry = tid < kSlackWords ? s_stage[kWinWords + tid] : 0u;
            __syncthreads();
            for (int i = tid; i < kWinWords + kSlackWords; i += kChunk) s_stage[i] = 0u;
            __syncthreads();
            if (tid < kSlackWords) s_stage[tid] = carry;
        }
        __syncthreads();
        const bool final_round = r == n_rounds - 1;
        if (own && myG >= wb && myG < wb + WB) {
            StageSink sink;
            sink.begin(s_stage, (uint32_t)(myG - wb), kWinWords + kSlackWords);
            walk_sparse<EAGER>(mlo, mhi, diff, myblk, tid & 7, s_dc + 16 * chroma, s_ac + 256 * chroma, sink);
            sink.finish();
        }
        if (final_round) {
            // lookahead: the first bits of the following blocks complete the last owned word
            const unsigned long long lim_bit = W1 << 5;
            if (!last && warp == 0) {
                // bit offsets of the lookahead blocks: serial prefix over at most 8 counts
                const long long lb = b1 + lane;
                const bool act = lane < kLookahead && lb < bl;
                uint32_t lmlo = 0u, lmhi = 0u;
                int ldiff = 0, lchroma = 0, lbits = 0;
                const uint8_t *lblk = s_coef + (int)(lb - b0) * 128;
                const int lswz = (int)(lb - b0) & 7;
                if (act) {
                    block_mask(lblk, lswz, lmlo, lmhi);
                    lchroma = (int)(lb % e.bpm) >= e.hv;
                    ldiff = coef_at(lblk, lswz, 0) - dc_of(prev_index(lb));
                    CountSink2 cs{0};
                    walk_sparse<EAGER>(lmlo, lmhi, ldiff, lblk, lswz, s_dc + 16 * lchroma, s_ac + 256 * lchroma, cs);
                    lbits = cs.bits;
                }
                const uint32_t linc = warp_incl_scan((uint32_t)lbits);
                const unsigned long long lG = Eend + (linc - (uint32_t)lbits);
                if (act && lG < lim_bit) {
                    StageSink sink;
                    sink.begin(s_stage, (uint32_t)(lG - wb), (int)((lim_bit - wb) >> 5));
                    walk_sparse<EAGER>(lmlo, lmhi, ldiff, lblk, lswz, s_dc + 16 * lchroma, s_ac + 256 * lchroma, sink);
                    sink.finish();
                }
            }
            if (last && tid == 0 && Epad > Eend) {                   // pad the final byte with 1s
                const uint32_t pos = (uint32_t)(Eend - wb), n = (uint32_t)(Epad - Eend);
                atomicOr(s_stage + (pos >> 5), ((1u << n) - 1u) << (32u - (pos & 31u) - n));
            }
        }
        __syncthreads();
        // store the owned words that are complete in this window (big-endian bytes)
        const unsigned long long wlo = wb >> 5;
        const unsigned long long whi = final_round ? W1 : wlo + kWinWords;
        const unsigned long long lo = W0 > wlo ? W0 : wlo, hi = W1 < whi ? W1 : whi;
        for (unsigned long long w = lo + tid; w < hi; w += kChunk)
            if (w < wcap) dst[w] = __byte_perm(s_stage[(int)(w - wlo)], 0u, 0x0123);
    }
    if (last && tid == 0) P.total_bits[f] = (int64_t)Eend;
}

}  // namespace b2j

using namespace b2j;

static size_t fused_status_bytes(const Geo &g)
{
    const size_t chunks = (size_t)((g.blocks_per_frame + kChunk - 1) / kChunk);
    return 256 + (size_t)g.n_frames * chunks * sizeof(unsigned long long);
}

extern "C" size_t b2j_entropy_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap)
{
    Geo g;
    if (make_geo(desc, &g) || out_cap < kHeaderBytes + 2) return 0;
    const long long ucap = ucap_of(out_cap);
    size_t o = (fused_status_bytes(g) + 255) & ~(size_t)255;
    o += ((size_t)g.n_frames * sizeof(int64_t) + 255) & ~(size_t)255;
    o += (size_t)g.n_frames * (size_t)ucap;
    o += (size_t)g.n_frames * stuff_chunks(ucap) * sizeof(uint32_t) + 512;
    return o;
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
    const size_t status_bytes = fused_status_bytes(g);
    FusedEntropyParams P;
    P.coefs = coefs;
    P.ticket = reinterpret_cast<unsigned int *>(ws);
    P.error_flag = reinterpret_cast<int *>(ws + 64);
    P.status = reinterpret_cast<unsigned long long *>(ws + 256);
    size_t o = (status_bytes + 255) & ~(size_t)255;
    P.total_bits = reinterpret_cast<int64_t *>(ws + o);
    o += ((size_t)g.n_frames * sizeof(int64_t) + 255) & ~(size_t)255;
    P.unstuffed = ws + o;
    o += (size_t)g.n_frames * (size_t)ucap;
    uint32_t *ffpart = reinterpret_cast<uint32_t *>(ws + o);
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
    const unsigned grid = (unsigned)((long long)chunks * g.n_frames);
    if (zrl == B2J_ZRL_EAGER) k_entropy_fused<true><<<grid, kChunk, 0, st>>>(P, L);
    else k_entropy_fused<false><<<grid, kChunk, 0, st>>>(P, L);
    B2J_LAUNCH_CHECK("k_entropy_fused");
    return launch_stuffing(P.unstuffed, P.total_bits, ffpart, sizes, out, ucap, out_cap, g.n_frames, H, P.error_flag, st);
}
