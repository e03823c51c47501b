// K1: fused colour conversion -> chroma subsampling -> 8x8 DCT (the model's
// float64 graph) -> quantise -> zig-zag, one pass over the RGB frames.
// replaces: the loop body test/model/enc_frame.py:203-217 of the reference.
//
// Output layout: coefs[frame][mcu][block_in_mcu][64] int16, zig-zag order, MCUs
// row-major, blocks of an MCU in scan order (luma raster, then Cb, then Cr).
//
// FIXED selects the [ext] gateware fixed-point DCT instead of the float64 model (b2j_block.cuh).
// Two variants:
//   simple - one thread per 8x8 block straight from global memory; any geometry;
//            kept as the cross-check and the un-tiled fallback.
//   tiled  - persistent CTAs walk tiles of 1024 x (8|16) pixels.  Per tile:
//            (load)  the RGB rows are brought into shared memory by bulk-async
//                    copies (cp.async.bulk, one per row, completion on an
//                    mbarrier) when rows are 16-byte aligned, else by a
//                    cooperative copy;
//            (A)     every thread converts items of 4 pixels x V rows to Y/Cb/Cr with the
//                    integer restatement of the model's colour formulas (the inputs
//                    where the float64 model differs are patched from tables) and
//                    writes 8-bit planes (chroma already subsampled) to smem;
//            (B)     every thread takes whole 8x8 blocks from the planes, runs
//                    the model's DCT graph in registers (FP64 pipe), quantises
//                    and stores 128 B per block.
//            The next tile's load is issued as soon as (A) has consumed the raw
//            buffer, so it overlaps (B), which is the long FP64 phase.
#include "b2j_common.cuh"
#include "b2j_math.cuh"
#include "b2j_block.cuh"

namespace b2j {

struct FusedParams {
    const uint8_t *rgb;
    int16_t *coefs;
    long long row_stride, frame_stride;
    int width, height;
    int mcus_x, mcus_y;
    int n_frames;
    int aligned;                  // rows may use bulk-async copies
    int bpp, ro;                  // [ext] source layout: bytes per pixel, byte of r in a pixel (b at 2 - ro); pair and simple variants
    QuantMul qm;
    QuantF32 qf;                  // pair variant only
};

// =============================================================== simple variant
template <int H, int V, bool FIXED>
__global__ void __launch_bounds__(128) k_fused_simple(const __grid_constant__ FusedParams P)
{
    pdl_launch_dependents();
    constexpr int BPM = H * V + 2;
    __shared__ uint32_t s_mul[2][64];
    if (threadIdx.x < 128) s_mul[threadIdx.x >> 6][threadIdx.x & 63] = P.qm.mul[threadIdx.x >> 6][threadIdx.x & 63];
    __syncthreads();
    const long long nblk = (long long)P.n_frames * P.mcus_x * P.mcus_y * BPM;
    const long long gb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gb >= nblk) return;
    const int k = (int)(gb % BPM);
    const long long gm = gb / BPM;
    const int mx = (int)(gm % P.mcus_x);
    const int my = (int)((gm / P.mcus_x) % P.mcus_y);
    const int f = (int)(gm / ((long long)P.mcus_x * P.mcus_y));
    const uint8_t *img = P.rgb + (long long)f * P.frame_stride;
    const int comp = k < H * V ? 0 : k - H * V + 1;
    uint2 px[8];
    for (int i = 0; i < 8; i++) {
        uint32_t wds[2] = {0u, 0u};
        for (int j = 0; j < 8; j++) {
            uint32_t s;
            if (comp == 0) {
                int yy = my * 8 * V + (k / H) * 8 + i, xx = mx * 8 * H + (k % H) * 8 + j;
                yy = min(yy, P.height - 1); xx = min(xx, P.width - 1);          // edge replication
                const uint8_t *q = img + (long long)yy * P.row_stride + (long long)P.bpp * xx;
                s = rgb2ycbcr_px(q[P.ro], q[1], q[2 - P.ro]) & 0xFFu;
            } else {
                int acc_v = 0;
                for (int dy = 0; dy < V; dy++) {
                    int acc_h = 0;
                    for (int dx = 0; dx < H; dx++) {
                        int yy = min(my * 8 * V + i * V + dy, P.height - 1);
                        int xx = min(mx * 8 * H + j * H + dx, P.width - 1);
                        const uint8_t *q = img + (long long)yy * P.row_stride + (long long)P.bpp * xx;
                        acc_h += (int)((rgb2ycbcr_px(q[P.ro], q[1], q[2 - P.ro]) >> (8 * comp)) & 0xFFu);
                    }
                    acc_v += H == 2 ? acc_h >> 1 : acc_h;                        // crs.py:39-54 pair mean
                }
                s = (uint32_t)(V == 2 ? acc_v >> 1 : acc_v);
            }
            wds[j >> 2] |= s << (8 * (j & 3));
        }
        px[i] = make_uint2(wds[0], wds[1]);
    }
    BlockPacker pk;
    pk.init(s_mul[comp ? 1 : 0]);
    if (FIXED) dct_quant_block_fixed(px, pk);
    else dct_quant_block(px, pk);
    pk.store(P.coefs + gb * 64);
}

// =============================================================== tiled variant
// Tile width / CTA size / occupancy hint are compile-time knobs (profiles/scripts/k1_variants.py):
// measured on B200 for 4:2:0: 1024 px x 128 threads x 3 CTAs/SM is the best of the variants tried
// (512-px tiles -7 %, 192 or 256 threads -6 % / -30 %, 8-row band loads with 4 CTAs/SM -16 %).
#ifndef B2J_TILE_W
#define B2J_TILE_W 1024
#endif
#ifndef B2J_K1_THREADS
#define B2J_K1_THREADS 128
#endif
#ifndef B2J_K1_MIN_CTAS
#define B2J_K1_MIN_CTAS 2
#endif
constexpr int kTileW = B2J_TILE_W;           // pixels
constexpr int kRawPitch = kTileW * 3;        // bytes
constexpr int kThreads = B2J_K1_THREADS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int H, int V>
struct TileShape {
    static constexpr int kRows = 8 * V;                       // pixel rows per tile
    static constexpr int kCW = kTileW / H;                    // chroma plane width
    static constexpr int kRawBytes = kRows * kRawPitch;
    static constexpr int kYBytes = kRows * kTileW;
    static constexpr int kCBytes = 8 * kCW;
    static constexpr int kSmem = kRawBytes + kYBytes + 2 * kCBytes + 2 * 64 * 4 + 256 + 16;
};

struct TileCoord {
    int frame, my, x0;            // MCU row, first pixel column
    int valid_px, need_px;        // source pixels available / pixels after padding to whole MCUs
    int valid_rows;
};

template <int H, int V>
__device__ __forceinline__ TileCoord tile_coord(const FusedParams &P, long long t, int tiles_x)
{
    TileCoord c;
    c.frame = (int)(t / ((long long)tiles_x * P.mcus_y));
    int r = (int)(t % ((long long)tiles_x * P.mcus_y));
    c.my = r / tiles_x;
    c.x0 = (r % tiles_x) * kTileW;
    c.valid_px = min(kTileW, P.width - c.x0);
    c.need_px = min(kTileW, P.mcus_x * 8 * H - c.x0);
    c.valid_rows = min(8 * V, P.height - c.my * 8 * V);
    return c;
}

template <int H, int V, bool FIXED>
__global__ void __launch_bounds__(kThreads, B2J_K1_MIN_CTAS) k_fused_tiled(const __grid_constant__ FusedParams P, long long n_tiles, int tiles_x)
{
    pdl_launch_dependents();
    using TS = TileShape<H, V>;
    constexpr int BPM = H * V + 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *raw = smem;
    uint8_t *s_y = raw + TS::kRawBytes;
    uint8_t *s_cb = s_y + TS::kYBytes;
    uint8_t *s_cr = s_cb + TS::kCBytes;
    uint32_t *s_mul = reinterpret_cast<uint32_t *>(s_cr + TS::kCBytes);      // [2][64]
    uint8_t *s_crx = reinterpret_cast<uint8_t *>(s_mul + 128);               // [256]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_crx + 256);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 128; i += kThreads) s_mul[i] = P.qm.mul[i >> 6][i & 63];
    for (int i = tid; i < 256; i += kThreads) s_crx[i] = (uint8_t)cr_exception_class((uint32_t)i);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](const TileCoord &c) -> bool {
        // bulk-async path needs 16-byte aligned rows and a 16-byte multiple of bytes per row
        const bool bulk = P.aligned && ((c.valid_px * 3) & 15) == 0;
        if (bulk && tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)(c.valid_rows * c.valid_px * 3));
            const uint8_t *src = P.rgb + (long long)c.frame * P.frame_stride +
                                 (long long)(c.my * 8 * V) * P.row_stride + 3LL * c.x0;
            for (int r = 0; r < c.valid_rows; r++)
                bulk_g2s(raw + r * kRawPitch, src + (long long)r * P.row_stride, (uint32_t)(c.valid_px * 3), bar);
        }
        return bulk;
    };

    long long t = blockIdx.x;
    if (t >= n_tiles) return;
    TileCoord cur = tile_coord<H, V>(P, t, tiles_x);
    bool cur_bulk = issue_load(cur);
    uint32_t phase = 0;

    for (; t < n_tiles; t += gridDim.x) {
        // ---------------- wait for / perform the load of the current tile
        if (cur_bulk) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            const uint8_t *src = P.rgb + (long long)cur.frame * P.frame_stride +
                                 (long long)(cur.my * 8 * V) * P.row_stride + 3LL * cur.x0;
            const int nbytes = cur.valid_px * 3;
            for (int r = 0; r < cur.valid_rows; r++)
                for (int i = tid; i < nbytes; i += kThreads) raw[r * kRawPitch + i] = src[(long long)r * P.row_stride + i];
            __syncthreads();
        }
        // ---------------- edge replication up to whole MCUs (uniform branch, edge tiles only)
        if (cur.valid_px < cur.need_px || cur.valid_rows < TS::kRows) {
            const int pad = (cur.need_px - cur.valid_px) * 3;
            for (int i = tid; i < cur.valid_rows * pad; i += kThreads) {
                const int r = i / pad, j = i % pad;
                raw[r * kRawPitch + cur.valid_px * 3 + j] = raw[r * kRawPitch + (cur.valid_px - 1) * 3 + j % 3];
            }
            __syncthreads();
            const int nb = cur.need_px * 3;
            for (int r = cur.valid_rows; r < TS::kRows; r++)
                for (int i = tid; i < nb; i += kThreads) raw[r * kRawPitch + i] = raw[(cur.valid_rows - 1) * kRawPitch + i];
            __syncthreads();
        }

        // ---------------- phase A: colour conversion + subsampling into 8-bit planes.
        // work item = 4 consecutive pixels of one row (4:2:0: of a pair of rows); 8 x 256 items per tile.
        // The fast path is straight-line integer code for all pixels of the item (no branches, so
        // the pixels interleave); the rare inputs where the float64 model differs from the exact
        // rational value are only flagged here and patched afterwards from the raw bytes.
#pragma unroll 1            // (unrolling by 2 pushes the kernel past the instruction cache: 1.56 -> 2.0 ms)
        for (int item = tid; item < 8 * (kTileW / 4); item += kThreads) {
            const int ry = item / (kTileW / 4), grp = item % (kTileW / 4);
            if (grp * 4 >= cur.need_px) continue;
            uint32_t sy[V][4], scb[V][4], scr[V][4];
            bool y_clean = true;
            uint32_t cr_flag = 0u;
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(raw + (ry * V + v) * kRawPitch + grp * 12);
                const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                uint32_t r[4], g[4], b[4];
                r[0] = __byte_perm(w0, 0u, 0x4440); g[0] = __byte_perm(w0, 0u, 0x4441); b[0] = __byte_perm(w0, 0u, 0x4442);
                r[1] = __byte_perm(w0, 0u, 0x4443); g[1] = __byte_perm(w1, 0u, 0x4440); b[1] = __byte_perm(w1, 0u, 0x4441);
                r[2] = __byte_perm(w1, 0u, 0x4442); g[2] = __byte_perm(w1, 0u, 0x4443); b[2] = __byte_perm(w2, 0u, 0x4440);
                r[3] = __byte_perm(w2, 0u, 0x4441); g[3] = __byte_perm(w2, 0u, 0x4442); b[3] = __byte_perm(w2, 0u, 0x4443);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    sy[v][i] = csc_sum_y(r[i], g[i], b[i]);
                    scb[v][i] = csc_sum_cb(r[i], g[i], b[i]);
                    scr[v][i] = csc_sum_cr(r[i], g[i], b[i]);
                    y_clean = y_clean && (sy[v][i] & B2J_CONST_EXC_MASK) != 0u;
                    // Cr candidates (g == b, r - g even) are common; only eight grey levels can differ
                    cr_flag |= (scr[v][i] & B2J_CONST_EXC_MASK) == 0u ? (uint32_t)s_crx[g[i]] : 0u;
                }
            }
            // luma: the component is the top byte of the sum
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t lo = __byte_perm(sy[v][0], sy[v][1], 0x0073), hi = __byte_perm(sy[v][2], sy[v][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_y + (ry * V + v) * kTileW + grp * 4) = __byte_perm(lo, hi, 0x5410);
            }
            if (H == 1) {
                const uint32_t lo = __byte_perm(scb[0][0], scb[0][1], 0x0073), hi = __byte_perm(scb[0][2], scb[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_cb + ry * TS::kCW + grp * 4) = __byte_perm(lo, hi, 0x5410);
                const uint32_t lo2 = __byte_perm(scr[0][0], scr[0][1], 0x0073), hi2 = __byte_perm(scr[0][2], scr[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_cr + ry * TS::kCW + grp * 4) = __byte_perm(lo2, hi2, 0x5410);
            } else {
                // pair mean (c0 + c1) >> 1 horizontally (crs.py:39-54), then vertically
                uint32_t cb0 = ((scb[0][0] >> 24) + (scb[0][1] >> 24)) >> 1, cb1 = ((scb[0][2] >> 24) + (scb[0][3] >> 24)) >> 1;
                uint32_t cr0 = ((scr[0][0] >> 24) + (scr[0][1] >> 24)) >> 1, cr1 = ((scr[0][2] >> 24) + (scr[0][3] >> 24)) >> 1;
                if (V == 2) {
                    cb0 = (cb0 + (((scb[V - 1][0] >> 24) + (scb[V - 1][1] >> 24)) >> 1)) >> 1;
                    cb1 = (cb1 + (((scb[V - 1][2] >> 24) + (scb[V - 1][3] >> 24)) >> 1)) >> 1;
                    cr0 = (cr0 + (((scr[V - 1][0] >> 24) + (scr[V - 1][1] >> 24)) >> 1)) >> 1;
                    cr1 = (cr1 + (((scr[V - 1][2] >> 24) + (scr[V - 1][3] >> 24)) >> 1)) >> 1;
                }
                *reinterpret_cast<uint16_t *>(s_cb + ry * TS::kCW + grp * 2) = (uint16_t)(cb0 | (cb1 << 8));
                *reinterpret_cast<uint16_t *>(s_cr + ry * TS::kCW + grp * 2) = (uint16_t)(cr0 | (cr1 << 8));
            }
            // ---- patch-ups: the sums are still in registers, r and g are re-read from the raw words.
            // Rare in noisy content, but every grey pixel (r = g = b) is a luma candidate, so this is
            // branch-free per row: four independent table reads, one word store.
            if (!y_clean) {
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(raw + (ry * V + v) * kRawPitch + grp * 12);
                    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                    uint32_t rg[4];                                   // r * 256 + g of the four pixels
                    rg[0] = __byte_perm(w0, 0u, 0x4401);
                    rg[1] = __byte_perm(w0, w1, 0x3434) & 0xFFFFu;
                    rg[2] = __byte_perm(w1, 0u, 0x4423);
                    rg[3] = __byte_perm(w2, 0u, 0x4412);
                    uint32_t packed = 0u;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t bits = 0u;
                        if ((sy[v][i] & B2J_CONST_EXC_MASK) == 0u) bits = __ldg(kYExceptionBits + (rg[i] >> 5));
                        const uint32_t y = (sy[v][i] >> 24) - ((bits >> (rg[i] & 31u)) & 1u);
                        packed |= y << (8 * i);
                    }
                    *reinterpret_cast<uint32_t *>(s_y + (ry * V + v) * kTileW + grp * 4) = packed;
                }
            }
            if (cr_flag != 0u) {
                uint32_t cr[V][4];
#pragma unroll
                for (int v = 0; v < V; v++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        cr[v][i] = scr[v][i] >> 24;
                        if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) {
                            const uint8_t *px = raw + (ry * V + v) * kRawPitch + grp * 12 + i * 3;
                            cr[v][i] -= cr_exception(px[0], px[1]);
                        }
                    }
                if (H == 1) {
                    *reinterpret_cast<uint32_t *>(s_cr + ry * TS::kCW + grp * 4) =
                        cr[0][0] | (cr[0][1] << 8) | (cr[0][2] << 16) | (cr[0][3] << 24);
                } else {
                    uint32_t c0 = (cr[0][0] + cr[0][1]) >> 1, c1 = (cr[0][2] + cr[0][3]) >> 1;
                    if (V == 2) {
                        c0 = (c0 + ((cr[V - 1][0] + cr[V - 1][1]) >> 1)) >> 1;
                        c1 = (c1 + ((cr[V - 1][2] + cr[V - 1][3]) >> 1)) >> 1;
                    }
                    *reinterpret_cast<uint16_t *>(s_cr + ry * TS::kCW + grp * 2) = (uint16_t)(c0 | (c1 << 8));
                }
            }
        }
        __syncthreads();        // planes complete; raw buffer is free again

        // ---------------- prefetch the next tile while phase B runs
        const long long tn = t + gridDim.x;
        TileCoord nxt = cur;
        bool nxt_bulk = false;
        if (tn < n_tiles) {
            nxt = tile_coord<H, V>(P, tn, tiles_x);
            nxt_bulk = issue_load(nxt);
        }

        // ---------------- phase B: one 8x8 block per thread, warp-uniform component class,
        // one code path for luma and chroma (the kernel must fit the instruction cache)
        {
            const int nbx = cur.need_px >> 3;                  // luma blocks per block row
            const int ncx = nbx / H;                           // chroma blocks per component
            const long long mcu_base = ((long long)cur.frame * P.mcus_y + cur.my) * P.mcus_x + cur.x0 / (8 * H);
            constexpr int kYPer = kTileW / 8 / 32;             // tasks per luma block row
            constexpr int kYTasks = V * kYPer;
            constexpr int kCPer = kTileW / H / 8 / 32;         // tasks per chroma component
            constexpr int kTasks = kYTasks + 2 * kCPer;
#pragma unroll 1
            for (int task = warp; task < kTasks; task += kThreads / 32) {
                const bool luma = task < kYTasks;
                const int ct = task - kYTasks;
                const int by = luma ? task / kYPer : 0;
                const int comp = luma ? 0 : ct / kCPer;
                const int bx = (luma ? task % kYPer : ct % kCPer) * 32 + lane;
                const int nb = luma ? nbx : ncx;
                const uint8_t *plane = luma ? s_y + by * 8 * kTileW : (comp ? s_cr : s_cb);
                const int pitch = luma ? kTileW : TS::kCW;
                const int k = luma ? by * H + (bx % H) : H * V + comp;
                const long long mcu = mcu_base + (luma ? bx / H : bx);
                const uint32_t *mul = luma ? s_mul : s_mul + 64;
                if (bx < nb) {
                    uint2 px[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) px[i] = *reinterpret_cast<const uint2 *>(plane + i * pitch + bx * 8);
                    BlockPacker pk;
                    pk.init(mul);
                    if (FIXED) dct_quant_block_fixed(px, pk);
                    else dct_quant_block(px, pk);
                    pk.store(P.coefs + (mcu * BPM + k) * 64);
                }
            }
        }
        __syncthreads();        // planes may be overwritten by the next phase A
        cur = nxt;
        cur_bulk = nxt_bulk;
    }
}

}  // namespace b2j
#include "b2j_fused_stream.cuh"     // =========================================================== stream variant
#include "b2j_fused_pair.cuh"       // =========================================================== pair variant
namespace b2j {

// The frames as a [frame][row][byte] tensor for the stream kernel's region loads.  Returns false when the
// layout cannot be described (unaligned base or strides, driver without the entry point): the kernel then
// copies every region byte-wise.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_frame_map(const FusedParams &P, int box_bytes, int box_rows, CUtensorMap *map)
{
    static EncodeTiledFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<EncodeTiledFn>(fn);
    }();
    memset(map, 0, sizeof *map);
    if (!encode || !P.aligned || P.row_stride >= (1LL << 40) || P.frame_stride >= (1LL << 40)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)P.width * P.bpp, (cuuint64_t)P.height, (cuuint64_t)P.n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)P.row_stride, (cuuint64_t)P.frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_bytes, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(P.rgb), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The pair variant for one source layout.  Returns 1 when the launch was left to the stream variant (a single frame, or more
// than 2^32 regions), 0 after a launch, < 0 on error.
template <int H, int V, int LAYOUT>
static int launch_pair(FusedParams &P, const Geo &g, cudaStream_t st, bool auto_pick)
{
    constexpr int BPP = (LAYOUT == B2J_LAYOUT_RGBA || LAYOUT == B2J_LAYOUT_BGRA) ? 4 : 3;
    using SS = PairShape<H, V, BPP>;
    const int regions_x = (g.mcus_x * 8 * H + SS::kRegW - 1) / SS::kRegW;
    const long long n_regions = (long long)regions_x * g.mcus_y * g.n_frames;
    if (n_regions >= 0xFFF00000LL) return 1;
    static LaunchInfo cache[kMaxDevices];
    LaunchInfo li;
    int rc = launch_info(k_fused_pair<H, V, LAYOUT>, kPairThreads, SS::kSmem, cache, &li);
    if (rc) return rc;
    // a few regions per warp only (a single frame): the stream variant's 32-pixel regions spread wider
    // (one 1080p frame: 10.7 us against 12.6 us)
    if (auto_pick && LAYOUT == B2J_LAYOUT_RGB && n_regions < 3 * (long long)li.per_sm * li.sms * kPairWarps) return 1;
    // regions per work item: the run length (<= kPairRun) with the smallest load of the busiest warp, an item
    // counted as its regions plus a fifth of a region for its own arithmetic
    const long long slots = (long long)li.per_sm * li.sms * kPairWarps;
    int run = 1;
    long long best = -1;
    for (int r = 1; r <= kPairRun && r <= regions_x; r++) {
        const long long items = n_regions / regions_x * ((regions_x + r - 1) / r);
        const long long cost = (items + slots - 1) / slots * (5 * r + 1);
        if (best < 0 || cost < best) { best = cost; run = r; }
    }
    const long long n_items = n_regions / regions_x * ((regions_x + run - 1) / run);
    const long long ctas_needed = (n_items + kPairWarps - 1) / kPairWarps;
    const long long grid = ctas_needed < (long long)li.per_sm * li.sms ? ctas_needed : (long long)li.per_sm * li.sms;
    CUtensorMap map;
    if (!make_frame_map(P, SS::kRawPitch, SS::kRows, &map)) P.aligned = 0;
    k_fused_pair<H, V, LAYOUT><<<(unsigned)grid, kPairThreads, SS::kSmem, st>>>(P, map, (unsigned)n_regions, regions_x, run);
    B2J_LAUNCH_CHECK("k_fused_pair");
    return 0;
}

template <int H, int V, bool FIXED>
static int launch_fused(const FusedParams &P0, const Geo &g, int variant, cudaStream_t st, bool auto_pick = false)
{
    FusedParams P = P0;
    auto simple = [&]() -> int {
        const long long nblk = (long long)g.n_frames * g.blocks_per_frame;
        k_fused_simple<H, V, FIXED><<<(unsigned)((nblk + 127) / 128), 128, 0, st>>>(P);
        B2J_LAUNCH_CHECK("k_fused_simple");
        return B2J_OK;
    };
    if (variant == B2J_FUSED_SIMPLE) return simple();
    if (variant == B2J_FUSED_PAIR && !FIXED) {
        int rc;
        switch (g.layout) {
        case B2J_LAYOUT_BGR: rc = launch_pair<H, V, B2J_LAYOUT_BGR>(P, g, st, auto_pick); break;
        case B2J_LAYOUT_RGBA: rc = launch_pair<H, V, B2J_LAYOUT_RGBA>(P, g, st, auto_pick); break;
        case B2J_LAYOUT_BGRA: rc = launch_pair<H, V, B2J_LAYOUT_BGRA>(P, g, st, auto_pick); break;
        default: rc = launch_pair<H, V, B2J_LAYOUT_RGB>(P, g, st, auto_pick); break;
        }
        if (rc <= 0) return rc;
        variant = g.layout == B2J_LAYOUT_RGB ? B2J_FUSED_STREAM : B2J_FUSED_SIMPLE;
    }
    if (variant == B2J_FUSED_PAIR) variant = g.layout == B2J_LAYOUT_RGB ? B2J_FUSED_STREAM : B2J_FUSED_SIMPLE;   // (gateware DCT)
    if (variant == B2J_FUSED_SIMPLE) return simple();
    if (variant == B2J_FUSED_STREAM) {
        using SS = StreamShape<H, V>;
        const int regions_x = (g.mcus_x * 8 * H + SS::kRegW - 1) / SS::kRegW;
        const long long n_regions = (long long)regions_x * g.mcus_y * g.n_frames;
        if (n_regions < 0xFFF00000LL) {
            static LaunchInfo cache[kMaxDevices];
            LaunchInfo li;
            int rc = launch_info(k_fused_stream<H, V, FIXED>, kStreamThreads, SS::kSmem, cache, &li);
            if (rc) return rc;
            // runs of regions per work item only when every resident warp still gets several items
            const long long slots = (long long)li.per_sm * li.sms * kStreamWarps;
            int run = (int)(n_regions / (slots * 4));
            run = run < 1 ? 1 : (run > kStreamRun ? kStreamRun : run);
            const long long n_items = n_regions / regions_x * ((regions_x + run - 1) / run);
            const long long ctas_needed = (n_items + kStreamWarps - 1) / kStreamWarps;
            const long long grid = ctas_needed < (long long)li.per_sm * li.sms ? ctas_needed : (long long)li.per_sm * li.sms;
            CUtensorMap map;
            if (!make_frame_map(P, SS::kRawPitch, SS::kRows, &map)) P.aligned = 0;
            k_fused_stream<H, V, FIXED><<<(unsigned)grid, kStreamThreads, SS::kSmem, st>>>(P, map, (unsigned)n_regions, regions_x, run);
            B2J_LAUNCH_CHECK("k_fused_stream");
            return B2J_OK;
        }
        // (more than 2^32 regions in one call: the tiled kernel has 64-bit tile indices)
    }
    using TS = TileShape<H, V>;
    const int tiles_x = (g.mcus_x * 8 * H + kTileW - 1) / kTileW;
    const long long n_tiles = (long long)tiles_x * g.mcus_y * g.n_frames;
    static LaunchInfo cache[kMaxDevices];
    LaunchInfo li;
    int rc = launch_info(k_fused_tiled<H, V, FIXED>, kThreads, TS::kSmem, cache, &li);
    if (rc) return rc;
    const long long grid = n_tiles < (long long)li.per_sm * li.sms ? n_tiles : (long long)li.per_sm * li.sms;   // persistent CTAs
    k_fused_tiled<H, V, FIXED><<<(unsigned)grid, kThreads, TS::kSmem, st>>>(P, n_tiles, tiles_x);
    B2J_LAUNCH_CHECK("k_fused_tiled");
    return B2J_OK;
}

}  // namespace b2j

using namespace b2j;

extern "C" int b2j_fused_block(const uint8_t *rgb, const b2j_image_desc *desc, const uint8_t qy_host[64],
                               const uint8_t qc_host[64], int16_t *coefs, int variant, void *stream)
{
    B2J_REQUIRE(rgb && coefs && qy_host && qc_host, B2J_E_ARG, "b2j_fused_block: null pointer");
    const bool fixed = (variant & B2J_FUSED_DCT_FIXED) != 0;
    const bool wide = (variant & B2J_FUSED_PAIR_WIDE) != 0;
    variant &= ~(B2J_FUSED_DCT_FIXED | B2J_FUSED_PAIR_WIDE);
    B2J_REQUIRE(variant >= B2J_FUSED_AUTO && variant <= B2J_FUSED_PAIR, B2J_E_ARG, "b2j_fused_block: bad variant %d", variant);
    B2J_REQUIRE(((uintptr_t)coefs & 15) == 0, B2J_E_ALIGN, "b2j_fused_block: coefs must be 16-byte aligned");
    Geo g;
    int rc = make_geo(desc, &g);
    if (rc) return rc;
    for (int i = 0; i < 64; i++)
        B2J_REQUIRE(qy_host[i] >= 1 && qc_host[i] >= 1, B2J_E_ARG, "b2j_fused_block: zero quantiser at %d", i);
    FusedParams P;
    P.rgb = rgb; P.coefs = coefs;
    P.row_stride = g.row_stride; P.frame_stride = g.frame_stride;
    P.width = g.width; P.height = g.height; P.mcus_x = g.mcus_x; P.mcus_y = g.mcus_y; P.n_frames = g.n_frames;
    P.aligned = (((uintptr_t)rgb & 15) == 0 && (g.row_stride & 15) == 0 && (g.frame_stride & 15) == 0) ? 1 : 0;
    P.bpp = g.bpp;
    P.ro = (g.layout == B2J_LAYOUT_BGR || g.layout == B2J_LAYOUT_BGRA) ? 2 : 0;
    make_quant_mul(qy_host, qc_host, &P.qm);
    const bool fast_ok = make_quant_f32(qy_host, qc_host, &P.qf, wide ? 0.25 : kFastPathErr);
    cudaStream_t st = as_stream(stream);
#ifndef B2J_FUSED_DEFAULT
#define B2J_FUSED_DEFAULT B2J_FUSED_PAIR
#endif
    const bool auto_pick = variant == B2J_FUSED_AUTO;
    if (variant == B2J_FUSED_AUTO) variant = B2J_FUSED_DEFAULT;
    if (variant == B2J_FUSED_PAIR && !fast_ok) variant = B2J_FUSED_STREAM;    // a divisor 1: exact arithmetic throughout
    if (g.layout != B2J_LAYOUT_RGB) {
        // [ext] other source layouts are read in place by the pair variant; what that variant does not cover (gateware
        // DCT, a divisor 1) goes to the one-thread-per-block kernel
        B2J_REQUIRE(auto_pick || variant == B2J_FUSED_PAIR || variant == B2J_FUSED_SIMPLE, B2J_E_ARG,
                    "b2j_fused_block: only the pair and simple variants read layouts other than RGB");
        if (variant != B2J_FUSED_PAIR || fixed) variant = B2J_FUSED_SIMPLE;
    }
    if (fixed) {
        if (g.h == 1) return launch_fused<1, 1, true>(P, g, variant, st);
        if (g.v == 1) return launch_fused<2, 1, true>(P, g, variant, st);
        return launch_fused<2, 2, true>(P, g, variant, st);
    }
    if (g.h == 1) return launch_fused<1, 1, false>(P, g, variant, st, auto_pick);
    if (g.v == 1) return launch_fused<2, 1, false>(P, g, variant, st, auto_pick);
    return launch_fused<2, 2, false>(P, g, variant, st, auto_pick);
}
