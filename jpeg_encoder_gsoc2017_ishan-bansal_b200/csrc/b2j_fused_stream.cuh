// K1, "stream" variant: every WARP is an independent worker on regions of 32 x (8|16) pixels
// (4:2:0: two MCUs = 12 blocks).  No CTA-wide barrier in the steady state, ~64 registers, so
// 24..32 warps per SM are resident and the integer-bound colour phase of one warp overlaps the
// FP64-bound DCT phase of others.
// replaces: the loop body test/model/enc_frame.py:203-217 of the reference (same arithmetic
// as the tiled variant: b2j_math.cuh).
//
// Per region (all of it private to the warp, __syncwarp only):
//   load   ONE tensor copy (cp.async.bulk.tensor.3d, UTMALDG; box = 112 bytes x rows of the
//          frames viewed as a [frame][row][byte] tensor) into the warp's raw buffer, completion on the
//          warp's own mbarrier; issued by one lane as soon as phase A has consumed the buffer, so it
//          lands during phase B.  (Per-row bulk copies cost ~10 issue slots per row - the operands
//          of UBLKCP are uniform registers -, 16-byte cp.async 96 shared-memory wavefronts per
//          region: both measured slower.)
//   A      lane = 4 pixels x V rows: integer colour sums, luma bytes and (pair-meaned) chroma
//          bytes into 8-bit planes; the rare inputs where the float64 model differs are
//          detected in line (luma: running minimum of the masked sums; Cr: class table of the
//          green byte) and patched from the tables;
//   B      four 8x8 blocks per round, 8 lanes per block: lane = one row (row pass), then one
//          column (column pass) of the model's DCT graph; the 8x8 intermediate goes through a
//          padded shared-memory transpose; the column lanes quantise and scatter their eight
//          coefficients to zig-zag positions of a 128-byte staging line per block, which the
//          warp then writes with full-line 16-byte stores.
#pragma once
#include <cuda.h>
#include "b2j_common.cuh"
#include "b2j_math.cuh"

namespace b2j {

#ifndef B2J_STREAM_THREADS
#define B2J_STREAM_THREADS 256
#endif
#ifndef B2J_STREAM_MIN_CTAS
#define B2J_STREAM_MIN_CTAS 3
#endif
#ifndef B2J_STREAM_UNROLL_A
#define B2J_STREAM_UNROLL_A 1
#endif
constexpr int kStreamUnrollA = B2J_STREAM_UNROLL_A;
#ifndef B2J_STREAM_RUN
#define B2J_STREAM_RUN 6
#endif
constexpr int kStreamRun = B2J_STREAM_RUN;        // regions per work item when there is work to spare
constexpr int kStreamThreads = B2J_STREAM_THREADS;
constexpr int kStreamWarps = kStreamThreads / 32;

static __device__ const uint8_t kZigzagDev[64] = B2J_ZIGZAG_INIT;
static __device__ const uint8_t kStgXor[4] = {0, 8, 20, 28};

template <int H, int V>
struct StreamShape {
    static constexpr int kRegW = 32;                              // pixels per region row
    static constexpr int kRows = 8 * V;
    static constexpr int kRowBytes = kRegW * 3;                   // 96
    static constexpr int kRawPitch = 112;                         // 28 words: conflict-free item reads (= the box width)
    static constexpr int kRawBytes = kRows * kRawPitch;
    static constexpr int kYBytes = kRows * 32;
    static constexpr int kCPitch = 32 / H;
    static constexpr int kCrOff = H == 2 ? 192 : 8 * kCPitch;     // H == 2: Cb and Cr rows on disjoint banks
    static constexpr int kCBytes = H == 2 ? 320 : 16 * kCPitch;
    static constexpr int kTRow = 80;                              // 8 doubles + 16 bytes
    static constexpr int kTBlk = 672;                             // 8 rows + 32 bytes
    static constexpr int kTBytes = 4 * kTBlk;
    static constexpr int kStgBytes = 4 * 128;
    static constexpr int kOffY = kRawBytes;
    static constexpr int kOffC = kOffY + kYBytes;
    static constexpr int kOffT = kOffC + kCBytes;
    static constexpr int kOffStg = kOffT + kTBytes;
    static constexpr int kOffBar = kOffStg + kStgBytes;            // the warp's mbarrier
    static constexpr int kWarpBytes = (kOffBar + 16 + 127) / 128 * 128;   // tensor copies want 128-byte aligned targets
    static constexpr int kMcus = 4 / H;                           // MCUs per region
    static constexpr int kChunks = kRows * 6;                     // 16-byte pieces of a region
    static constexpr int kMulBytes = 2 * 8 * 48;                  // multipliers, [class][column][8 + pad]
    static constexpr int kTableBytes = kMulBytes + 256;           // + Cr exception class of every green value
    static constexpr int kSmem = kTableBytes + kStreamWarps * kWarpBytes;
};

__device__ __forceinline__ uint32_t sm_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(sm_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

// ---- one round of phase B: four blocks, 8 lanes each.
// src: this lane's 8 samples (row lane>>2 of block lane&3).  t_row / t_col: the lane's write row and
// read column in the transpose buffer.  stg_addr[x]: shared address of coefficient (x, u = lane>>2)
// of block lane&3.  mul8: the lane's eight multipliers (column u).
template <bool FIXED>
__device__ __forceinline__ void stream_round(const uint8_t *src, uint8_t *t_row, const uint8_t *t_col,
                                             const uint32_t (&stg_addr)[8], const uint32_t *mul8)
{
    const uint2 px = *reinterpret_cast<const uint2 *>(src);
    if (FIXED) {
        int s[8];
        s[0] = (int)__byte_perm(px.x, 0u, 0x4440) - 128; s[1] = (int)__byte_perm(px.x, 0u, 0x4441) - 128;
        s[2] = (int)__byte_perm(px.x, 0u, 0x4442) - 128; s[3] = (int)__byte_perm(px.x, 0u, 0x4443) - 128;
        s[4] = (int)__byte_perm(px.y, 0u, 0x4440) - 128; s[5] = (int)__byte_perm(px.y, 0u, 0x4441) - 128;
        s[6] = (int)__byte_perm(px.y, 0u, 0x4442) - 128; s[7] = (int)__byte_perm(px.y, 0u, 0x4443) - 128;
        dct_1d_fixed(s);
        int4 *o = reinterpret_cast<int4 *>(t_row);
        o[0] = make_int4(s[0], s[1], s[2], s[3]);
        o[1] = make_int4(s[4], s[5], s[6], s[7]);
    } else {
        // row pass (enc_frame.py:92-100); outputs 0 and 4 are integers, exact as doubles
        int r0, r4;
        double r1, r2, r3, r5, r6, r7;
#ifndef B2J_NO_PACKED_ROWS
        dct_1d_bytes(px, r0, r4, r1, r2, r3, r5, r6, r7);
#else
        int s[8];
        s[0] = (int)__byte_perm(px.x, 0u, 0x4440) - 128; s[1] = (int)__byte_perm(px.x, 0u, 0x4441) - 128;
        s[2] = (int)__byte_perm(px.x, 0u, 0x4442) - 128; s[3] = (int)__byte_perm(px.x, 0u, 0x4443) - 128;
        s[4] = (int)__byte_perm(px.y, 0u, 0x4440) - 128; s[5] = (int)__byte_perm(px.y, 0u, 0x4441) - 128;
        s[6] = (int)__byte_perm(px.y, 0u, 0x4442) - 128; s[7] = (int)__byte_perm(px.y, 0u, 0x4443) - 128;
        dct_1d_int(s, r0, r4, r1, r2, r3, r5, r6, r7);
#endif
        double2 *o = reinterpret_cast<double2 *>(t_row);
        o[0] = make_double2(i2d(r0), r1);
        o[1] = make_double2(r2, r3);
        o[2] = make_double2(i2d(r4), r5);
        o[3] = make_double2(r6, r7);
    }
    __syncwarp();
    const uint4 m0 = *reinterpret_cast<const uint4 *>(mul8), m1 = *reinterpret_cast<const uint4 *>(mul8 + 4);
    const uint32_t mul[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    if (FIXED) {
        int c[8];
#pragma unroll
        for (int x = 0; x < 8; x++) c[x] = *reinterpret_cast<const int *>(t_col + x * 80);
        dct_1d_fixed(c);
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const int o = c[x] >> 3;                                  // new_dct6.py:107-140
            const uint32_t q = quant_mag((uint32_t)abs(o), mul[x]);
            sts16(stg_addr[x], o < 0 ? 0u - q : q);
        }
    } else {
        // column pass (enc_frame.py:103-108), int(v / 8), quantise (:111-129)
        double c[8];
#pragma unroll
        for (int x = 0; x < 8; x++) c[x] = *reinterpret_cast<const double *>(t_col + x * 80);
        dct_1d(c);
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const uint32_t q = quant_mag(abs_div8_trunc(c[x]), mul[x]);
            sts16(stg_addr[x], __double2hiint(c[x]) < 0 ? 0u - q : q);
        }
    }
    __syncwarp();
}

template <int H, int V, bool FIXED>
__global__ void __launch_bounds__(kStreamThreads, B2J_STREAM_MIN_CTAS)
k_fused_stream(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap tmap, unsigned n_regions, int regions_x,
               int kRun)
{
    pdl_launch_dependents();                                         // (the entropy kernel's CTAs may queue up behind this grid)
    using SS = StreamShape<H, V>;
    constexpr int BPM = H * V + 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_mul = reinterpret_cast<uint32_t *>(smem);            // [class][u][12]: mul[class][x * 8 + u] at x
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *s_crx = smem + SS::kMulBytes;                           // [256]
    for (int i = tid; i < 128; i += kStreamThreads) {
        const int cls = i >> 6, x = (i >> 3) & 7, u = i & 7;
        s_mul[(cls * 8 + u) * 12 + x] = P.qm.mul[cls][x * 8 + u];
    }
    for (int i = tid; i < 256; i += kStreamThreads) s_crx[i] = (uint8_t)cr_exception_class((uint32_t)i);

    uint8_t *wb = smem + SS::kTableBytes + warp * SS::kWarpBytes;
    uint8_t *raw = wb, *s_y = wb + SS::kOffY, *s_c = wb + SS::kOffC, *s_t = wb + SS::kOffT, *s_stg = wb + SS::kOffStg;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wb + SS::kOffBar);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- lane constants of phase B
    const int blk = lane & 3, idx = lane >> 2;                       // block of the round; row, then column
    uint8_t *t_row = s_t + blk * SS::kTBlk + idx * SS::kTRow;
    const uint8_t *t_col = s_t + blk * SS::kTBlk + idx * (FIXED ? 4 : 8);
    uint32_t stg_addr[8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        const int zz = kZigzagDev[x * 8 + idx];
        // the four blocks of a round write the same zig-zag position at the same time: the word index
        // inside a block's 128-byte line is XOR-ed with a per-block constant (fewest bank conflicts of
        // all constants that keep 16-byte pieces whole; found by exhaustive search)
        stg_addr[x] = sm_u32(s_stg) + blk * 128 + ((((zz >> 1) ^ kStgXor[blk]) & 31) << 2) + ((zz & 1) << 1);
    }
    const int oblk = lane >> 3, ochunk = lane & 7;                   // copy-out: block, 16-byte piece
    const uint4 *stg_out = reinterpret_cast<const uint4 *>(s_stg + oblk * 128 + ((((ochunk * 4) ^ kStgXor[oblk]) & 31) << 2));

    // ---- region walk.  A work item is a run of up to kRun consecutive regions of one MCU row (kRun: 1 when there
    // are few regions per warp, e.g. a single frame; kStreamRun for batches); item
    // q = (frame * mcus_y + my) * runs_x + run goes to warp q mod (number of warps).  Inside a run the next region is
    // rx + 1 (a pointer increment); the item arithmetic runs once per run.
    const unsigned total_warps = gridDim.x * kStreamWarps;
    const int runs_x = (regions_x + kRun - 1) / kRun;
    const unsigned n_items = n_regions / (unsigned)regions_x * (unsigned)runs_x;
    unsigned q = blockIdx.x * kStreamWarps + warp;
    if (q >= n_items) return;
    int run = (int)(q % (unsigned)runs_x);
    unsigned rowidx = q / (unsigned)runs_x;
    int my = (int)(rowidx % (unsigned)P.mcus_y), frame = (int)(rowidx / (unsigned)P.mcus_y);
    const int step_run = (int)(total_warps % (unsigned)runs_x);
    const unsigned step_rows = total_warps / (unsigned)runs_x;
    const int step_my = (int)(step_rows % (unsigned)P.mcus_y), step_frame = (int)(step_rows / (unsigned)P.mcus_y);
    int rx = run * kRun, left = min(kRun, regions_x - rx);

    auto region_fast = [&](int rx_) { return P.aligned && (rx_ + 1) * SS::kRegW <= P.width; };
    // rows below the image and bytes right of it arrive as zeros (the former are replicated after landing)
    auto issue_load = [&](int frame_, int my_, int rx_) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, SS::kRows * SS::kRawPitch);
            tma_load_3d(raw, &tmap, rx_ * SS::kRowBytes, my_ * SS::kRows, frame_, bar);
        }
    };
    uint32_t phase = 0;
    if (region_fast(rx)) issue_load(frame, my, rx);

    for (;;) {
        // ---------------- the region's pixels
        if (region_fast(rx)) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            const int vr = P.height - my * SS::kRows;
            if (vr < SS::kRows) {                                   // last MCU row: replicate the last image row
                uint32_t *raw32 = reinterpret_cast<uint32_t *>(raw);
                for (int i = lane; i < (SS::kRows - vr) * (SS::kRawPitch / 4); i += 32) {
                    const int row = i / (SS::kRawPitch / 4), w = i - row * (SS::kRawPitch / 4);
                    raw32[(vr + row) * (SS::kRawPitch / 4) + w] = raw32[(vr - 1) * (SS::kRawPitch / 4) + w];
                }
            }
        } else {
            // right edge / unaligned layouts: byte-wise with clamped coordinates
            const int vr = min(SS::kRows, P.height - my * SS::kRows);
            const uint8_t *img = P.rgb + (long long)frame * P.frame_stride + (long long)(my * SS::kRows) * P.row_stride;
            for (int i = lane; i < SS::kRows * SS::kRowBytes; i += 32) {
                const int row = i / SS::kRowBytes, j = i - row * SS::kRowBytes;
                const int px = j / 3, ch = j - 3 * px;
                const int sx = min(rx * SS::kRegW + px, P.width - 1);
                raw[row * SS::kRawPitch + j] = img[(long long)min(row, vr - 1) * P.row_stride + 3LL * sx + ch];
            }
        }
        __syncwarp();

        // ---------------- phase A: colour conversion + subsampling; item = 4 pixels x V rows
#pragma unroll kStreamUnrollA
        for (int item = lane; item < 64; item += 32) {
            // the +128 of the chroma sums as a register addend of the first multiply-add (as a literal the
            // compiler turns it into a separate XOR per sum)
            uint32_t c_off;
            asm("mov.u32 %0, 0x80000000;" : "=r"(c_off));
            const int ry = item >> 3, grp = item & 7;
            uint32_t sy[V][4], scb[V][4], scr[V][4];
            uint32_t cand[V], cr_flag[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
                cand[v] = 0xFFFFFFFFu;
                cr_flag[v] = 0u;
                const uint32_t *q = reinterpret_cast<const uint32_t *>(raw + (ry * V + v) * SS::kRawPitch + grp * 12);
                const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                uint32_t cr[4], cg[4], cb[4];
                cr[0] = __byte_perm(w0, 0u, 0x4440); cg[0] = __byte_perm(w0, 0u, 0x4441); cb[0] = __byte_perm(w0, 0u, 0x4442);
                cr[1] = __byte_perm(w0, 0u, 0x4443); cg[1] = __byte_perm(w1, 0u, 0x4440); cb[1] = __byte_perm(w1, 0u, 0x4441);
                cr[2] = __byte_perm(w1, 0u, 0x4442); cg[2] = __byte_perm(w1, 0u, 0x4443); cb[2] = __byte_perm(w2, 0u, 0x4440);
                cr[3] = __byte_perm(w2, 0u, 0x4441); cg[3] = __byte_perm(w2, 0u, 0x4442); cb[3] = __byte_perm(w2, 0u, 0x4443);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    sy[v][i] = csc_sum_y(cr[i], cg[i], cb[i]);
                    scb[v][i] = cb[i] * B2J_CONST_CX_HALF - cr[i] * B2J_CONST_CB_R - cg[i] * B2J_CONST_CB_G + c_off;
                    scr[v][i] = cr[i] * B2J_CONST_CX_HALF - cg[i] * B2J_CONST_CR_G - cb[i] * B2J_CONST_CR_B + c_off;
                    // inputs where the float64 model may differ from the exact value (b2j_math.cuh).  Luma
                    // candidates are rare (every grey pixel is one); Cr candidates (g == b, r - g even) are
                    // common, but only eight green values can really differ
#ifndef B2J_NO_ROW_FLAGS
                    cand[v] = min(cand[v], sy[v][i] & B2J_CONST_EXC_MASK);
                    if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) cr_flag[v] |= s_crx[cg[i]];
#else
                    cand[0] = min(cand[0], sy[v][i] & B2J_CONST_EXC_MASK);
                    if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) cr_flag[0] |= s_crx[cg[i]];
#endif
                }
            }
#pragma unroll
            for (int v = 0; v < V; v++) {
#ifndef B2J_NO_ROW_FLAGS
                if (cand[v] == 0u || cr_flag[v] != 0u) {
#else
                if (cand[0] == 0u || cr_flag[0] != 0u) {
#endif
                    // some pixel of this row of the item is a candidate (every grey pixel is): patch the sums so that
                    // the top byte is the model's value.  Rare in photographic content.
                    const uint8_t *q = raw + (ry * V + v) * SS::kRawPitch + grp * 12;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if ((sy[v][i] & B2J_CONST_EXC_MASK) == 0u) sy[v][i] -= y_exception(q[3 * i], q[3 * i + 1]) << 24;
                        if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) scr[v][i] -= cr_exception(q[3 * i], q[3 * i + 1]) << 24;
                    }
                }
            }
            // luma: the component is the top byte of the sum
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t lo = __byte_perm(sy[v][0], sy[v][1], 0x0073), hi = __byte_perm(sy[v][2], sy[v][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_y + (ry * V + v) * 32 + grp * 4) = __byte_perm(lo, hi, 0x5410);
            }
            if (H == 1) {
                const uint32_t lo = __byte_perm(scb[0][0], scb[0][1], 0x0073), hi = __byte_perm(scb[0][2], scb[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + ry * 32 + grp * 4) = __byte_perm(lo, hi, 0x5410);
                const uint32_t lo2 = __byte_perm(scr[0][0], scr[0][1], 0x0073), hi2 = __byte_perm(scr[0][2], scr[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + SS::kCrOff + ry * 32 + grp * 4) = __byte_perm(lo2, hi2, 0x5410);
            } else {
                // pair mean (c0 + c1) >> 1 horizontally (crs.py:39-54), then vertically, on packed bytes:
                // e = [cb0, cb2, cr0, cr2], o = [cb1, cb3, cr1, cr3] of a row; floor((a + b) / 2) per byte
                uint32_t mean[V];
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const uint32_t e = __byte_perm(__byte_perm(scb[v][0], scb[v][2], 0x0073), __byte_perm(scr[v][0], scr[v][2], 0x0073), 0x5410);
                    const uint32_t o = __byte_perm(__byte_perm(scb[v][1], scb[v][3], 0x0073), __byte_perm(scr[v][1], scr[v][3], 0x0073), 0x5410);
                    mean[v] = (e & o) + (((e ^ o) & 0xFEFEFEFEu) >> 1);
                }
                uint32_t m = mean[0];
                if (V == 2) m = (mean[0] & mean[1]) + (((mean[0] ^ mean[1]) & 0xFEFEFEFEu) >> 1);
                *reinterpret_cast<uint16_t *>(s_c + ry * 16 + grp * 2) = (uint16_t)m;
                *reinterpret_cast<uint16_t *>(s_c + SS::kCrOff + ry * 16 + grp * 2) = (uint16_t)(m >> 16);
            }
        }
        __syncwarp();                 // planes complete; raw buffer free

        // ---------------- next region: coordinates, prefetch
        const int cur_frame = frame, cur_my = my, cur_rx = rx;
        bool more = true;
        if (left > 1) {
            left--;
            rx++;
        } else {
            q += total_warps;
            more = q < n_items && q >= total_warps;                  // (wrap-around guard)
            if (more) {
                run += step_run;
                const int carry = run >= runs_x ? 1 : 0;
                run -= carry ? runs_x : 0;
                my += step_my + carry;
                frame += step_frame;
                if (my >= P.mcus_y) { my -= P.mcus_y; frame++; }
                rx = run * kRun;
                left = min(kRun, regions_x - rx);
            }
        }
        if (more && region_fast(rx)) issue_load(frame, my, rx);

        // ---------------- phase B
        {
            const int mcu0 = cur_rx * SS::kMcus;
            const int nm = min(SS::kMcus, P.mcus_x - mcu0);                          // MCUs of this region inside the frame
            int16_t *dst0 = P.coefs + (((long long)cur_frame * P.mcus_y + cur_my) * P.mcus_x + mcu0) * (BPM * 64) + ochunk * 8;
#pragma unroll
            for (int by = 0; by < V; by++) {
                stream_round<FIXED>(s_y + (by * 8 + idx) * 32 + blk * 8, t_row, t_col, stg_addr, s_mul + idx * 12);
                const int mcu = oblk / H, k = by * H + (oblk % H);
                const uint4 val = *stg_out;
                if (mcu < nm) *reinterpret_cast<uint4 *>(dst0 + (mcu * BPM + k) * 64) = val;
            }
            if (H == 1) {
#pragma unroll
                for (int comp = 0; comp < 2; comp++) {
                    stream_round<FIXED>(s_c + comp * SS::kCrOff + idx * 32 + blk * 8, t_row, t_col, stg_addr, s_mul + (8 + idx) * 12);
                    const int mcu = oblk, k = H * V + comp;
                    const uint4 val = *stg_out;
                    if (mcu < nm) *reinterpret_cast<uint4 *>(dst0 + (mcu * BPM + k) * 64) = val;
                }
            } else {
                stream_round<FIXED>(s_c + (blk >> 1) * SS::kCrOff + idx * 16 + (blk & 1) * 8, t_row, t_col, stg_addr, s_mul + (8 + idx) * 12);
                const int mcu = oblk & 1, k = H * V + (oblk >> 1);
                const uint4 val = *stg_out;
                if (mcu < nm) *reinterpret_cast<uint4 *>(dst0 + (mcu * BPM + k) * 64) = val;
            }
        }
        if (!more) break;
    }
}

}  // namespace b2j
