// K1, "wide" variant: every WARP is an independent worker on regions of 256 x (8|16) pixels
// (4:2:0: 16 MCUs = 96 blocks = three rounds of one 8x8 block per lane).  One CTA of 12 warps per
// SM, no CTA-wide barrier in the steady state: the integer-bound colour phase of one warp overlaps
// the FP64-bound DCT phase of the others, and there is no shared-memory transpose (the block stays
// in the registers of its lane, b2j_block.cuh).
// replaces: the loop body test/model/enc_frame.py:203-217 of the reference.
//
// Per region (private to the warp, __syncwarp only):
//   load   eight tensor copies (cp.async.bulk.tensor.3d, UTMALDG; box = 96 bytes x rows of the frames
//          viewed as a [frame][row][byte] tensor), one per 32-pixel sub-tile, on the warp's own
//          mbarrier; issued by one lane as soon as phase A has consumed the raw buffer, so they land
//          during phase B.  Bytes outside the frame arrive as zeros and are replaced by replicas of
//          the last column / row after landing (edge regions only);
//   A      lane = 4 pixels x V rows: integer colour sums, luma bytes and (pair-meaned) chroma bytes
//          into 8-bit planes; the rare inputs where the float64 model differs are detected in line
//          (luma: running minimum of the masked sums; Cr: class table of the green byte) and
//          patched from the tables;
//   B      lane = one 8x8 block: the model's DCT graph in registers, quantise, zig-zag, 8 x 16-byte
//          stores (b2j_block.cuh).
#pragma once
#include <cuda.h>
#include "b2j_block.cuh"
#include "b2j_common.cuh"
#include "b2j_math.cuh"

namespace b2j {

#ifndef B2J_WIDE_WARPS
#define B2J_WIDE_WARPS 12
#endif
constexpr int kWideWarps = B2J_WIDE_WARPS;
constexpr int kWideThreads = kWideWarps * 32;

template <int H, int V>
struct WideShape {
    static constexpr int kRegW = 256;                             // pixels per region row
    static constexpr int kSubW = 32;                              // pixels per sub-tile (one tensor copy)
    static constexpr int kSubs = kRegW / kSubW;
    static constexpr int kRows = 8 * V;
    static constexpr int kSubPitch = kSubW * 3;                   // 96: sub-tiles are dense
    static constexpr int kSubBytes = kRows * kSubPitch;
    static constexpr int kRawBytes = kSubs * kSubBytes;
    static constexpr int kYPitch = kRegW + 16;                    // 68 words: conflict-free plane stores
    static constexpr int kYBytes = kRows * kYPitch;
    static constexpr int kCPitch = kRegW / H + 16;
    static constexpr int kCBytes = 8 * kCPitch;                   // one chroma component
    static constexpr int kOffY = kRawBytes;
    static constexpr int kOffC = kOffY + kYBytes;
    static constexpr int kOffBar = kOffC + 2 * kCBytes;
    static constexpr int kWarpBytes = (kOffBar + 16 + 127) / 128 * 128;   // tensor copies want 128-byte aligned targets
    static constexpr int kMcus = kRegW / (8 * H);                 // MCUs per region
    static constexpr int kTableBytes = 256 + 2 * 64 * 4;          // Cr exception class of every green value; multipliers
    static constexpr int kSmem = kTableBytes + kWideWarps * kWarpBytes;
};

template <int H, int V, bool FIXED>
__global__ void __launch_bounds__(kWideThreads, 1)
k_fused_wide(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap tmap, unsigned n_regions, int regions_x)
{
    using WS = WideShape<H, V>;
    constexpr int BPM = H * V + 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);          // (tells the compiler that it is warp-uniform)
    if (tid == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    uint8_t *wb = smem + warp * WS::kWarpBytes;
    uint8_t *raw = wb, *s_y = wb + WS::kOffY, *s_c = wb + WS::kOffC;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wb + WS::kOffBar);
    uint8_t *s_crx = smem + kWideWarps * WS::kWarpBytes;             // [256]
    uint32_t *s_mul = reinterpret_cast<uint32_t *>(s_crx + 256);     // [2][64]
    for (int i = tid; i < 256; i += kWideThreads) s_crx[i] = (uint8_t)cr_exception_class((uint32_t)i);
    for (int i = tid; i < 128; i += kWideThreads) s_mul[i] = P.qm.mul[i >> 6][i & 63];
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- region walk: r = (frame * mcus_y + my) * regions_x + rx, stepped by the number of warps
    const unsigned total_warps = gridDim.x * kWideWarps;
    unsigned r = blockIdx.x * kWideWarps + warp;
    if (r >= n_regions) return;
    int rx = (int)(r % (unsigned)regions_x);
    unsigned rowidx = r / (unsigned)regions_x;
    int my = (int)(rowidx % (unsigned)P.mcus_y), frame = (int)(rowidx / (unsigned)P.mcus_y);
    const int step_rx = (int)(total_warps % (unsigned)regions_x);
    const unsigned step_rows = total_warps / (unsigned)regions_x;
    const int step_my = (int)(step_rows % (unsigned)P.mcus_y), step_frame = (int)(step_rows / (unsigned)P.mcus_y);

    // sub-tiles that hold source pixels / that are needed after padding to whole MCUs
    auto subs_valid = [&](int rx_) { return min(WS::kSubs, (P.width - rx_ * WS::kRegW + WS::kSubW - 1) / WS::kSubW); };
    auto issue_load = [&](int frame_, int my_, int rx_) {
        if (lane == 0) {
            const int ns = subs_valid(rx_);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)(ns * WS::kSubBytes));
            for (int s = 0; s < ns; s++)
                tma_load_3d(raw + s * WS::kSubBytes, &tmap, (rx_ * WS::kSubs + s) * WS::kSubPitch, my_ * WS::kRows, frame_, bar);
        }
    };
    uint32_t phase = 0;
    if (P.aligned) issue_load(frame, my, rx);

    for (;;) {
        const int x0 = rx * WS::kRegW;
        const int valid_px = min(WS::kRegW, P.width - x0);                         // source pixels in the region
        const int need_px = min(WS::kRegW, P.mcus_x * 8 * H - x0);                 // after padding to whole MCUs
        const int vr = min(WS::kRows, P.height - my * WS::kRows);
        // ---------------- the region's pixels
        if (P.aligned) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            // layouts a tensor map cannot describe: byte-wise
            const uint8_t *img = P.rgb + (long long)frame * P.frame_stride + (long long)(my * WS::kRows) * P.row_stride + 3LL * x0;
            for (int row = 0; row < vr; row++)
                for (int i = lane; i < valid_px * 3; i += 32) {
                    const int sub = i / WS::kSubPitch;
                    raw[sub * WS::kSubBytes + row * WS::kSubPitch + (i - sub * WS::kSubPitch)] = img[(long long)row * P.row_stride + i];
                }
            __syncwarp();
        }
        if (valid_px < need_px || vr < WS::kRows) {
            // edge regions: replicate the last column, then the last row (enc_frame.py has no padding rule:
            // DESIGN.md "[ext] definitions")
            __syncwarp();
            const int pad = need_px - valid_px;
            for (int i = lane; i < vr * pad * 3; i += 32) {
                const int row = i / (pad * 3), j = i - row * (pad * 3);
                const int px = valid_px + j / 3, ch = j % 3;
                const int sp = px / WS::kSubW, sl = (valid_px - 1) / WS::kSubW;
                raw[sp * WS::kSubBytes + row * WS::kSubPitch + (px - sp * WS::kSubW) * 3 + ch] =
                    raw[sl * WS::kSubBytes + row * WS::kSubPitch + (valid_px - 1 - sl * WS::kSubW) * 3 + ch];
            }
            __syncwarp();
            uint32_t *raw32 = reinterpret_cast<uint32_t *>(raw);
            constexpr int kRowWords = WS::kSubPitch / 4;
            for (int s = 0; s * WS::kSubW < need_px; s++)
                for (int i = lane; i < (WS::kRows - vr) * kRowWords; i += 32) {
                    const int row = i / kRowWords, w = i - row * kRowWords;
                    raw32[s * (WS::kSubBytes / 4) + (vr + row) * kRowWords + w] = raw32[s * (WS::kSubBytes / 4) + (vr - 1) * kRowWords + w];
                }
        }
        __syncwarp();

        // ---------------- phase A: colour conversion + subsampling; item = 4 pixels x V rows,
        // 64 items per sub-tile
        const int n_iter = 2 * ((need_px + WS::kSubW - 1) / WS::kSubW);
#pragma unroll 1
        for (int it = 0; it < n_iter; it++) {
            const int sub = it >> 1, item = lane + 32 * (it & 1);
            const int ry = item >> 3, grp = item & 7;
            // the +128 of the chroma sums as a register addend of the first multiply-add (as a literal the
            // compiler turns it into a separate XOR per sum)
            uint32_t c_off;
            asm("mov.u32 %0, 0x80000000;" : "=r"(c_off));
            uint32_t sy[V][4], scb[V][4], scr[V][4];
            uint32_t cand = 0xFFFFFFFFu, cr_flag = 0u;
            const uint8_t *rsub = raw + sub * WS::kSubBytes + grp * 12;
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(rsub + (ry * V + v) * WS::kSubPitch);
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
                    cand = min(cand, sy[v][i] & B2J_CONST_EXC_MASK);
                    if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) cr_flag |= s_crx[cg[i]];
                }
            }
            if (cand == 0u || cr_flag != 0u) {
                // some pixel of the item is a candidate: patch the sums so that the top byte is the model's value
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const uint8_t *q = rsub + (ry * V + v) * WS::kSubPitch;
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
                *reinterpret_cast<uint32_t *>(s_y + (ry * V + v) * WS::kYPitch + sub * 32 + grp * 4) = __byte_perm(lo, hi, 0x5410);
            }
            if (H == 1) {
                const uint32_t lo = __byte_perm(scb[0][0], scb[0][1], 0x0073), hi = __byte_perm(scb[0][2], scb[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + ry * WS::kCPitch + sub * 32 + grp * 4) = __byte_perm(lo, hi, 0x5410);
                const uint32_t lo2 = __byte_perm(scr[0][0], scr[0][1], 0x0073), hi2 = __byte_perm(scr[0][2], scr[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + WS::kCBytes + ry * WS::kCPitch + sub * 32 + grp * 4) = __byte_perm(lo2, hi2, 0x5410);
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
                *reinterpret_cast<uint16_t *>(s_c + ry * WS::kCPitch + sub * 16 + grp * 2) = (uint16_t)m;
                *reinterpret_cast<uint16_t *>(s_c + WS::kCBytes + ry * WS::kCPitch + sub * 16 + grp * 2) = (uint16_t)(m >> 16);
            }
        }
        __syncwarp();                 // planes complete; raw buffer free

        // ---------------- next region: coordinates, prefetch
        const int cur_frame = frame, cur_my = my;
        r += total_warps;
        const bool more = r < n_regions && r >= total_warps;     // (wrap-around guard)
        if (more) {
            rx += step_rx;
            const int carry = rx >= regions_x ? 1 : 0;
            rx -= carry ? regions_x : 0;
            my += step_my + carry;
            frame += step_frame;
            if (my >= P.mcus_y) { my -= P.mcus_y; frame++; }
            if (P.aligned) issue_load(frame, my, rx);
        }

        // ---------------- phase B: one 8x8 block per lane, warp-uniform component class,
        // one code path for luma and chroma
        {
            const int nbx = need_px >> 3;                      // luma blocks per block row
            const int ncx = nbx / H;                           // chroma blocks per component
            const long long mcu_base = ((long long)cur_frame * P.mcus_y + cur_my) * P.mcus_x + x0 / (8 * H);
            constexpr int kRounds = V + (H == 1 ? 2 : 1);
#pragma unroll 1
            for (int round = 0; round < kRounds; round++) {
                const bool luma = round < V;
                int comp, bx;
                if (H == 1) { comp = round - V; bx = lane; }
                else { comp = lane >> 4; bx = lane & 15; }
                if (luma) bx = lane;
                const int nb = luma ? nbx : ncx;
                const uint8_t *plane = luma ? s_y + round * 8 * WS::kYPitch : s_c + comp * WS::kCBytes;
                const int pitch = luma ? WS::kYPitch : WS::kCPitch;
                const int k = luma ? round * H + (bx % H) : H * V + comp;
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
        __syncwarp();                 // planes may be overwritten by the next phase A
        if (!more) break;
    }
}

}  // namespace b2j
