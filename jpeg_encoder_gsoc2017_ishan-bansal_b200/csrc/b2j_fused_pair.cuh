// K1, "pair" variant: the stream kernel's organisation (every WARP an independent worker on regions,
// tensor-copy loads, no CTA barrier in the steady state) with regions of 64 x (8|16) pixels and a
// phase B that works on EIGHT 8x8 blocks per round: lane = one row, then one column, of a PAIR of
// blocks, the two blocks in the halves of packed single-precision registers (FADD2 / FMUL2 / FFMA2:
// one issue slot for two operations).
// replaces: the loop body test/model/enc_frame.py:203-217 of the reference.
//
// Exactness.  The model's DCT is a float64 flow graph followed by int(v / 8) and a rounding division
// (enc_frame.py:86-129); the float32 evaluation of the same graph is only an approximation a32 of
// a = v / 8 with |a32 - a| <= E_A = 1/1024 (b2j_common.cuh kFastPathErr; the worst-case forward error bound
// of the graph plus the roundings of the quantiser is 7.8e-4: tests/test_fastpath_model.py derives it and
// checks it on random and extreme blocks, where the largest distance seen is 1.0e-4).  The quantised value is a
// step function of |a| with steps at the integers T_j = j d - floor(d / 2).  The kernel quantises the
// two ends of the interval [|a32| - E_A, |a32| + E_A]: when both give the same integer, that integer
// is the model's value; when they differ (a few blocks per thousand), the warp recomputes the four
// blocks concerned with the exact float64 graph (stream_round of the stream variant).  The four
// coefficients (0|4, 0|4) are sums of integers, exact in float32, and need no interval.
// Tables with a divisor 1 (IJG quality >= 93 or so) are not supported: the host uses the stream
// variant for them (odd divisors shift |a| by 1/2; for d = 1 that puts every zero coefficient on a
// rounding tie).
// Measured on B200, 256 x 1080p 4:2:0 q75 (profiles/r02/experiments.txt): 1.085 ms against 1.31 ms of the stream
// variant, 0.92 G instead of 1.20 G warp-instructions; 0.4 % of the blocks take the exact path.
#pragma once
#include "b2j_fused_stream.cuh"

namespace b2j {

#ifndef B2J_PAIR_THREADS
#define B2J_PAIR_THREADS 256
#endif
#ifndef B2J_PAIR_MIN_CTAS
#define B2J_PAIR_MIN_CTAS 2
#endif
#ifndef B2J_PAIR_RUN
#define B2J_PAIR_RUN 5
#endif
constexpr int kPairThreads = B2J_PAIR_THREADS;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kPairRun = B2J_PAIR_RUN;          // regions per work item when there is work to spare

typedef unsigned long long f32x2;               // two floats in an aligned register pair

__device__ __forceinline__ f32x2 f2_pack(uint32_t lo, uint32_t hi)
{
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, uint32_t &lo, uint32_t &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_both(float c) { return f2_pack(__float_as_uint(c), __float_as_uint(c)); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// cos / sin of n pi / 16 (enc_frame.py:38-47) rounded to float32; the sqrt2 of the third rotation is folded
// into its constants, the sqrt2 of outputs 3 and 5 (enc_frame.py:76-77) into the quantiser's multipliers.
#define B2J_F32_C1 0.98078528040323043f
#define B2J_F32_S1 0.19509032201612825f
#define B2J_F32_C3 0.83146961230254524f
#define B2J_F32_S3 0.55557023301960218f
#define B2J_F32_C6R2 0.54119610014619701f
#define B2J_F32_S6R2 1.30656296487637650f

// dct_1d (enc_frame.py:54-84) on a pair of vectors; v[k] becomes r[k], r[3] and r[5] without their sqrt2.
__device__ __forceinline__ void dct_1d_pair(f32x2 (&v)[8])
{
    const f32x2 a3 = f2_add(v[3], v[4]), a4 = f2_sub(v[3], v[4]);
    const f32x2 a2 = f2_add(v[2], v[5]), a5 = f2_sub(v[2], v[5]);
    const f32x2 a1 = f2_add(v[1], v[6]), a6 = f2_sub(v[1], v[6]);
    const f32x2 a0 = f2_add(v[0], v[7]), a7 = f2_sub(v[0], v[7]);
    const f32x2 b0 = f2_add(a0, a3), b3 = f2_sub(a0, a3);
    const f32x2 b1 = f2_add(a1, a2), b2 = f2_sub(a1, a2);
    const f32x2 c1 = f2_both(B2J_F32_C1), s1 = f2_both(B2J_F32_S1), c3 = f2_both(B2J_F32_C3), s3 = f2_both(B2J_F32_S3);
    const f32x2 c6 = f2_both(B2J_F32_C6R2), s6 = f2_both(B2J_F32_S6R2);
    const f32x2 b4 = f2_fma(a4, c3, f2_mul(a7, s3)), b7 = f2_sub(f2_mul(a7, c3), f2_mul(a4, s3));
    const f32x2 b5 = f2_fma(a5, c1, f2_mul(a6, s1)), b6 = f2_sub(f2_mul(a6, c1), f2_mul(a5, s1));
    const f32x2 e7 = f2_add(b7, b5), e5 = f2_sub(b7, b5);
    const f32x2 e4 = f2_add(b4, b6), e6 = f2_sub(b4, b6);
    v[2] = f2_fma(b2, c6, f2_mul(b3, s6));
    v[6] = f2_sub(f2_mul(b3, c6), f2_mul(b2, s6));
    v[0] = f2_add(b0, b1);
    v[4] = f2_sub(b0, b1);
    v[1] = f2_add(e7, e4);
    v[7] = f2_sub(e7, e4);
    v[3] = e5;
    v[5] = e6;
}

template <int H, int V, int BPP = 3>
struct PairShape {
    static constexpr int kRegW = 64;                              // pixels per region row
    static constexpr int kRows = 8 * V;
    static constexpr int kRowBytes = kRegW * BPP;                 // 192 (256 for the four-byte layouts)
    // pitches: two lanes of an item pass work on rows two apart; 56 and 24 words put them on disjoint banks (four-byte
    // pixels are read as one 16-byte piece per lane: any pitch is conflict free)
    static constexpr int kRawPitch = BPP == 3 ? 224 : 256;        // = the box width of the tensor copy
    static constexpr int kRawBytes = kRows * kRawPitch;
    static constexpr int kYPitch = 96;
    static constexpr int kYBytes = kRows * kYPitch;
    static constexpr int kCPitch = H == 2 ? 32 : kYPitch;
    static constexpr int kCrOff = 8 * kCPitch;
    static constexpr int kCBytes = 16 * kCPitch;
    static constexpr int kTRow = 80;                              // 8 pairs + 16 bytes (= the float64 row of the fallback)
    static constexpr int kTBlk = 672;                             // 8 rows + 32 bytes
    static constexpr int kTBytes = 4 * kTBlk;
    static constexpr int kStgBytes = 8 * 128;
    static constexpr int kOffY = kRawBytes;
    static constexpr int kOffC = kOffY + kYBytes;
    static constexpr int kOffT = kOffC + kCBytes;
    static constexpr int kOffStg = kOffT + kTBytes;
    static constexpr int kOffBar = kOffStg + kStgBytes;            // the warp's mbarrier
    static constexpr int kWarpBytes = (kOffBar + 16 + 127) / 128 * 128;   // tensor copies want 128-byte aligned targets
    static constexpr int kMcus = 8 / H;                           // MCUs per region
    static constexpr int kMulBytes = 2 * 8 * 48;                  // integer multipliers of the fallback, [class][column][8 + pad]
    static constexpr int kQfBytes = 2 * 64 * 16;                  // float constants, [class][x][u] x (R, Ha, Hb - Ha, Hb - Ha)
    static constexpr int kTableBytes = kMulBytes + kQfBytes;
    static constexpr int kSmem = kTableBytes + kPairWarps * kWarpBytes;
    static constexpr int kRounds = H == 1 ? 3 : V + 1;            // rounds of eight blocks per region
};

// ---- one round of phase B: eight blocks, 8 lanes per pair (A, B) of blocks.
// srcA / srcB: this lane's 8 samples of each block (row lane>>2 of pair lane&3).  t_row / t_col: the lane's write
// row and read column in the transpose buffer.  stg_addr[x]: shared address of coefficient (x, u = lane>>2) of block A
// (block B: + 512).  kq[x]: (R, Ha, Hb - Ha twice) of coefficient (x, u).  Returns non-zero bits where an interval straddles a
// step of the quantiser: low word block A, high word block B.
__device__ __forceinline__ uint2 pair_round(const uint8_t *srcA, const uint8_t *srcB, uint8_t *t_row, const uint8_t *t_col,
                                            const uint32_t (&stg_addr)[8], const float4 (&kq)[8])
{
    const uint2 pa = *reinterpret_cast<const uint2 *>(srcA), pb = *reinterpret_cast<const uint2 *>(srcB);
    f32x2 v[8];
    // sample byte -> float 32768 + byte (the byte lands in mantissa bits 8..15 of 2^15)
    v[0] = f2_pack(__byte_perm(pa.x, 0x47000000u, 0x7504), __byte_perm(pb.x, 0x47000000u, 0x7504));
    v[1] = f2_pack(__byte_perm(pa.x, 0x47000000u, 0x7514), __byte_perm(pb.x, 0x47000000u, 0x7514));
    v[2] = f2_pack(__byte_perm(pa.x, 0x47000000u, 0x7524), __byte_perm(pb.x, 0x47000000u, 0x7524));
    v[3] = f2_pack(__byte_perm(pa.x, 0x47000000u, 0x7534), __byte_perm(pb.x, 0x47000000u, 0x7534));
    v[4] = f2_pack(__byte_perm(pa.y, 0x47000000u, 0x7504), __byte_perm(pb.y, 0x47000000u, 0x7504));
    v[5] = f2_pack(__byte_perm(pa.y, 0x47000000u, 0x7514), __byte_perm(pb.y, 0x47000000u, 0x7514));
    v[6] = f2_pack(__byte_perm(pa.y, 0x47000000u, 0x7524), __byte_perm(pb.y, 0x47000000u, 0x7524));
    v[7] = f2_pack(__byte_perm(pa.y, 0x47000000u, 0x7534), __byte_perm(pb.y, 0x47000000u, 0x7534));
    // row pass (enc_frame.py:92-100).  Differences cancel the offsets; only output 0 carries 8 x (32768 + 128)
    dct_1d_pair(v);
    v[0] = f2_add(v[0], f2_both(-263168.0f));
    {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(t_row);
        o[0] = make_ulonglong2(v[0], v[1]);
        o[1] = make_ulonglong2(v[2], v[3]);
        o[2] = make_ulonglong2(v[4], v[5]);
        o[3] = make_ulonglong2(v[6], v[7]);
    }
    __syncwarp();
    // column pass (enc_frame.py:103-108)
#pragma unroll
    for (int x = 0; x < 8; x++) v[x] = *reinterpret_cast<const f32x2 *>(t_col + x * 80);
    dct_1d_pair(v);
    // int(v / 8) and the rounding division (:108, :111-129) at both ends of the error interval
    f32x2 acc = f2_both(0.0f);
    const f32x2 magic = f2_both(12582912.0f);                      // 1.5 * 2^23: float32 ulp 1
#pragma unroll
    for (int x = 0; x < 8; x++) {
        uint32_t oA, oB;
        f2_unpack(v[x], oA, oB);
        const float fA = fabsf(__uint_as_float(oA)), fB = fabsf(__uint_as_float(oB));
        const f32x2 hi = f2_pack(__float_as_uint(fA - kq[x].y), __float_as_uint(fB - kq[x].y));     // |out| - (H - e)
        const f32x2 lo = f2_sub(hi, f2_pack(__float_as_uint(kq[x].z), __float_as_uint(kq[x].w)));   // |out| - (H + e)
        const uint32_t r = __float_as_uint(kq[x].x);
        const f32x2 rr = f2_pack(r ^ (oA & 0x80000000u), r ^ (oB & 0x80000000u));                   // R with the sign of out
        const f32x2 z1 = f2_fma(hi, rr, magic), z2 = f2_fma(lo, rr, magic);
        uint32_t z1A, z1B;
        f2_unpack(z1, z1A, z1B);
        const f32x2 dz = f2_sub(z1, z2);                           // 0 or +-1 (as floats: on the multiplier pipe, the ALU pipe
        acc = f2_fma(dz, dz, acc);                                 // is the busier one; XOR / OR of the words measured 2 % slower)
        sts16(stg_addr[x], z1A);                                   // low half of 0x4B400000 + q = q as int16
        sts16(stg_addr[x] + 512u, z1B);
    }
    __syncwarp();
    uint32_t accA, accB;
    f2_unpack(acc, accA, accB);
    return make_uint2(accA, accB);
}

// The exact float64 graph for the four blocks of one half of a round (rare: kept out of line so that the loop body
// stays small).  off: 0 for the A blocks, 512 for the B blocks of the staging lines.
struct StgAddr {
    uint32_t a[8];
};
__device__ __noinline__ void pair_exact_round(const uint8_t *src, uint8_t *t_row, const uint8_t *t_col, StgAddr sa, uint32_t off,
                                              const uint32_t *mul8)
{
    uint32_t s[8];
#pragma unroll
    for (int x = 0; x < 8; x++) s[x] = sa.a[x] + off;
    stream_round<false>(src, t_row, t_col, s, mul8);
}

// LAYOUT: B2J_LAYOUT_RGB | BGR | RGBA | BGRA of the source frames ([ext] SURVEY 8f-4: read in place, no ingest pass).
template <int H, int V, int LAYOUT = B2J_LAYOUT_RGB>
__global__ void __launch_bounds__(kPairThreads, B2J_PAIR_MIN_CTAS)
k_fused_pair(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap tmap, unsigned n_regions, int regions_x,
             int kRun)
{
    pdl_launch_dependents();                                         // (the entropy kernel's CTAs may queue up behind this grid)
    constexpr int BPP = (LAYOUT == B2J_LAYOUT_RGBA || LAYOUT == B2J_LAYOUT_BGRA) ? 4 : 3;
    constexpr int RO = (LAYOUT == B2J_LAYOUT_BGR || LAYOUT == B2J_LAYOUT_BGRA) ? 2 : 0, BO = 2 - RO;   // byte of r and b in a pixel
    using SS = PairShape<H, V, BPP>;
    constexpr int BPM = H * V + 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_mul = reinterpret_cast<uint32_t *>(smem);            // [class][u][12]: mul[class][x * 8 + u] at x
    float4 *s_qf = reinterpret_cast<float4 *>(smem + SS::kMulBytes); // [class][x][u]
    // static: their shared addresses are compile-time constants (an immediate offset of the predicated byte load)
    __shared__ uint8_t s_crx[256];
    __shared__ uint32_t s_exc[4096];                                 // kYcExceptionBits
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 128; i += kPairThreads) {
        const int cls = i >> 6, x = (i >> 3) & 7, u = i & 7;
        s_mul[(cls * 8 + u) * 12 + x] = P.qm.mul[cls][x * 8 + u];
        const float w2 = P.qf.hb[cls][x * 8 + u] - P.qf.ha[cls][x * 8 + u];           // width of the interval (a register pair)
        s_qf[i] = make_float4(P.qf.r[cls][x * 8 + u], P.qf.ha[cls][x * 8 + u], w2, w2);
    }
    for (int i = tid; i < 256; i += kPairThreads) s_crx[i] = (uint8_t)cr_exception_class((uint32_t)i);
    for (int i = tid; i < 4096; i += kPairThreads) s_exc[i] = kYcExceptionBits[i];

    uint8_t *wb = smem + SS::kTableBytes + warp * SS::kWarpBytes;
    uint8_t *raw = wb, *s_y = wb + SS::kOffY, *s_c = wb + SS::kOffC, *s_t = wb + SS::kOffT, *s_stg = wb + SS::kOffStg;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wb + SS::kOffBar);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- lane constants of phase B
    const int blk = lane & 3, idx = lane >> 2;                       // pair of the round; row, then column
    uint8_t *t_row = s_t + blk * SS::kTBlk + idx * SS::kTRow;
    const uint8_t *t_col = s_t + blk * SS::kTBlk + idx * 8;
    uint32_t stg_addr[8];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        const int zz = kZigzagDev[x * 8 + idx];
        // (the word index inside a block's 128-byte line is XOR-ed with a per-block constant: see the stream kernel)
        stg_addr[x] = sm_u32(s_stg) + blk * 128 + ((((zz >> 1) ^ kStgXor[blk]) & 31) << 2) + ((zz & 1) << 1);
    }
    const int oblk = lane >> 3, ochunk = lane & 7;                   // copy-out: block (of each half), 16-byte piece
    const uint4 *stg_out = reinterpret_cast<const uint4 *>(s_stg + oblk * 128 + ((((ochunk * 4) ^ kStgXor[oblk]) & 31) << 2));

    // ---- region walk: as in the stream kernel (runs of up to kRun consecutive regions of one MCU row per work item)
    const unsigned total_warps = gridDim.x * kPairWarps;
    const int runs_x = (regions_x + kRun - 1) / kRun;
    const unsigned n_items = n_regions / (unsigned)regions_x * (unsigned)runs_x;
    unsigned q = blockIdx.x * kPairWarps + warp;
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
                const int px = j / BPP, ch = j - BPP * px;
                const int sx = min(rx * SS::kRegW + px, P.width - 1);
                raw[row * SS::kRawPitch + j] = img[(long long)min(row, vr - 1) * P.row_stride + (long long)BPP * sx + ch];
            }
        }
        __syncwarp();

        // ---------------- phase A: colour conversion + subsampling; item = 4 pixels x V rows.  Lanes 0..15 and 16..31
        // of a pass work on rows two apart (see the pitches)
#pragma unroll 1
        for (int pass = 0; pass < 4; pass++) {
            // the +128 of the chroma sums as a register addend of the first multiply-add (as a literal the
            // compiler turns it into a separate XOR per sum)
            uint32_t c_off;
            asm("mov.u32 %0, 0x80000000;" : "=r"(c_off));
            const int grp = lane & 15;
            const int row0 = V == 2 ? 2 * (lane >> 4) + 4 * pass : 2 * (lane >> 4) + (pass & 1) + 4 * (pass >> 1);
            uint32_t sy[V][4], scb[V][4], scr[V][4];
            uint32_t cr_flag[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
                cr_flag[v] = 0u;
                uint32_t cr[4], cg[4], cb[4];
                if (BPP == 3) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(raw + (row0 + v) * SS::kRawPitch + grp * 12);
                    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                    uint32_t c0[4], c2[4];                           // first and third byte of every pixel
                    c0[0] = __byte_perm(w0, 0u, 0x4440); cg[0] = __byte_perm(w0, 0u, 0x4441); c2[0] = __byte_perm(w0, 0u, 0x4442);
                    c0[1] = __byte_perm(w0, 0u, 0x4443); cg[1] = __byte_perm(w1, 0u, 0x4440); c2[1] = __byte_perm(w1, 0u, 0x4441);
                    c0[2] = __byte_perm(w1, 0u, 0x4442); cg[2] = __byte_perm(w1, 0u, 0x4443); c2[2] = __byte_perm(w2, 0u, 0x4440);
                    c0[3] = __byte_perm(w2, 0u, 0x4441); cg[3] = __byte_perm(w2, 0u, 0x4442); c2[3] = __byte_perm(w2, 0u, 0x4443);
#pragma unroll
                    for (int i = 0; i < 4; i++) { cr[i] = RO == 0 ? c0[i] : c2[i]; cb[i] = RO == 0 ? c2[i] : c0[i]; }
                } else {
                    const uint4 q = *reinterpret_cast<const uint4 *>(raw + (row0 + v) * SS::kRawPitch + grp * 16);
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        cr[i] = __byte_perm(w[i], 0u, 0x4440 + RO); cg[i] = __byte_perm(w[i], 0u, 0x4441);
                        cb[i] = __byte_perm(w[i], 0u, 0x4440 + BO);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    sy[v][i] = csc_sum_y(cr[i], cg[i], cb[i]);
                    scb[v][i] = cb[i] * B2J_CONST_CX_HALF - cr[i] * B2J_CONST_CB_R - cg[i] * B2J_CONST_CB_G + c_off;
                    scr[v][i] = cr[i] * B2J_CONST_CX_HALF - cg[i] * B2J_CONST_CR_G - cb[i] * B2J_CONST_CR_B + c_off;
                    // inputs where the float64 model may differ from the exact value (b2j_math.cuh)
                    if ((scr[v][i] & B2J_CONST_EXC_MASK) == 0u) cr_flag[v] |= s_crx[cg[i]];
                }
            }
            // luma candidates: bits 10..23 of a sum are zero, i.e. its bits 8..23 as a 16-bit number are < 4
            uint32_t hit[V], any_hit = 0u;
#pragma unroll
            for (int v = 0; v < V; v++) {
                uint32_t m;
                const uint32_t a = __byte_perm(sy[v][0], sy[v][1], 0x6521), b = __byte_perm(sy[v][2], sy[v][3], 0x6521);
                asm("min.u16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
                hit[v] = ((m - 0x00040004u) & ~m & 0x80008000u) | cr_flag[v];      // some half < 4, or a Cr class
                any_hit |= hit[v];
            }
            if (any_hit) {
                // some pixel of the item is a candidate (every grey pixel is): patch the sums of its row so that the top
                // byte is the model's value.  Rare in photographic content.
#pragma unroll
                for (int v = 0; v < V; v++) {
                    if (hit[v]) {
                        const uint8_t *q = raw + (row0 + v) * SS::kRawPitch + grp * (4 * BPP);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const bool cy = (sy[v][i] & B2J_CONST_EXC_MASK) == 0u, cc = (scr[v][i] & B2J_CONST_EXC_MASK) == 0u;
                            if (cy || cc) {
                                const uint32_t rg = q[BPP * i + RO] * 256u + q[BPP * i + 1];
                                const uint32_t e2 = s_exc[rg >> 4] >> ((rg & 15u) * 2u);  // bit 0: Y, bit 1: Cr of the candidate
                                if (cy) sy[v][i] -= (e2 & 1u) << 24;
                                if (cc) scr[v][i] -= (e2 & 2u) << 23;
                            }
                        }
                    }
                }
            }
            // luma: the component is the top byte of the sum
#pragma unroll
            for (int v = 0; v < V; v++) {
                const uint32_t lo = __byte_perm(sy[v][0], sy[v][1], 0x0073), hi = __byte_perm(sy[v][2], sy[v][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_y + (row0 + v) * SS::kYPitch + grp * 4) = __byte_perm(lo, hi, 0x5410);
            }
            if (H == 1) {
                const uint32_t lo = __byte_perm(scb[0][0], scb[0][1], 0x0073), hi = __byte_perm(scb[0][2], scb[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + row0 * SS::kCPitch + grp * 4) = __byte_perm(lo, hi, 0x5410);
                const uint32_t lo2 = __byte_perm(scr[0][0], scr[0][1], 0x0073), hi2 = __byte_perm(scr[0][2], scr[0][3], 0x0073);
                *reinterpret_cast<uint32_t *>(s_c + SS::kCrOff + row0 * SS::kCPitch + grp * 4) = __byte_perm(lo2, hi2, 0x5410);
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
                const int crow = V == 2 ? row0 >> 1 : row0;
                *reinterpret_cast<uint16_t *>(s_c + crow * SS::kCPitch + grp * 2) = (uint16_t)m;
                *reinterpret_cast<uint16_t *>(s_c + SS::kCrOff + crow * SS::kCPitch + grp * 2) = (uint16_t)(m >> 16);
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

        // ---------------- phase B: rounds of eight blocks
        {
            const int mcu0 = cur_rx * SS::kMcus;
            const int nm = min(SS::kMcus, P.mcus_x - mcu0);                          // MCUs of this region inside the frame
            int16_t *dst0 = P.coefs + (((long long)cur_frame * P.mcus_y + cur_my) * P.mcus_x + mcu0) * (BPM * 64) + ochunk * 8;
            float4 kq[8];
#ifdef B2J_PAIR_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int r = 0; r < SS::kRounds; r++) {
                // blocks of the round: pair p = (A, B); source rows; MCU and block index of A and B of pair `oblk`
                const uint8_t *srcA, *srcB;
                int cls, mcuA, kA, mcuB, kB;
                if (H == 2 && V == 2) {
                    if (r < 2) {                                      // luma: upper and lower block of a column of blocks
                        srcA = s_y + idx * SS::kYPitch + (4 * r + blk) * 8; srcB = srcA + 8 * SS::kYPitch; cls = 0;
                        mcuA = mcuB = (4 * r + oblk) >> 1; kA = oblk & 1; kB = 2 + (oblk & 1);
                    } else {                                          // chroma: Cb and Cr block of an MCU
                        srcA = s_c + idx * SS::kCPitch + blk * 8; srcB = srcA + SS::kCrOff; cls = 1;
                        mcuA = mcuB = oblk; kA = 4; kB = 5;
                    }
                } else if (H == 2) {
                    if (r == 0) {                                     // luma: blocks p and p + 4 of the row of blocks
                        srcA = s_y + idx * SS::kYPitch + blk * 8; srcB = srcA + 32; cls = 0;
                        mcuA = oblk >> 1; mcuB = 2 + (oblk >> 1); kA = kB = oblk & 1;
                    } else {
                        srcA = s_c + idx * SS::kCPitch + blk * 8; srcB = srcA + SS::kCrOff; cls = 1;
                        mcuA = mcuB = oblk; kA = 2; kB = 3;
                    }
                } else {                                              // 4:4:4: blocks p and p + 4 of one component
                    const uint8_t *plane = r == 0 ? s_y : s_c + (r - 1) * SS::kCrOff;
                    srcA = plane + idx * SS::kYPitch + blk * 8; srcB = srcA + 32; cls = r == 0 ? 0 : 1;
                    mcuA = oblk; mcuB = 4 + oblk; kA = kB = r;
                }
                if (r == 0 || (cls == 1 && (H == 2 ? r == SS::kRounds - 1 : r == 1))) {
#pragma unroll
                    for (int x = 0; x < 8; x++) kq[x] = s_qf[(cls * 8 + x) * 8 + idx];
                }
                const uint2 acc = pair_round(srcA, srcB, t_row, t_col, stg_addr, kq);
                if (__any_sync(0xFFFFFFFFu, (acc.x | acc.y) != 0u)) {
                    // an interval straddles a step somewhere: the exact float64 graph for the four blocks of that half
                    StgAddr sa;
#pragma unroll
                    for (int x = 0; x < 8; x++) sa.a[x] = stg_addr[x];
                    if (__any_sync(0xFFFFFFFFu, acc.x != 0u)) pair_exact_round(srcA, t_row, t_col, sa, 0u, s_mul + (cls * 8 + idx) * 12);
                    if (__any_sync(0xFFFFFFFFu, acc.y != 0u)) pair_exact_round(srcB, t_row, t_col, sa, 512u, s_mul + (cls * 8 + idx) * 12);
                }
                const uint4 valA = stg_out[0], valB = stg_out[32];
                if (mcuA < nm) *reinterpret_cast<uint4 *>(dst0 + (mcuA * BPM + kA) * 64) = valA;
                if (mcuB < nm) *reinterpret_cast<uint4 *>(dst0 + (mcuB * BPM + kB) * 64) = valB;
            }
        }
        if (!more) break;
    }
}

}  // namespace b2j
