/* b2j.h - C ABI of libb2j.so, the B200 (sm_100a) baseline-JPEG encode path.
 *
 * This is the drop-in boundary.  The reference (LiteJPEG) has no FFI of its
 * own: its operator API is "a class with .sink/.source whose behaviour a bench
 * observes as Packet(list[int]) in -> list[int] out" (test/new_zigzag_tb.py:22-51
 * of the reference).  Each entry point below replaces the datapath of one such
 * operator, or of the software model function with the same job; the citation
 * after "replaces:" is the reference interface, relative to /root/reference.
 * INTEGRATION.md shows the ctypes stub a maintainer would add per bench.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no C++/torch types;
 *   - every `*_dev` / unmarked data pointer is caller-allocated DEVICE memory on
 *     the current CUDA device; `*_host` pointers are host memory; quantisation
 *     tables are 64-entry HOST arrays in natural (row-major) order, read at call
 *     time and passed by value to the kernel;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); work is
 *     asynchronous on that stream unless stated; the library keeps no global
 *     mutable state, so calls are re-entrant, one stream per caller thread;
 *   - return value: 0 = OK, < 0 = argument error (B2J_E_*), > 0 = cudaError_t;
 *     never throws, never aborts.  b2j_last_error() gives a message for the
 *     calling thread.
 */
#ifndef B2J_H
#define B2J_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2J_OK 0
#define B2J_E_ARG (-1)        /* null pointer, bad enum, negative size */
#define B2J_E_SHAPE (-2)      /* unsupported geometry (e.g. side > 65535) */
#define B2J_E_WORKSPACE (-3)  /* workspace too small */
#define B2J_E_CAPACITY (-4)   /* output slot too small for a frame (sizes[] holds the need, negated) */
#define B2J_E_ALIGN (-5)      /* a pointer that must be 16-byte aligned is not */

#define B2J_SUB_444 444
#define B2J_SUB_422 422
#define B2J_SUB_420 420

/* ZRL policy of the run-length stage */
#define B2J_ZRL_EAGER 0     /* gateware rule, litejpeg/core/rle/rlecore.py:101-105 */
#define B2J_ZRL_STANDARD 1  /* T.81 / libjpeg rule: ZRL only in front of a non-zero */

/* colour model of b2j_rgb2ycbcr */
#define B2J_CSC_JFIF 0      /* software model, test/model/enc_frame.py:27-35 */
#define B2J_CSC_XAPP930 1   /* gateware, litejpeg/core/csc.py:85-160 (coef_w = 8) */

/* rounding of b2j_quantize */
#define B2J_Q_ROUND 0       /* model, test/model/enc_frame.py:111-129 (half away from zero) */
#define B2J_Q_TRUNC 1       /* older model, test/model/encoder.py:84-88 (int(v/d)) */
#define B2J_Q_RECIPROCAL 2  /* gateware, litejpeg/core/quantization.py:124-168 */

/* variant of b2j_fused_block */
#define B2J_FUSED_AUTO 0    /* the fastest variant (pair) */
#define B2J_FUSED_SIMPLE 1  /* one thread per 8x8 block straight from global memory (cross-check) */
#define B2J_FUSED_TILED 2   /* CTA-wide shared-memory tiles, bulk-async loads, one thread per block */
#define B2J_FUSED_STREAM 3  /* warp-private 32-pixel regions, 8 lanes per block, no CTA barriers */
#define B2J_FUSED_PAIR 4    /* stream organisation, 64-pixel regions; DCT of two blocks per lane in packed float32 with an
                             * error interval per coefficient and the exact float64 graph wherever the interval
                             * straddles a step of the quantiser (same results as every other variant); falls back to
                             * STREAM for B2J_FUSED_DCT_FIXED and for tables with a divisor 1 */
/* OR-ed into `variant`: [ext] DCT in the gateware's fixed-point arithmetic
 * (litejpeg/core/new_dct6.py:171-396: its constants and shift points, the two
 * wiring slips :262-263 / :294 corrected, 32-bit sums) instead of the float64
 * model.  NOT parity with the model (<= 4 LSB per coefficient before
 * quantisation); checked against its own restatement oracle/gateware.py. */
#define B2J_FUSED_DCT_FIXED 0x100
/* [test] OR-ed into B2J_FUSED_PAIR: widen the error interval of the single-precision path from 1/1024 to 1/4, so that
 * nearly every round of blocks takes the exact float64 path (same results; exercises the recomputation). */
#define B2J_FUSED_PAIR_WIDE 0x200
/* The `zrl` argument of b2j_entropy_encode, b2j_encode and b2j_encoder_create is an option word:
 *   bits 0..7    B2J_ZRL_*
 *   bit  8       B2J_OPT_DCT_FIXED (b2j_encode / b2j_encoder_create: K1 uses B2J_FUSED_DCT_FIXED)
 *   bits 12..14  B2J_OPT_RST_PHASE(p): the first restart marker written is RSTp instead of RST0
 *                (a band of a larger frame, encoded on its own: see INTEGRATION.md "one frame on
 *                several GPUs")
 *   bits 16..30  B2J_OPT_RESTART(n): [ext] restart interval of n MCUs, 1..32767 (T.81 B.2.4.4,
 *                E.1.4; the reference has no counterpart): a DRI segment in the
 *                header, and every n MCUs the byte is padded with 1s, RSTm (0xFFD0 + m mod 8)
 *                is written and the DC predictors restart from 0.  Workspaces must then be
 *                sized with the *_workspace_bytes_opts() queries. */
#define B2J_OPT_DCT_FIXED 0x100
#define B2J_OPT_RST_PHASE(p) (((int)(p) & 7) << 12)
#define B2J_OPT_RESTART(n) ((int)(n) << 16)

/* A batch of equally sized, interleaved 8-bit RGB frames (H x W x 3, like the
 * array the model reads at test/model/enc_frame.py:7-10).
 * [ext] Other interleaved pixel layouts are read in place by K1 (no ingest pass): OR
 * B2J_DESC_LAYOUT(layout) into `subsampling`, layout = B2J_LAYOUT_BGR, B2J_LAYOUT_RGBA or B2J_LAYOUT_BGRA
 * (declared with b2j_ingest below; alpha is ignored, as enc_frame.py:8-10 does).  Frames are then
 * H x W x 3 or H x W x 4 bytes and the strides count those bytes.  Results are those of the RGB
 * frames b2j_ingest would produce.  Planar and grey sources still go through b2j_ingest. */
#define B2J_DESC_LAYOUT(layout) ((int32_t)(layout) << 16)
typedef struct b2j_image_desc {
    int32_t width;          /* source pixels, 1..65535 */
    int32_t height;         /* source pixels, 1..65535 */
    int32_t subsampling;    /* B2J_SUB_*, optionally OR-ed with B2J_DESC_LAYOUT(B2J_LAYOUT_BGR | _RGBA | _BGRA) */
    int32_t n_frames;       /* >= 1 */
    int64_t row_stride;     /* bytes between rows, >= 3*width (0 = packed) */
    int64_t frame_stride;   /* bytes between frames, >= height*row_stride (0 = packed) */
} b2j_image_desc;

/* Derived geometry (what the scan order and buffer sizes follow from). */
typedef struct b2j_geometry {
    int32_t h, v;                 /* luma sampling factors */
    int32_t mcus_x, mcus_y;       /* MCUs per row / column */
    int32_t blocks_per_mcu;       /* h*v + 2 */
    int32_t padded_width, padded_height;
    int64_t blocks_per_frame;     /* mcus_x*mcus_y*blocks_per_mcu */
    int32_t header_bytes;         /* SOI..SOS length written in front of the scan: 623; an encode with a restart interval
                                   * adds a 6-byte DRI segment (629), which this field does not include */
    int32_t reserved;
} b2j_geometry;

const char *b2j_version(void);
/* Copies the calling thread's last error message; returns its length. */
size_t b2j_last_error(char *buf, size_t cap);
int b2j_get_geometry(const b2j_image_desc *desc, b2j_geometry *geo);
/* [ext] IJG scaling of the Annex-K tables; quality 85 gives the model's tables
 * (test/model/enc_frame.py:141-159), 50 the gateware's (core/quantization.py:20-27). */
int b2j_quant_tables(int quality, uint8_t qy_host[64], uint8_t qc_host[64]);
/* Host only: the constants of K1's single-precision path (B2J_FUSED_PAIR) for a pair of tables, per natural index
 * x * 8 + u and class (0 luma, 1 chroma): out[0] multipliers r, out[1] and out[2] the offsets of the upper and lower
 * end of the error interval (csrc/b2j_common.cuh QuantF32; restated by tests/fastpath_model.py).  Returns 1 when the
 * tables are supported by that path, 0 when a divisor is 1 (the exact variant is used), < 0 on error.
 * replaces: the divider walk of quantize(), test/model/enc_frame.py:111-129, for that path. */
int b2j_fastpath_constants(const uint8_t qy_host[64], const uint8_t qc_host[64], float out[3][2][64]);

/* ---------------------------------------------------------------- K0: one kernel per stage operator */

/* replaces: RGB2YCbCr datapath, litejpeg/core/csc.py:41-160 (mode XAPP930) and
 * rgb2ycbcr, test/model/enc_frame.py:27-35 (mode JFIF).  rgb: n_pixels x 3
 * interleaved; y/cb/cr: n_pixels each. */
int b2j_rgb2ycbcr(const uint8_t *rgb, uint8_t *y, uint8_t *cb, uint8_t *cr,
                  int64_t n_pixels, int mode, void *stream);

/* replaces: YCbCr444to422 datapath, litejpeg/core/crs.py:8-65.  Streams of
 * n_pixels (even); cb_cr carries mean(Cb) on even and mean(Cr) on odd pixels. */
int b2j_ycbcr444to422(const uint8_t *y, const uint8_t *cb, const uint8_t *cr,
                      uint8_t *y_out, uint8_t *cb_cr_out, int64_t n_pixels, void *stream);

/* Decode-direction operators (they let the reference's round-trip benches
 * test/ycbcr2rgb_tb.py and test/ycbcr_resampling_tb.py run against this library).
 * replaces: YCbCr2RGB datapath, litejpeg/core/csc.py:219-310 (mode XAPP930 = XAPP931 integer
 * pipeline, saturated) and RAWImage.ycbcr2rgb, test/common.py:172-178 (mode JFIF: float64,
 * int(), not clamped - hence int16 outputs). */
int b2j_ycbcr2rgb(const uint8_t *y, const uint8_t *cb, const uint8_t *cr,
                  int16_t *r, int16_t *g, int16_t *b, int64_t n_pixels, int mode, void *stream);

/* replaces: YCbCr422to444 datapath, litejpeg/core/crs.py:88-146 (sample repeat). */
int b2j_ycbcr422to444(const uint8_t *y, const uint8_t *cb_cr,
                      uint8_t *y_out, uint8_t *cb_out, uint8_t *cr_out, int64_t n_pixels, void *stream);

/* [ext] ingest: converts other pixel layouts on the device into the packed,
 * interleaved 8-bit RGB the encode path reads.  replaces: the host-side unpacking of the
 * reference (test/model/enc_frame.py:7-10 `arr[..., 0..2]`; test/common.py:47-51
 * `zip(*img.getdata())`).  Strides in bytes, 0 = packed; src_plane_stride only for
 * B2J_LAYOUT_PLANAR_RGB (R plane, G plane, B plane); rgb: n_frames x height x width x 3. */
#define B2J_LAYOUT_RGB 0
#define B2J_LAYOUT_BGR 1
#define B2J_LAYOUT_RGBA 2       /* alpha ignored, as enc_frame.py:8-10 does for RGBA PNGs */
#define B2J_LAYOUT_BGRA 3
#define B2J_LAYOUT_PLANAR_RGB 4
#define B2J_LAYOUT_GRAY 5       /* r = g = b = sample */
int b2j_ingest(const uint8_t *src, int layout, int32_t width, int32_t height, int32_t n_frames,
               int64_t src_row_stride, int64_t src_plane_stride, int64_t src_frame_stride,
               uint8_t *rgb, void *stream);

/* [ext] PNG ingest on the device.  replaces: PIL decoding the file on the host,
 * test/model/enc_frame.py:7 (`Image.open`) and test/common.py:47-51 (`img.getdata()`).
 * The host walks the chunk list (b2j_png_info: size, samples per pixel and the total IDAT payload of one
 * file held in host memory); b2j_png_decode copies the zlib streams of n_files equally sized files to the
 * device, where one warp per file inflates (RFC 1951: stored, fixed and dynamic blocks) and undoes the
 * scanline filters (None, Sub, Up, Average, Paeth), writing packed RGB frames n_files x height x width x 3
 * (alpha dropped as enc_frame.py:8-10 does, grey replicated) - the input of b2j_fused_block / b2j_encode.
 * Decoded on the device: 8 bits per sample, colour types 0 / 2 / 6 (channels 1 / 3 / 4), non-interlaced;
 * any other file returns B2J_E_SHAPE (decode it with PIL as the reference does, then b2j_ingest).
 * status_host (may be NULL; n_files entries, the call then waits for the stream): 0 = decoded, otherwise
 * 1 bad zlib header, 2 bad Huffman code / distance, 3 more data than the image holds, 4 stream ends early,
 * 5 less data than the image holds, 6 unknown filter type; the frame of such a file is undefined.
 * The call returns after the host-to-device copy of the streams; the kernels run on `stream`. */
int b2j_png_info(const uint8_t *png_host, int64_t png_bytes, int32_t *width, int32_t *height,
                 int32_t *channels, int64_t *zlib_bytes);
size_t b2j_png_workspace_bytes(int32_t width, int32_t height, int32_t channels, int32_t n_files,
                               int64_t max_zlib_bytes);
int b2j_png_decode(const uint8_t *const *png_host, const int64_t *png_bytes, int32_t n_files,
                   int32_t width, int32_t height, int32_t channels, void *workspace,
                   size_t workspace_bytes, uint8_t *rgb, int32_t *status_host, void *stream);

/* [ext] plane subsampling with the same pair mean, horizontally (422) and then
 * vertically (420).  in: height x width, out: (height/v) x (width/2). */
int b2j_chroma_subsample(const uint8_t *in, uint8_t *out, int32_t height, int32_t width,
                         int32_t subsampling, void *stream);

/* replaces: DCT, litejpeg/core/new_dct6.py:400-525 (operator surface) with the
 * arithmetic of dct, test/model/enc_frame.py:42-109 (float64 graph, int(v/8)).
 * in: n_blocks x 64 samples 0..255, row-major; out: n_blocks x 64, [v][u]. */
int b2j_dct(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream);

/* [ext] DCT stage in the gateware's fixed-point arithmetic: DCTDatapath,
 * litejpeg/core/new_dct6.py:107-145 and dct1D :171-396 (see B2J_FUSED_DCT_FIXED). */
int b2j_dct_fixed(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream);

/* replaces: Quantization, litejpeg/core/quantization.py:29-182 and quantize,
 * test/model/enc_frame.py:111-129. */
int b2j_quantize(const int16_t *in, int16_t *out, int64_t n_blocks,
                 const uint8_t table_host[64], int mode, void *stream);

/* replaces: ZigZag, litejpeg/core/zigzag.py:37-212 and zigzag, enc_frame.py:161-165. */
int b2j_zigzag(const int16_t *in, int16_t *out, int64_t n_blocks, void *stream);

/* replaces: RunLength, litejpeg/core/rle/rlecore.py:30-146.  in: n_blocks x 64
 * zig-zag ordered blocks of ONE component in scan order; prev_dc seeds the DC
 * predictor.  words: n_blocks x 64 slots, amp[0:12] run[12:16] valid[16].
 * The 4-bit zero counter starts from 0 in every block (the gateware register of
 * rlecore.py:84-132 persists across blocks and after an EOB; one block at a time,
 * which is all the reference bench drives, the two agree: oracle/gateware.py). */
int b2j_run_length(const int16_t *in, uint32_t *words, int64_t n_blocks, int32_t prev_dc,
                   void *stream);

/* replaces: rle_code, test/model/enc_frame.py:175-176 (the model's generic
 * run-length code, itertools.groupby): (symbol, run) pairs, flattened.
 * pairs: n_blocks x 128 int16 (only the first counts[b] values of row b are
 * written); counts: n_blocks, the number of values (2 x pairs) of each block. */
int b2j_rle_code(const int16_t *in, int16_t *pairs, int32_t *counts, int64_t n_blocks, void *stream);

/* replaces: EntropyCoder, litejpeg/core/rle/entropycoder.py:19-65: size = bit
 * length of |x|. */
int b2j_entropy_size(const int16_t *in, uint8_t *size, int64_t n, void *stream);

/* replaces: RLEMain, litejpeg/core/rle/rlemain.py:24-80: amp[0:12] size[12:16]
 * run[16:20] valid[20]; size is that of the input coefficient of the slot. */
int b2j_rle_main(const int16_t *in, uint32_t *words, int64_t n_blocks, int32_t prev_dc,
                 void *stream);

/* ---------------------------------------------------------------- K1..K4: the fused encode path */

/* K1.  replaces: the loop body test/model/enc_frame.py:203-217
 * (rgb2ycbcr -> dct -> quantize -> zigzag per block) plus [ext] chroma
 * subsampling and edge replication.  coefs: n_frames x blocks_per_frame x 64
 * int16, zig-zag order, blocks in scan (MCU-interleaved) order, DC not yet
 * differenced.  coefs must be 16-byte aligned. */
int b2j_fused_block(const uint8_t *rgb, const b2j_image_desc *desc,
                    const uint8_t qy_host[64], const uint8_t qc_host[64],
                    int16_t *coefs, int variant, void *stream);

/* K2.  replaces: DC differencing enc_frame.py:219-230, RunLength
 * rlecore.py:84-132, EntropyCoder entropycoder.py:55-65, and [ext] the Annex-K
 * code-length lookup.  bits: n_frames x blocks_per_frame Huffman-coded sizes. */
int b2j_block_bits(const int16_t *coefs, const b2j_image_desc *desc, int zrl,
                   int32_t *bits, void *stream);

/* K3.  [ext] exclusive prefix sum of bits within each frame.
 * offsets: n_frames x blocks_per_frame; total_bits: n_frames.
 * workspace: b2j_scan_workspace_bytes(desc) bytes. */
size_t b2j_scan_workspace_bytes(const b2j_image_desc *desc);
int b2j_scan(const int32_t *bits, const b2j_image_desc *desc, uint32_t *offsets,
             int64_t *total_bits, void *workspace, void *stream);

/* K4.  replaces: huffman, enc_frame.py:179-180 (an identity stub) and the
 * header sketch :189-200; [ext] Annex-K Huffman codes at the scanned offsets,
 * MSB-first packing, 1-padding, 0xFF00 stuffing, JFIF header and EOI.
 * out: n_frames slots of out_cap bytes; sizes[f] = bytes of frame f, or minus
 * the bytes needed when out_cap is too small (the call is asynchronous and returns
 * B2J_OK either way: read sizes[]). */
size_t b2j_bitpack_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap);
int b2j_bitpack(const int16_t *coefs, const uint32_t *offsets, const int64_t *total_bits,
                const b2j_image_desc *desc, const uint8_t qy_host[64], const uint8_t qc_host[64],
                int zrl, void *workspace, uint8_t *out, int64_t out_cap, int64_t *sizes,
                void *stream);

/* K2 + K3 + K4 fused into one pass over the coefficients (what b2j_encode uses): per chunk of
 * 256 blocks the sizes are counted, scanned inside the CTA, chained across chunks by a
 * decoupled look-back, and the bits are packed at their final offsets; 0xFF stuffing, JFIF
 * header and EOI follow.  Same outputs as b2j_block_bits + b2j_scan + b2j_bitpack.
 * workspace: b2j_entropy_workspace_bytes() bytes, 256-byte aligned. */
/* Coefficients must be baseline-codable (what K1 produces): |AC| <= 1023 and DC differences within
 * +-2047.  Anything else yields an undefined stream or sizes[] = INT64_MIN, never out-of-bounds writes. */
size_t b2j_entropy_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap);
size_t b2j_entropy_workspace_bytes_opts(const b2j_image_desc *desc, int64_t out_cap, int opts);
int b2j_entropy_encode(const int16_t *coefs, const b2j_image_desc *desc,
                       const uint8_t qy_host[64], const uint8_t qc_host[64], int zrl,
                       void *workspace, uint8_t *out, int64_t out_cap, int64_t *sizes, void *stream);

/* The whole path K1..K4 on device buffers.  workspace: b2j_encode_workspace_bytes().
 * quality 1..100 picks the [ext] IJG tables, or pass quality = 0 and explicit
 * tables.  zrl: B2J_ZRL_*, optionally | B2J_OPT_DCT_FIXED. */
size_t b2j_encode_workspace_bytes(const b2j_image_desc *desc, int64_t out_cap);
size_t b2j_encode_workspace_bytes_opts(const b2j_image_desc *desc, int64_t out_cap, int opts);
int64_t b2j_default_out_cap(const b2j_image_desc *desc);
int b2j_encode(const uint8_t *rgb, const b2j_image_desc *desc, int quality,
               const uint8_t qy_host[64], const uint8_t qc_host[64], int zrl,
               void *workspace, size_t workspace_bytes,
               uint8_t *out, int64_t out_cap, int64_t *sizes, void *stream);

/* Number of kernel launches b2j_encode issues for this geometry (3; one more with restart intervals). */
int b2j_encode_launch_count(const b2j_image_desc *desc);

/* ---------------------------------------------------------------- host-buffer encoder (end-to-end)
 * Owns its device workspace, pinned staging buffers and streams (the only
 * entry points that allocate).  encode_host: rgb_host -> JFIF files in
 * out_host (n_frames slots of out_cap bytes), sizes_host[n_frames].  Copies
 * and kernels of consecutive frame chunks overlap; the call returns when the
 * last byte is in out_host. */
typedef struct b2j_encoder b2j_encoder;
int b2j_encoder_create(b2j_encoder **enc, const b2j_image_desc *desc, int quality, int zrl,
                       int64_t out_cap, int32_t frames_per_chunk);
int b2j_encoder_encode_host(b2j_encoder *enc, const uint8_t *rgb_host, int32_t n_frames,
                            uint8_t *out_host, int64_t *sizes_host);
int b2j_encoder_destroy(b2j_encoder *enc);

/* One call, several GPUs of one box.  replaces: the per-image loop test/model/enc_frame.py:203-238
 * (the unit that shards): frames [g*n/G, (g+1)*n/G) of the host batch are encoded on devices[g]
 * (contiguous shards, no exchange between devices), each by a b2j_encoder driven from its own host
 * thread, so copies and kernels of all devices overlap.  A device may be listed more than once.
 * Buffers as for b2j_encoder_encode_host; pinned host memory must be portable
 * (cudaHostAllocPortable / cudaHostRegisterPortable, or any pinned allocation under unified addressing).
 * b2j_multi_encoder_h2d_probe issues only the host-to-device copies of such a call (`repeats` passes)
 * and reports the wall-clock seconds of one pass: the copy ceiling of the box for this batch, against
 * which an end-to-end rate can be judged. */
typedef struct b2j_multi_encoder b2j_multi_encoder;
int b2j_multi_encoder_create(b2j_multi_encoder **enc, const int32_t *devices, int32_t n_devices,
                             const b2j_image_desc *desc, int quality, int zrl, int64_t out_cap,
                             int32_t frames_per_chunk);
int b2j_multi_encoder_encode_host(b2j_multi_encoder *enc, const uint8_t *rgb_host, int32_t n_frames,
                                  uint8_t *out_host, int64_t *sizes_host);
int b2j_multi_encoder_h2d_probe(b2j_multi_encoder *enc, const uint8_t *rgb_host, int32_t n_frames,
                                int32_t repeats, double *seconds_per_pass);
int b2j_multi_encoder_destroy(b2j_multi_encoder *enc);

#ifdef __cplusplus
}
#endif
#endif /* B2J_H */
