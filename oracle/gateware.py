"""Integer restatements of the gateware stage operators (litejpeg/core/*).

TEST INFRASTRUCTURE (see oracle/__init__.py).  LiteX/Migen is not installed and
not vendored, so the gateware itself cannot be simulated; these follow the
datapath equations in the source.  Pipeline latency / handshake timing is not
modelled: one packet in, one packet out, registers start from reset.
All paths are relative to /root/reference.
"""
import numpy as np


def coef(value, cw=None):
    """litejpeg/core/common.py:91-92."""
    return int(value * 2 ** cw) if cw is not None else value


def rgb2ycbcr_coefs(dw, cw=None):
    """litejpeg/core/csc.py:22-35."""
    return {
        "ca": coef(0.1819, cw), "cb": coef(0.0618, cw),
        "cc": coef(0.6495, cw), "cd": coef(0.5512, cw),
        "yoffset": 2 ** (dw - 4), "coffset": 2 ** (dw - 1),
        "ymax": 2 ** dw - 1, "cmax": 2 ** dw - 1, "ymin": 0, "cmin": 0,
    }


def rgb2ycbcr_gateware(r, g, b, rgb_w=8, ycbcr_w=8, coef_w=8):
    """RGB2YCbCrDatapath, litejpeg/core/csc.py:85-160 (XAPP930 integer pipeline).

    `sig[coef_w:]` on a signed Migen signal is an arithmetic right shift once
    the result is re-interpreted in the (wide enough) signed target signal."""
    c = rgb2ycbcr_coefs(ycbcr_w, coef_w)
    r = np.asarray(r).astype(np.int64)
    g = np.asarray(g).astype(np.int64)
    b = np.asarray(b).astype(np.int64)
    r_minus_g = r - g                                   # :87-92
    b_minus_g = b - g
    carg_plus_cbbg = r_minus_g * c["ca"] + b_minus_g * c["cb"]   # :96-108
    yraw = (carg_plus_cbbg >> coef_w) + g               # :112-115
    b_minus_yraw = b - yraw                             # :120-127
    r_minus_yraw = r - yraw
    y = yraw + c["yoffset"]                             # :144-152
    cb = ((b_minus_yraw * c["cc"]) >> coef_w) + c["coffset"]
    cr = ((r_minus_yraw * c["cd"]) >> coef_w) + c["coffset"]
    return (np.clip(y, c["ymin"], c["ymax"]),           # :156-160 saturate
            np.clip(cb, c["cmin"], c["cmax"]),
            np.clip(cr, c["cmin"], c["cmax"]))


def rgb2ycbcr_gateware_model(r, g, b, dw=8):
    """test/common.py:133-142 `RAWImage.rgb2ycbcr_model`: float coefs, int(), no clamp."""
    c = rgb2ycbcr_coefs(dw)
    r = np.asarray(r).astype(np.float64)
    g = np.asarray(g).astype(np.float64)
    b = np.asarray(b).astype(np.float64)
    yraw = c["ca"] * (r - g) + c["cb"] * (b - g) + g
    y = np.trunc(yraw + c["yoffset"]).astype(np.int64)
    cb = np.trunc(c["cc"] * (b - yraw) + c["coffset"]).astype(np.int64)
    cr = np.trunc(c["cd"] * (r - yraw) + c["coffset"]).astype(np.int64)
    return y, cb, cr


def ycbcr2rgb_coefs(dw, cw=None):
    """litejpeg/core/csc.py:194-216."""
    ca, cb, cc, cd = 0.1819, 0.0618, 0.6495, 0.5512
    xcoef_w = None if cw is None else cw - 2
    return {
        "ca": coef(ca, cw), "cb": coef(cb, cw), "cc": coef(cc, cw), "cd": coef(cd, cw),
        "yoffset": 2 ** (dw - 4), "coffset": 2 ** (dw - 1),
        "ymax": 2 ** dw - 1, "cmax": 2 ** dw - 1, "ymin": 0, "cmin": 0,
        "acoef": coef(1 / cd, xcoef_w), "bcoef": coef(-cb / (cc * (1 - ca - cb)), xcoef_w),
        "ccoef": coef(-ca / (cd * (1 - ca - cb)), xcoef_w), "dcoef": coef(1 / cc, xcoef_w),
    }


def ycbcr2rgb_gateware(y, cb, cr, ycbcr_w=8, rgb_w=8, coef_w=8):
    """YCbCr2RGBDatapath, litejpeg/core/csc.py:219-310 (XAPP931): products sliced [coef_w-2:]
    (arithmetic shift once re-interpreted in the signed sum), saturated to 0..2^rgb_w-1."""
    c = ycbcr2rgb_coefs(rgb_w, coef_w)
    sh = coef_w - 2
    y = np.asarray(y).astype(np.int64)
    cb = np.asarray(cb).astype(np.int64)
    cr = np.asarray(cr).astype(np.int64)
    cbm, crm, ym = cb - c["coffset"], cr - c["coffset"], y - c["yoffset"]     # :266-283
    r = ym + ((crm * c["acoef"]) >> sh)                                       # :291-297
    g = ym + ((cbm * c["bcoef"]) >> sh) + ((crm * c["ccoef"]) >> sh)
    b = ym + ((cbm * c["dcoef"]) >> sh)
    hi = 2 ** rgb_w - 1
    return np.clip(r, 0, hi), np.clip(g, 0, hi), np.clip(b, 0, hi)            # :301-305


def ycbcr2rgb_gateware_model(y, cb, cr, dw=8):
    """test/common.py:158-168 `RAWImage.ycbcr2rgb_model`: float coefficients, int(), no clamp."""
    c = ycbcr2rgb_coefs(dw)
    y = np.asarray(y).astype(np.float64)
    cb = np.asarray(cb).astype(np.float64)
    cr = np.asarray(cr).astype(np.float64)
    r = np.trunc(y - c["yoffset"] + (cr - c["coffset"]) * c["acoef"])
    g = np.trunc(y - c["yoffset"] + (cb - c["coffset"]) * c["bcoef"] + (cr - c["coffset"]) * c["ccoef"])
    b = np.trunc(y - c["yoffset"] + (cb - c["coffset"]) * c["dcoef"])
    return r.astype(np.int64), g.astype(np.int64), b.astype(np.int64)


def ycbcr2rgb_jfif(y, cb, cr):
    """test/common.py:172-178 `RAWImage.ycbcr2rgb` (the Wikipedia inverse): float64, int(), no clamp."""
    y = np.asarray(y).astype(np.float64)
    cb = np.asarray(cb).astype(np.float64)
    cr = np.asarray(cr).astype(np.float64)
    r = np.trunc(y + (cr - 128) * 1.402)
    g = np.trunc(y + (cb - 128) * -0.34414 + (cr - 128) * -0.71414)
    b = np.trunc(y + (cb - 128) * 1.772)
    return r.astype(np.int64), g.astype(np.int64), b.astype(np.int64)


def ycbcr422to444(y, cb_cr):
    """YCbCr422to444Datapath, litejpeg/core/crs.py:88-128: sample repeat; the even pixel of a
    pair carries Cb, the odd one Cr."""
    y = np.asarray(y).astype(np.int64)
    cc = np.asarray(cb_cr).astype(np.int64)
    assert y.size % 2 == 0
    return y, np.repeat(cc[0::2], 2), np.repeat(cc[1::2], 2)


def ycbcr444to422(y, cb, cr):
    """YCbCr444to422Datapath, litejpeg/core/crs.py:38-65.

    1-D streams of even length.  Output stream (y, cb_cr): even pixel carries
    the pair mean of Cb, odd pixel the pair mean of Cr; mean = (c0 + c1) >> 1."""
    y = np.asarray(y).astype(np.int64)
    cb = np.asarray(cb).astype(np.int64)
    cr = np.asarray(cr).astype(np.int64)
    assert y.size % 2 == 0
    cb_mean = (cb[0::2] + cb[1::2]) >> 1
    cr_mean = (cr[0::2] + cr[1::2]) >> 1
    cb_cr = np.empty_like(y)
    cb_cr[0::2] = cb_mean
    cb_cr[1::2] = cr_mean
    return y, cb_cr


def dct_1d_fixed(v):
    """One pass of `DCTDatapath.dct1D`, litejpeg/core/new_dct6.py:171-396, on int64 arrays v[0..7]
    (any trailing shape), with the datapath's own constants and shift points (:219-230: 1004/1024,
    213/256, 49/128, 25/128, 569/1024, 473/512; :345, :370 sqrt2 = 181/128; every product is
    floor(c*x / 2^s), `>>` on a sign-extended signal) but
      * the two wiring slips of the source corrected so that the graph is the model's
        (enc_frame.py:54-84): :262-263 feeds vector8[4] instead of vector8[7] into the sine product
        of the (4,7) rotation, and :294 subtracts vector84Finalcos where v4*sin belongs;
      * sums kept in wide integers (the 12-bit signals of the source wrap when |mean-128| > 32).
    This is the [ext] "fixed" DCT mode of SURVEY 8f-3: NOT parity with the float64 model."""
    v = [np.asarray(x).astype(np.int64) for x in v]

    def bf(i, j):
        v[i], v[j] = v[i] + v[j], v[i] - v[j]
    bf(3, 4); bf(2, 5); bf(1, 6); bf(0, 7)                                  # stage 1  :192-195
    bf(0, 3); bf(1, 2)                                                      # stage 2  :201-202
    v4, v7, v5, v6 = v[4], v[7], v[5], v[6]
    v[4] = ((213 * v4) >> 8) + ((569 * v7) >> 10)                           # :245-263, :291 (corrected)
    v[7] = ((213 * v7) >> 8) - ((569 * v4) >> 10)                           # :294 (corrected)
    v[5] = ((1004 * v5) >> 10) + ((25 * v6) >> 7)                           # :265-293
    v[6] = ((1004 * v6) >> 10) - ((25 * v5) >> 7)
    bf(7, 5); bf(4, 6); bf(0, 1)                                            # stage 3  :300-302
    fa = ((49 * v[2]) >> 7) + ((473 * v[3]) >> 9)                           # :304-336
    fs = ((49 * v[3]) >> 7) - ((473 * v[2]) >> 9)
    v[2], v[3] = (181 * fa) >> 7, (181 * fs) >> 7                           # :338-360
    r = [None] * 8                                                          # stage 4  :364-396
    r[1], r[7] = v[7] + v[4], v[7] - v[4]
    r[5], r[3] = (181 * v[6]) >> 7, (181 * v[5]) >> 7
    r[6], r[2], r[4], r[0] = v[3], v[2], v[1], v[0]
    return r


def dct_fixed(blocks):
    """`DCTDatapath`, new_dct6.py:107-145, with dct_1d_fixed: level shift -128 (:111), dct1D on the
    rows (:114-115), on the columns (:118-127), arithmetic >> 3 (:136-140).  blocks (..., 64)
    samples 0..255 row-major -> (..., 64) int64."""
    b = np.asarray(blocks).astype(np.int64)
    shp = b.shape
    m = b.reshape(-1, 8, 8) - 128
    rows = dct_1d_fixed([m[:, :, k] for k in range(8)])                    # along each row
    m = np.stack(rows, axis=2)
    cols = dct_1d_fixed([m[:, k, :] for k in range(8)])                    # along each column
    m = np.stack(cols, axis=1)
    return (m >> 3).reshape(shp)


# litejpeg/core/quantization.py:20-27 (Annex-K luminance, unscaled)
quant_values = np.array([
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)


def quantize_gateware(block, table=quant_values):
    """Quantization, litejpeg/core/quantization.py:58, 124-168.

    inv = int(2**16/q); m = |x|*inv; out = sign * (m[16:28] + m[15])."""
    x = np.asarray(block).astype(np.int64)
    inv = (2 ** 16 / np.asarray(table).astype(np.float64)).astype(np.int64)
    m = np.abs(x) * inv
    mag = ((m >> 16) & 0xFFF) + ((m >> 15) & 1)
    return np.where(x < 0, -mag, mag)


def run_length(block, prev_dc=0):
    """RLEDatapath, litejpeg/core/rle/rlecore.py:84-132.

    block: 64 ints in zig-zag order.  Returns (symbols, dc) where symbols is a
    list of 64 slots `(valid, run, amp)` aligned with the input index, and dc
    the value to carry as prev_dc for the next block of the same component.
    The 4-bit zero counter is taken as 0 at the start of each block (the
    register is not cleared at write_cnt==0 in the source; a single-packet
    bench starting from reset cannot observe the difference)."""
    block = [int(v) for v in block]
    slots = []
    zero_count = 0
    for idx, v in enumerate(block):
        if idx == 0:                                   # :86-98
            slots.append((1, 0, v - prev_dc))
        elif v == 0:
            if zero_count == 15:                       # :101-105 eager ZRL
                slots.append((1, 15, 0))
                zero_count = 0                         # 4-bit wrap of 15+1
            elif idx == 63:                            # :107-114 EOB
                slots.append((1, 0, 0))
            else:                                      # :119-120
                zero_count += 1
                slots.append((0, 0, 0))
        else:                                          # :127-131
            slots.append((1, zero_count, v))
            zero_count = 0
    return slots, block[0]


def entropy_size(x):
    """EntropyDatapath, litejpeg/core/rle/entropycoder.py:55-65, 142-144:
    number of bits of |x| (0 for 0)."""
    return int(abs(int(x))).bit_length()


def pack_rle_word(valid, run, amp):
    """17-bit RunLength word, rlecore.py:139-146: amp[0:12] run[12:16] valid[16]."""
    return (amp & 0xFFF) | ((run & 0xF) << 12) | ((valid & 1) << 16)


def pack_rlemain_word(valid, run, size, amp):
    """21-bit RLEMain word, rlemain.py:63-74: amp[0:12] size[12:16] run[16:20] valid[20]."""
    return (amp & 0xFFF) | ((size & 0xF) << 12) | ((run & 0xF) << 16) | ((valid & 1) << 20)


def rle_main(block, prev_dc=0):
    """RLEMain, litejpeg/core/rle/rlemain.py:57-80: RunLength + EntropyCoder on
    the same input, merged per slot.  The size field is taken from the *input*
    coefficient of that slot (the EntropyCoder never sees the DC difference)."""
    slots, dc = run_length(block, prev_dc)
    words = []
    for (valid, run, amp), v in zip(slots, block):
        words.append(pack_rlemain_word(valid, run, entropy_size(v), amp))
    return words, dc


def to_signed12(x):
    """test/common.py:353-361 decoder for 12-bit two's complement words."""
    x = int(x) & 0xFFF
    return x - 4096 if x & 0x800 else x
