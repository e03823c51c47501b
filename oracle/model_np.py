"""NumPy float64 restatement of the reference software model.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function follows
/root/reference/test/model/enc_frame.py at the cited lines, in the same
operation order, so that float64 rounding is reproduced bit for bit.  NumPy
ufuncs evaluate one IEEE operation per call, hence no FMA contraction.

Block convention (enc_frame.py:6-24): blocks are 64-element row-major lists,
blocks of an image are row-major too; image sides must be multiples of 8.
"""
import math

import numpy as np

# enc_frame.py:38-40
pi = 3.14159265358979323846
sqrt2 = 1.41421356237309504880

# enc_frame.py:131-139 (same table as litejpeg/core/zigzag.py:19-27)
zig_zag_order = np.array([
    0, 1, 5, 6, 14, 15, 27, 28,
    2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43,
    9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,
    20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,
    35, 36, 48, 49, 57, 58, 62, 63], dtype=np.int64)

# natural index of the coefficient that lands at zig-zag position i
# (litejpeg/core/zigzag.py:30-34 `zigzag_rom`)
zigzag_rom = np.argsort(zig_zag_order)

# enc_frame.py:141-159
y_quant_table = np.array([
    0x05, 0x03, 0x03, 0x05, 0x07, 0x0c, 0x0f, 0x12,
    0x04, 0x04, 0x04, 0x06, 0x08, 0x11, 0x12, 0x11,
    0x04, 0x04, 0x05, 0x07, 0x0c, 0x11, 0x15, 0x11,
    0x04, 0x05, 0x07, 0x09, 0x0f, 0x1a, 0x18, 0x13,
    0x05, 0x07, 0x0b, 0x11, 0x14, 0x21, 0x1f, 0x17,
    0x07, 0x0b, 0x11, 0x13, 0x18, 0x1f, 0x22, 0x1c,
    0x0f, 0x13, 0x17, 0x1a, 0x1f, 0x24, 0x24, 0x1e,
    0x16, 0x1c, 0x1d, 0x1d, 0x22, 0x1e, 0x1f, 0x1e], dtype=np.int64)

cbcr_quant_table = np.array([
    0x05, 0x05, 0x07, 0x0e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x05, 0x06, 0x08, 0x14, 0x1e, 0x1e, 0x1e, 0x1e,
    0x07, 0x08, 0x11, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x0e, 0x14, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e,
    0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e, 0x1e], dtype=np.int64)


def dct_constants():
    """cos/sin exactly as enc_frame.py:43-44 evaluates them: math.cos((n*pi)/16)."""
    out = {}
    for n in (1, 3, 6):
        out["c%d" % n] = math.cos((n * pi) / 16)
        out["s%d" % n] = math.sin((n * pi) / 16)
    out["sqrt2"] = sqrt2
    return out


def image_to_blocks(plane):
    """enc_frame.py:6-24 block split of one H x W plane -> (nblk, 64), row-major."""
    h, w = plane.shape
    assert h % 8 == 0 and w % 8 == 0
    return (plane.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)
            .reshape(-1, 64))


def blocks_to_image(blocks, h, w):
    return (blocks.reshape(h // 8, w // 8, 8, 8).transpose(0, 2, 1, 3)
            .reshape(h, w))


def rgb2ycbcr(r, g, b):
    """enc_frame.py:27-35.  Any-shape integer arrays -> int64 arrays (y, cb, cr).

    Left-to-right float64 evaluation then int() (truncation toward zero)."""
    r = np.asarray(r).astype(np.float64)
    g = np.asarray(g).astype(np.float64)
    b = np.asarray(b).astype(np.float64)
    y = 0.299 * r
    y = y + 0.587 * g
    y = y + 0.114 * b
    cb = -0.1687 * r
    cb = cb - 0.3313 * g
    cb = cb + 0.5 * b
    cb = cb + 128
    cr = 0.5 * r
    cr = cr - 0.4187 * g
    cr = cr - 0.0813 * b
    cr = cr + 128
    return (np.trunc(y).astype(np.int64), np.trunc(cb).astype(np.int64),
            np.trunc(cr).astype(np.int64))


def _rotation(i0, i1, k, n):
    # enc_frame.py:42-47
    cos = math.cos((n * pi) / 16)
    sin = math.sin((n * pi) / 16)
    out0 = k * (i0 * cos + i1 * sin)
    out1 = k * (i1 * cos - i0 * sin)
    return out0, out1


def _butterfly(i0, i1):
    # enc_frame.py:49-52
    return i0 + i1, i0 - i1


def dct_1d(v):
    """enc_frame.py:54-84.  v: list of 8 float64 arrays (broadcast over blocks)."""
    v = list(v)
    # stage 1
    v[3], v[4] = _butterfly(v[3], v[4])
    v[2], v[5] = _butterfly(v[2], v[5])
    v[1], v[6] = _butterfly(v[1], v[6])
    v[0], v[7] = _butterfly(v[0], v[7])
    # stage 2
    v[0], v[3] = _butterfly(v[0], v[3])
    v[1], v[2] = _butterfly(v[1], v[2])
    v[4], v[7] = _rotation(v[4], v[7], 1, 3)
    v[5], v[6] = _rotation(v[5], v[6], 1, 1)
    # stage 3
    v[7], v[5] = _butterfly(v[7], v[5])
    v[4], v[6] = _butterfly(v[4], v[6])
    v[0], v[1] = _butterfly(v[0], v[1])
    v[2], v[3] = _rotation(v[2], v[3], sqrt2, 6)
    # stage 4
    r = [None] * 8
    r[1], r[7] = _butterfly(v[7], v[4])
    r[5] = sqrt2 * v[6]
    r[3] = sqrt2 * v[5]
    r[6] = v[3]
    r[2] = v[2]
    r[4] = v[1]
    r[0] = v[0]
    return r


def dct(blocks):
    """enc_frame.py:86-109.  (..., 64) ints in 0..255 -> (..., 64) int64."""
    blocks = np.asarray(blocks)
    shp = blocks.shape
    m = blocks.reshape(-1, 8, 8).astype(np.float64) - 128  # :94 level shift
    rows = dct_1d([m[:, :, y] for y in range(8)])           # :92-100 row pass
    m = np.stack(rows, axis=2)
    cols = dct_1d([m[:, x, :] for x in range(8)])           # :103-108 column pass
    m = np.stack(cols, axis=1)
    return np.trunc(m / 8).astype(np.int64).reshape(shp)     # :108 int(v/8)


def quantize(blocks, table):
    """enc_frame.py:111-129: round-half-away-from-zero of value/divider.

    Integer restatement sign(v)*((2|v|+d)//(2d)); make_golden.py checks it
    against the model's float/ceil code for every d in 1..255, v in -2048..2047."""
    v = np.asarray(blocks).astype(np.int64)
    d = np.asarray(table).astype(np.int64)
    a = (2 * np.abs(v) + d) // (2 * d)
    return np.where(v < 0, -a, a)


def quantize_trunc(blocks, table):
    """Older model test/model/encoder.py:84-88: int(value/divider)."""
    v = np.asarray(blocks).astype(np.int64)
    d = np.asarray(table).astype(np.int64)
    a = np.abs(v) // d
    return np.where(v < 0, -a, a)


def zigzag(blocks):
    """enc_frame.py:161-165: out[zig_zag_order[i]] = in[i]."""
    blocks = np.asarray(blocks)
    out = np.empty_like(blocks)
    out[..., zig_zag_order] = blocks
    return out


def dc_diff(zz_blocks):
    """enc_frame.py:202, 219-230: per-component predictor, initial 0, never reset.

    zz_blocks: (nblk, 64) blocks of ONE component in scan order."""
    out = np.array(zz_blocks, dtype=np.int64, copy=True)
    dc = out[:, 0].copy()
    out[1:, 0] = dc[1:] - dc[:-1]
    return out


def rle_code(block):
    """enc_frame.py:175-176: generic (symbol, run) pairs, flattened."""
    block = [int(x) for x in block]
    out = []
    i = 0
    while i < len(block):
        j = i
        while j < len(block) and block[j] == block[i]:
            j += 1
        out += [block[i], j - i]
        i = j
    return out


def encode_planes_444(rgb, ytab=y_quant_table, ctab=cbcr_quant_table):
    """The loop body enc_frame.py:203-217 over a whole H x W x 3 image.

    Returns zig-zag ordered quantised coefficients before DC differencing:
    three (nblk, 64) int64 arrays (y, cb, cr), blocks row-major."""
    rgb = np.asarray(rgb)
    y, cb, cr = rgb2ycbcr(rgb[..., 0], rgb[..., 1], rgb[..., 2])
    out = []
    for plane, tab in ((y, ytab), (cb, ctab), (cr, ctab)):
        out.append(zigzag(quantize(dct(image_to_blocks(plane)), tab)))
    return tuple(out)
