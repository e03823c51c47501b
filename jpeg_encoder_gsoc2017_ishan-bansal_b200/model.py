"""Function-level mirror of the reference's software model, test/model/enc_frame.py.

Same names, argument meaning and return types as the model (lists of 64 ints per
block), so a script written against `enc_frame` runs against this module
unchanged - but every stage executes on the GPU through the C ABI of
include/b2j.h.  Each function also accepts a *batch*: a list of blocks (or a
tensor (n, 64)) in, a list of blocks out; that is how the functions are meant to
be used at speed (`encode_blocks` runs the whole loop body enc_frame.py:203-238
for all blocks of an image in a handful of launches).

There is no CPU fallback: without the CUDA library these functions raise.
"""
import torch

from . import _lib
from .common import require_cuda, stream_ptr

# y_quant_table / cbcr_quant_table (enc_frame.py:141-159: the Annex-K tables scaled to IJG quality 85)
# are module attributes served by __getattr__ below from the library's b2j_quant_tables(85).
_TABLES = None

# enc_frame.py:150-159 zig_zag_order[i] = position of natural-order coefficient i in the zig-zag sequence
zig_zag_order = [
    0, 1, 5, 6, 14, 15, 27, 28,
    2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43,
    9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,
    20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,
    35, 36, 48, 49, 57, 58, 62, 63]


def _tables():
    global _TABLES
    if _TABLES is None:
        qy, qc = _lib.quant_tables(85)
        _TABLES = (list(qy), list(qc))
    return _TABLES


def __getattr__(name):            # y_quant_table / cbcr_quant_table come from the library (b2j_quant_tables)
    if name in ("y_quant_table", "cbcr_quant_table"):
        t = _tables()
        return t[0] if name == "y_quant_table" else t[1]
    raise AttributeError(name)


def _blocks(x, dtype):
    """list of 64 | list of lists of 64 | tensor (..., 64)  ->  (CUDA tensor (n, 64), was_single, was_tensor)."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x.to(device="cuda", dtype=dtype).contiguous()
        if t.numel() == 0 or t.shape[-1] != 64:
            raise ValueError("blocks must have 64 values")
        return t.reshape(-1, 64), t.dim() == 1, True
    x = list(x)
    single = len(x) > 0 and not hasattr(x[0], "__len__")
    rows = [x] if single else x
    if not rows or any(len(r) != 64 for r in rows):
        raise ValueError("blocks must have 64 values")
    return torch.tensor([[int(v) for v in r] for r in rows], dtype=torch.int64).to(dtype).cuda(), single, False


def _ret(t, single, tensor):
    if tensor:
        return t[0] if single else t
    rows = t.cpu().tolist()
    return rows[0] if single else rows


def readRgbImageBlocks(name):
    """enc_frame.py:6-24: image file -> (r_blocks, g_blocks, b_blocks), blocks row-major, 64 pixels
    row-major each.  Like the model, dimensions are floored to multiples of 8 blocks (int(W/8))."""
    from PIL import Image
    import numpy as np
    arr = np.array(Image.open(name))
    h8, w8 = arr.shape[0] // 8, arr.shape[1] // 8
    a = arr[:h8 * 8, :w8 * 8, :3].reshape(h8, 8, w8, 8, 3).transpose(4, 0, 2, 1, 3).reshape(3, h8 * w8, 64)
    return tuple(a[c].tolist() for c in range(3))


def rgb2ycbcr(r_block, g_block, b_block):
    """enc_frame.py:27-35, the JFIF float64 conversion with int() truncation -> (y, cb, cr)."""
    r, single, tensor = _blocks(r_block, torch.uint8)
    g, _, _ = _blocks(g_block, torch.uint8)
    b, _, _ = _blocks(b_block, torch.uint8)
    if r.shape != g.shape or r.shape != b.shape:
        raise ValueError("r, g and b must have the same shape")
    rgb = torch.stack([r, g, b], dim=-1).contiguous()
    n = r.numel()
    planes = torch.empty((3, n), dtype=torch.uint8, device=rgb.device)
    _lib.check(_lib.lib().b2j_rgb2ycbcr(rgb.data_ptr(), planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(),
                                         n, _lib.CSC["jfif"], stream_ptr(None)))
    conv = (lambda p: p.reshape(r.shape)) if tensor else (lambda p: p.reshape(r.shape).to(torch.int64))
    return tuple(_ret(conv(planes[c]), single, tensor) for c in range(3))


def dct(block):
    """enc_frame.py:86-109: level shift, Loeffler row and column passes in float64, int(v/8)."""
    t, single, tensor = _blocks(block, torch.int16)
    out = torch.empty_like(t)
    _lib.check(_lib.lib().b2j_dct(t.data_ptr(), out.data_ptr(), t.shape[0], stream_ptr(None)))
    return _ret(out, single, tensor)


def quantize(block, table):
    """enc_frame.py:111-129: round half away from zero of value / divider (table in natural order)."""
    t, single, tensor = _blocks(block, torch.int16)
    table = [int(v) for v in table]
    if len(table) != 64 or min(table) < 1 or max(table) > 255:
        raise ValueError("table must hold 64 divisors in 1..255")
    out = torch.empty_like(t)
    _lib.check(_lib.lib().b2j_quantize(t.data_ptr(), out.data_ptr(), t.shape[0], _lib.table_arg(table),
                                        _lib.QMODE["round"], stream_ptr(None)))
    return _ret(out, single, tensor)


def zigzag(block):
    """enc_frame.py:161-165: r[zig_zag_order[i]] = block[i]."""
    t, single, tensor = _blocks(block, torch.int16)
    out = torch.empty_like(t)
    _lib.check(_lib.lib().b2j_zigzag(t.data_ptr(), out.data_ptr(), t.shape[0], stream_ptr(None)))
    return _ret(out, single, tensor)


def rle_code(block):
    """enc_frame.py:175-176: generic run-length code, flattened (symbol, run) pairs (DC included)."""
    t, single, _ = _blocks(block, torch.int16)
    pairs = torch.empty((t.shape[0], 128), dtype=torch.int16, device=t.device)
    counts = torch.empty((t.shape[0],), dtype=torch.int32, device=t.device)
    _lib.check(_lib.lib().b2j_rle_code(t.data_ptr(), pairs.data_ptr(), counts.data_ptr(), t.shape[0], stream_ptr(None)))
    pairs, counts = pairs.cpu().tolist(), counts.cpu().tolist()
    out = [row[:n] for row, n in zip(pairs, counts)]
    return out[0] if single else out


def huffman(block):
    """enc_frame.py:179-180: the model's Huffman stage is the identity (the [ext] Annex-K coder is
    `litejpeg_b200.Huffman` / `JpegEncoder`)."""
    return block


def encode_blocks(r_blocks, g_blocks, b_blocks):
    """The loop enc_frame.py:200-238 over all blocks of an image: per component, zig-zag ordered
    quantised blocks with the DC replaced by its difference to the previous block (predictors
    start at 0), and their rle_code.  Returns ((y_zig, cb_zig, cr_zig), (y_rle, cb_rle, cr_rle))."""
    ty, tc = _tables()
    y, cb, cr = rgb2ycbcr(r_blocks, g_blocks, b_blocks)
    zigs, rles = [], []
    for comp, table in ((y, ty), (cb, tc), (cr, tc)):
        z = zigzag(quantize(dct(torch.tensor(comp, dtype=torch.int16)), table)).to(torch.int64)
        dc = z[:, 0].clone()
        z[1:, 0] = dc[1:] - dc[:-1]
        zigs.append(z.cpu().tolist())
        rles.append(rle_code(z.to(torch.int16)))
    return tuple(zigs), tuple(rles)
