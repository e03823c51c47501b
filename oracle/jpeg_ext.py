"""[ext] stages the reference never implemented: chroma subsampling to 4:2:2 /
4:2:0, IJG quality scaling, JPEG-style RLE symbols, Annex-K Huffman coding,
bit packing with 0xFF stuffing, the JFIF container, and a baseline decoder.

TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: /root/reference has no implementation and no test of anything
in this file (`huffman` is `return block`, test/model/enc_frame.py:179-180; the
header sketch :189-200 cannot run).  What anchors the definitions instead:

* RLE symbol rules follow the only JPEG-style RLE in the tree, the gateware
  `RLEDatapath` (litejpeg/core/rle/rlecore.py:84-132): eager ZRL (15,0) on
  every 16th consecutive zero, EOB (0,0) when index 63 is zero (and did not
  just complete a ZRL).  `zrl="standard"` gives the libjpeg policy instead.
* size category = bit length of |x| (litejpeg/core/rle/entropycoder.py:55-65).
* DC prediction: one predictor per component, initial 0, never reset
  (enc_frame.py:202, 219-230).
* 4:2:2 = horizontal pair mean (c0+c1)>>1 (litejpeg/core/crs.py:39-54);
  4:2:0 cascades the same mean horizontally then vertically.
* quality q = IJG scaling of the Annex-K tables; q=85 reproduces the model's
  tables exactly (enc_frame.py:141-159), q=50 the gateware's
  (litejpeg/core/quantization.py:20-27).
* Huffman tables, additional-bit format, MSB-first packing, 1-padding, 0xFF00
  stuffing and marker layout are ITU-T T.81 Annex K / B / F; checked by
  decoding the files with this module's decoder, PIL and OpenCV.
* Edge padding to a whole MCU replicates the last column / row of RGB.
"""
import struct

import numpy as np

from . import model_np as M

# ---------------------------------------------------------------- tables

# Annex K.1 luminance (== litejpeg/core/quantization.py:20-27, test/common.py:319-326)
BASE_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)

# Annex K.2 chrominance (== test/common.py:329-336)
BASE_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64)


def ijg_quant_table(base, quality):
    """IJG jcparam.c scaling: s = 5000//q (q<50) else 200-2q; clamp((b*s+50)//100, 1, 255)."""
    q = int(quality)
    if not 1 <= q <= 100:
        raise ValueError("quality must be in 1..100")
    s = 5000 // q if q < 50 else 200 - 2 * q
    return np.clip((np.asarray(base, dtype=np.int64) * s + 50) // 100, 1, 255)


def quant_tables(quality):
    return ijg_quant_table(BASE_LUMA, quality), ijg_quant_table(BASE_CHROMA, quality)


# Annex K.3 typical Huffman tables: (BITS[1..16], HUFFVAL)
DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d], [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])
AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77], [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])


def derive_codes(bits, vals):
    """T.81 Annex C: canonical code assignment.  Returns (code[256], length[256])."""
    code = np.zeros(256, dtype=np.int64)
    length = np.zeros(256, dtype=np.int64)
    c = 0
    k = 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            code[vals[k]] = c
            length[vals[k]] = ln
            c += 1
            k += 1
        c <<= 1
    assert k == len(vals)
    return code, length


HUFF = {name: derive_codes(*tab) for name, tab in
        (("dc_luma", DC_LUMA), ("dc_chroma", DC_CHROMA),
         ("ac_luma", AC_LUMA), ("ac_chroma", AC_CHROMA))}

# ---------------------------------------------------------------- geometry

SUBSAMPLING = {"444": (1, 1), "422": (2, 1), "420": (2, 2)}   # (h, v) of luma


def geometry(height, width, subsampling):
    h, v = SUBSAMPLING[subsampling]
    mcu_w, mcu_h = 8 * h, 8 * v
    mcus_x = (width + mcu_w - 1) // mcu_w
    mcus_y = (height + mcu_h - 1) // mcu_h
    return dict(h=h, v=v, mcu_w=mcu_w, mcu_h=mcu_h, mcus_x=mcus_x, mcus_y=mcus_y,
                padded_w=mcus_x * mcu_w, padded_h=mcus_y * mcu_h,
                blocks_per_mcu=h * v + 2, n_mcu=mcus_x * mcus_y)


def pad_rgb(rgb, geo):
    """Replicate the last column / row up to a whole number of MCUs."""
    hgt, wid = rgb.shape[:2]
    return np.pad(rgb, ((0, geo["padded_h"] - hgt), (0, geo["padded_w"] - wid), (0, 0)),
                  mode="edge")


def subsample(cb, cr, subsampling):
    """Pair mean (c0+c1)>>1 (crs.py:39-54) horizontally, then (4:2:0) vertically."""
    h, v = SUBSAMPLING[subsampling]
    out = []
    for c in (cb, cr):
        c = np.asarray(c).astype(np.int64)
        if h == 2:
            c = (c[:, 0::2] + c[:, 1::2]) >> 1
        if v == 2:
            c = (c[0::2, :] + c[1::2, :]) >> 1
        out.append(c)
    return out


def mcu_order_blocks(yb, cbb, crb, geo):
    """Arrange per-component row-major block arrays into scan order:
    (n_mcu, blocks_per_mcu, 64); luma blocks of an MCU are raster within it."""
    h, v = geo["h"], geo["v"]
    mx, my = geo["mcus_x"], geo["mcus_y"]
    yb = yb.reshape(my, v, mx, h, 64).transpose(0, 2, 1, 3, 4).reshape(my * mx, v * h, 64)
    return np.concatenate([yb, cbb.reshape(-1, 1, 64), crb.reshape(-1, 1, 64)], axis=1)


def component_of_block(geo):
    """0 (Y) / 1 (Cb) / 2 (Cr) for each block slot of an MCU."""
    n = geo["h"] * geo["v"]
    return np.array([0] * n + [1, 2])


def encode_coefficients(rgb, subsampling="444", qy=None, qc=None, quality=None):
    """RGB (H, W, 3) u8 -> quantised zig-zag coefficients in scan order
    (n_mcu, blocks_per_mcu, 64) int16, DC *not* yet differenced.

    Colour, DCT, quantise, zig-zag are the model's (oracle.model_np)."""
    rgb = np.asarray(rgb)
    if quality is not None:
        qy, qc = quant_tables(quality)
    if qy is None:
        qy, qc = M.y_quant_table, M.cbcr_quant_table
    geo = geometry(rgb.shape[0], rgb.shape[1], subsampling)
    p = pad_rgb(rgb, geo)
    y, cb, cr = M.rgb2ycbcr(p[..., 0], p[..., 1], p[..., 2])
    cb, cr = subsample(cb, cr, subsampling)
    blocks = []
    for plane, tab in ((y, qy), (cb, qc), (cr, qc)):
        blocks.append(M.zigzag(M.quantize(M.dct(M.image_to_blocks(plane)), tab)))
    return mcu_order_blocks(blocks[0], blocks[1], blocks[2], geo).astype(np.int16), geo


# ---------------------------------------------------------------- entropy coding

def _bit_length(a):
    a = np.asarray(a).astype(np.int64)
    out = np.zeros(a.shape, dtype=np.int64)
    t = a.copy()
    for _ in range(16):
        nzm = t > 0
        out += nzm
        t >>= 1
    return out


def dc_differences(coefs, geo, restart_interval=0):
    """DC of every block minus DC of the previous block of the same component
    in scan order (first block: minus 0).  [ext] restart_interval > 0: the
    predictors restart from 0 every `restart_interval` MCUs (T.81 F.1.1.5.1 / E.1.4)."""
    nm, bpm, _ = coefs.shape
    comp = component_of_block(geo)
    dc = coefs[:, :, 0].astype(np.int64)
    diff = np.empty_like(dc)
    for c in range(3):
        slots = np.nonzero(comp == c)[0]
        seq = dc[:, slots].reshape(-1)
        d = seq.copy()
        d[1:] = seq[1:] - seq[:-1]
        d = d.reshape(nm, len(slots))
        if restart_interval > 0:
            first = np.arange(0, nm, restart_interval)
            d[first, 0] = dc[first, slots[0]]                  # first block of the component in the interval
        diff[:, slots] = d
    return diff


def block_symbols(coefs, geo, zrl="eager", restart_interval=0):
    """Per coefficient position, the Huffman-coded bits it emits.

    Returns (value, nbits) uint64/int64 arrays of shape (nblk, 64, 2): slot 0 is
    an optional ZRL prefix (only used by zrl="standard"), slot 1 the main
    symbol (Huffman code followed by the additional bits)."""
    nm, bpm, _ = coefs.shape
    comp = np.tile(component_of_block(geo), nm)               # per block
    is_chroma = comp > 0
    c = coefs.reshape(-1, 64).astype(np.int64)
    nblk = c.shape[0]
    diff = dc_differences(coefs, geo, restart_interval).reshape(-1)
    amp = c.copy()
    amp[:, 0] = diff
    size = _bit_length(np.abs(amp))
    extra = np.where(amp < 0, amp - 1, amp) & ((1 << size) - 1)   # T.81 F.1.2.1/2

    idx = np.arange(64)[None, :]
    nz = c != 0
    anchor = nz.copy()
    anchor[:, 0] = True                                        # DC resets the run
    last = np.maximum.accumulate(np.where(anchor, idx, 0), axis=1)
    zr = idx - last                                            # zeros up to and incl. k (0 at anchors)
    prev_zr = np.concatenate([np.zeros((nblk, 1), np.int64), zr[:, :-1]], axis=1)

    def lut(name_l, name_c, sym):
        cl, ll = HUFF[name_l]
        cc, lc = HUFF[name_c]
        code = np.where(is_chroma[:, None], cc[sym], cl[sym])
        ln = np.where(is_chroma[:, None], lc[sym], ll[sym])
        return code, ln

    val = np.zeros((nblk, 64, 2), dtype=np.uint64)
    nb = np.zeros((nblk, 64, 2), dtype=np.int64)

    # DC symbol
    code, ln = lut("dc_luma", "dc_chroma", size[:, :1])
    val[:, 0, 1] = ((code[:, 0] << size[:, 0]) | extra[:, 0]).astype(np.uint64)
    nb[:, 0, 1] = ln[:, 0] + size[:, 0]

    # AC non-zero coefficients
    run = prev_zr & 15
    rs = np.where(nz, (run << 4) | size, 0)
    code, ln = lut("ac_luma", "ac_chroma", rs)
    acnz = nz.copy()
    acnz[:, 0] = False
    v = ((code << size) | extra).astype(np.uint64)
    val[:, :, 1] = np.where(acnz, v, val[:, :, 1])
    nb[:, :, 1] = np.where(acnz, ln + size, nb[:, :, 1])

    zrl_code, zrl_len = lut("ac_luma", "ac_chroma", np.full((nblk, 1), 0xF0))
    eob_code, eob_len = lut("ac_luma", "ac_chroma", np.full((nblk, 1), 0x00))
    zero = ~nz
    zero[:, 0] = False
    if zrl == "eager":
        # rlecore.py:101-105: every 16th consecutive zero emits (15,0) at once
        is_zrl = zero & (zr % 16 == 0)
        # rlecore.py:107-114: zero at index 63 that is not a ZRL emits (0,0)
        is_eob = zero & (idx == 63) & ~is_zrl
    elif zrl == "standard":
        n_zrl = np.where(acnz, prev_zr >> 4, 0)
        for n in (1, 2, 3):
            rep = np.zeros((nblk, 1), dtype=np.uint64)
            for _ in range(n):
                rep = (rep << zrl_len.astype(np.uint64)) | zrl_code.astype(np.uint64)
            m = n_zrl == n
            val[:, :, 0] = np.where(m, rep, val[:, :, 0])
            nb[:, :, 0] = np.where(m, n * zrl_len, nb[:, :, 0])
        is_zrl = np.zeros_like(zero)
        is_eob = zero & (idx == 63)
    else:
        raise ValueError(zrl)
    val[:, :, 1] = np.where(is_zrl, zrl_code.astype(np.uint64), val[:, :, 1])
    nb[:, :, 1] = np.where(is_zrl, zrl_len, nb[:, :, 1])
    val[:, :, 1] = np.where(is_eob, eob_code.astype(np.uint64), val[:, :, 1])
    nb[:, :, 1] = np.where(is_eob, eob_len, nb[:, :, 1])
    return val, nb


def block_bit_lengths(coefs, geo, zrl="eager"):
    """Huffman-coded size of every block in bits, shape (n_mcu, blocks_per_mcu)."""
    _, nb = block_symbols(coefs, geo, zrl)
    return nb.sum(axis=(1, 2)).reshape(coefs.shape[0], coefs.shape[1])


def pack_bits(val, nb):
    """Concatenate (value, nbits) fields MSB first; pad the last byte with 1s."""
    val = val.reshape(-1)
    nb = nb.reshape(-1)
    keep = nb > 0
    val, nb = val[keep], nb[keep]
    total = int(nb.sum())
    starts = np.cumsum(nb) - nb
    sym = np.repeat(np.arange(nb.size), nb)
    pos = np.arange(total) - starts[sym]
    bits = ((val[sym] >> (nb[sym] - 1 - pos).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
    pad = (-total) % 8
    bits = np.concatenate([bits, np.ones(pad, dtype=np.uint8)])
    return np.packbits(bits), total


def stuff_bytes(data):
    """0xFF -> 0xFF 0x00 (T.81 B.1.1.5)."""
    data = np.asarray(data, dtype=np.uint8)
    ff = data == 0xFF
    pos = np.arange(data.size) + np.cumsum(ff) - ff
    out = np.zeros(data.size + int(ff.sum()), dtype=np.uint8)
    out[pos] = data
    return out


def entropy_encode(coefs, geo, zrl="eager", restart_interval=0):
    """Scan data (stuffed, 1-padded) for coefficients in scan order.

    [ext] restart_interval > 0 (T.81 B.2.4.4, E.1.4): every `restart_interval` MCUs the current
    byte is padded with 1s, a marker RSTm (0xFFD0 + m, m counting modulo 8) is written and the
    DC predictors start again from 0; there is no marker after the last interval.  Returns the
    scan bytes and the number of entropy-coded bits (padding and markers not counted)."""
    val, nb = block_symbols(coefs, geo, zrl, restart_interval)
    if restart_interval <= 0:
        packed, total_bits = pack_bits(val, nb)
        return stuff_bytes(packed).tobytes(), total_bits
    nm, bpm = coefs.shape[0], coefs.shape[1]
    val = val.reshape(nm, bpm, 64, 2)
    nb = nb.reshape(nm, bpm, 64, 2)
    out, total_bits = b"", 0
    for k, m0 in enumerate(range(0, nm, restart_interval)):
        packed, bits = pack_bits(val[m0:m0 + restart_interval], nb[m0:m0 + restart_interval])
        total_bits += bits
        if k:
            out += bytes([0xFF, 0xD0 + ((k - 1) & 7)])
        out += stuff_bytes(packed).tobytes()
    return out, total_bits


# ---------------------------------------------------------------- container

def _seg(marker, payload):
    return struct.pack(">BBH", 0xFF, marker, len(payload) + 2) + payload


def jfif_header(height, width, subsampling, qy, qc, restart_interval=0):
    """SOI, APP0(JFIF 1.01), DQT x2 (zig-zag order), SOF0, DHT x4, [DRI,] SOS."""
    h, v = SUBSAMPLING[subsampling]
    out = b"\xFF\xD8"
    out += _seg(0xE0, b"JFIF\x00" + bytes([1, 1, 0]) + struct.pack(">HH", 1, 1) + bytes([0, 0]))
    for tq, tab in ((0, qy), (1, qc)):
        zz = np.asarray(tab, dtype=np.int64)[M.zigzag_rom]
        out += _seg(0xDB, bytes([tq]) + bytes(int(x) for x in zz))
    out += _seg(0xC0, bytes([8]) + struct.pack(">HH", height, width) + bytes([3]) +
                bytes([1, (h << 4) | v, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for tc_th, tab in ((0x00, DC_LUMA), (0x10, AC_LUMA), (0x01, DC_CHROMA), (0x11, AC_CHROMA)):
        out += _seg(0xC4, bytes([tc_th]) + bytes(tab[0]) + bytes(tab[1]))
    if restart_interval > 0:
        out += _seg(0xDD, struct.pack(">H", restart_interval))
    out += _seg(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def encode_jpeg(rgb, subsampling="444", quality=85, zrl="eager", restart_interval=0):
    """Full RGB -> JFIF bytes."""
    if not 0 <= restart_interval <= 65535:
        raise ValueError("restart_interval must be in 0..65535")
    qy, qc = quant_tables(quality)
    coefs, geo = encode_coefficients(rgb, subsampling, qy, qc)
    scan, _ = entropy_encode(coefs, geo, zrl, restart_interval)
    return jfif_header(rgb.shape[0], rgb.shape[1], subsampling, qy, qc, restart_interval) + scan + b"\xFF\xD9"


# ---------------------------------------------------------------- decoder

class _BitReader:
    def __init__(self, data):
        self.data = data
        self.pos = 0
        self.acc = 0
        self.n = 0

    def restart(self, expect):
        """Discard the 1-padding of the current byte and consume the marker RST<expect>."""
        while self.n:
            assert self.bit() == 1, "non-1 padding before a restart marker"
        assert self.data[self.pos] == 0xFF and self.data[self.pos + 1] == 0xD0 + expect, "missing RST%d" % expect
        self.pos += 2

    def bit(self):
        if self.n == 0:
            b = self.data[self.pos]
            self.pos += 1
            if b == 0xFF:
                nxt = self.data[self.pos]
                if nxt != 0:
                    raise ValueError("marker 0xFF%02X inside scan" % nxt)
                self.pos += 1
            self.acc = b
            self.n = 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


def _decode_table(bits, vals):
    code, length = derive_codes(bits, vals)
    return {(int(length[s]), int(code[s])): s for s in vals}


def decode_jpeg_coefficients(data):
    """Baseline JFIF -> dict(height, width, subsampling, qy, qc, coefs) where coefs
    is (n_mcu, blocks_per_mcu, 64) int16 zig-zag order with DC prediction undone."""
    data = bytes(data)
    assert data[:2] == b"\xFF\xD8"
    pos = 2
    qt = {}
    ht = {}
    frame = None
    restart_interval = 0
    while True:
        assert data[pos] == 0xFF, "expected marker at %d" % pos
        m = data[pos + 1]
        ln = struct.unpack(">H", data[pos + 2:pos + 4])[0]
        seg = data[pos + 4:pos + 2 + ln]
        pos += 2 + ln
        if m == 0xDB:
            i = 0
            while i < len(seg):
                assert seg[i] >> 4 == 0
                zz = np.frombuffer(seg[i + 1:i + 65], dtype=np.uint8).astype(np.int64)
                nat = np.empty(64, dtype=np.int64)
                nat[M.zigzag_rom] = zz
                qt[seg[i] & 15] = nat
                i += 65
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                bits = list(seg[i + 1:i + 17])
                n = sum(bits)
                ht[seg[i]] = _decode_table(bits, list(seg[i + 17:i + 17 + n]))
                i += 17 + n
        elif m == 0xC0:
            prec, hgt, wid, nc = struct.unpack(">BHHB", seg[:6])
            assert prec == 8 and nc == 3
            frame = dict(height=hgt, width=wid,
                         comps=[(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k])
                                for k in range(3)])
        elif m == 0xDD:
            restart_interval = struct.unpack(">H", seg[:2])[0]
        elif m == 0xDA:
            break
    h, v = frame["comps"][0][1], frame["comps"][0][2]
    sub = {(1, 1): "444", (2, 1): "422", (2, 2): "420"}[(h, v)]
    geo = geometry(frame["height"], frame["width"], sub)
    rd = _BitReader(data[pos:])
    comp = component_of_block(geo)
    pred = [0, 0, 0]
    coefs = np.zeros((geo["n_mcu"], geo["blocks_per_mcu"], 64), dtype=np.int16)

    def sym(table):
        ln, code = 0, 0
        while True:
            code = (code << 1) | rd.bit()
            ln += 1
            s = table.get((ln, code))
            if s is not None:
                return s
            assert ln < 16, "bad Huffman code"

    def extend(vv, s):
        return vv if s == 0 or vv >= (1 << (s - 1)) else vv - (1 << s) + 1

    for mcu in range(geo["n_mcu"]):
        if restart_interval and mcu and mcu % restart_interval == 0:
            rd.restart((mcu // restart_interval - 1) & 7)
            pred = [0, 0, 0]
        for b in range(geo["blocks_per_mcu"]):
            c = int(comp[b])
            dct, act = (ht[0x00], ht[0x10]) if c == 0 else (ht[0x01], ht[0x11])
            s = sym(dct)
            pred[c] += extend(rd.bits(s), s)
            coefs[mcu, b, 0] = pred[c]
            k = 1
            while k < 64:
                rs = sym(act)
                r, s = rs >> 4, rs & 15
                if s == 0:
                    if r == 15:
                        k += 16
                        continue
                    break
                k += r
                coefs[mcu, b, k] = extend(rd.bits(s), s)
                k += 1
    # remaining bits of the last byte must be 1-padding, then EOI
    while rd.n:
        assert rd.bit() == 1, "non-1 padding"
    tail = data[pos + rd.pos:]
    assert tail[:2] == b"\xFF\xD9", "missing EOI"
    return dict(height=frame["height"], width=frame["width"], subsampling=sub,
                qy=qt[0], qc=qt[1], coefs=coefs, geo=geo, restart_interval=restart_interval)
