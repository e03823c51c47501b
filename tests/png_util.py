"""A small PNG writer for the ingest tests: chosen scanline filter per row, chosen zlib strategy / level, IDAT split.
(PIL writes adaptive filters and one strategy only; the decoder under test has to take every filter type and every
DEFLATE block type.)  Test infrastructure."""
import struct
import zlib

import numpy as np


def _paeth(a, b, c):
    p = a.astype(np.int32) + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def filter_rows(arr, filters):
    """arr: (H, W, C) uint8; filters: per-row filter type 0..4 -> the filtered scanlines as bytes."""
    h, w, c = arr.shape
    rows = arr.reshape(h, w * c).astype(np.int32)
    out = bytearray()
    zero = np.zeros(w * c, dtype=np.int32)
    for r in range(h):
        cur = rows[r]
        up = rows[r - 1] if r else zero
        left = np.concatenate([np.zeros(c, np.int32), cur[:-c]])
        upleft = np.concatenate([np.zeros(c, np.int32), up[:-c]])
        ft = int(filters[r])
        if ft == 0:
            pred = zero
        elif ft == 1:
            pred = left
        elif ft == 2:
            pred = up
        elif ft == 3:
            pred = (left + up) >> 1
        elif ft == 4:
            pred = _paeth(left, up, upleft)
        else:
            raise ValueError(ft)
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
    return bytes(out)


def chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def write_png(arr, filters=None, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, idat_split=0, depth=8, interlace=0, raw_override=None):
    """arr: (H, W), (H, W, 3) or (H, W, 4) uint8 -> PNG file bytes."""
    if arr.ndim == 2:
        arr = arr[..., None]
    h, w, c = arr.shape
    ctype = {1: 0, 3: 2, 4: 6}[c]
    if filters is None:
        filters = [0] * h
    raw = filter_rows(arr, filters) if raw_override is None else raw_override
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    z = co.compress(raw) + co.flush()
    parts = [z] if not idat_split else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    out += chunk(b"tEXt", b"Comment\x00b2j test")
    for p in parts:
        out += chunk(b"IDAT", p)
    return out + chunk(b"IEND", b"")
