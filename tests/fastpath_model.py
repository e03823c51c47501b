"""NumPy model of K1's single-precision fast path (csrc/b2j_fused_pair.cuh) - test infrastructure.

The kernel evaluates the DCT flow graph of test/model/enc_frame.py:54-109 in float32 (fused
multiply-adds, sqrt2 factors folded into constants), quantises with one multiply-add per bound of an
uncertainty interval and falls back to the exact float64 graph whenever the two bounds round to
different integers.  This module restates that arithmetic so that CPU tests can
  * measure the distance between the float32 and the float64 transform (the bound E_A the guard uses),
  * check that "both bounds agree" implies "equal to the model" on random and adversarial blocks,
  * report how often the guard fires.
float32 fma(a, b, c) is emulated as float32(float64(a) * float64(b) + float64(c)): the product is
exact in float64; the double rounding of the sum can differ from the hardware in the last float32 bit
only, which is far inside the margin the bound keeps.
"""
import math

import numpy as np

f32 = np.float32
PI = 3.14159265358979323846
SQRT2 = 1.41421356237309504880

# bound on |float32 transform / 8 - float64 transform / 8| the guard assumes (see tests/test_fastpath_model.py)
E_A = 1.0 / 1024
MAGIC = 12582912.0          # 1.5 * 2^23: float32 ulp is 1 in [2^23, 2^24)


def fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + np.asarray(c, dtype=np.float64)).astype(f32)


def consts32():
    c = {}
    for n in (1, 3):
        c["c%d" % n] = f32(math.cos(n * PI / 16))
        c["s%d" % n] = f32(math.sin(n * PI / 16))
    # the sqrt2 of the third rotation is folded into its constants
    c["c6"] = f32(SQRT2 * math.cos(6 * PI / 16))
    c["s6"] = f32(SQRT2 * math.sin(6 * PI / 16))
    return c


def dct_1d_f32(v):
    """8 float32 arrays -> 8 float32 arrays; outputs 3 and 5 lack their factor sqrt2."""
    k = consts32()
    a3, a4 = v[3] + v[4], v[3] - v[4]
    a2, a5 = v[2] + v[5], v[2] - v[5]
    a1, a6 = v[1] + v[6], v[1] - v[6]
    a0, a7 = v[0] + v[7], v[0] - v[7]
    b0, b3 = a0 + a3, a0 - a3
    b1, b2 = a1 + a2, a1 - a2
    b4 = fma(a4, k["c3"], (a7 * k["s3"]).astype(f32))
    b7 = fma(a7, k["c3"], -(a4 * k["s3"]).astype(f32))
    b5 = fma(a5, k["c1"], (a6 * k["s1"]).astype(f32))
    b6 = fma(a6, k["c1"], -(a5 * k["s1"]).astype(f32))
    e7, e5 = b7 + b5, b7 - b5
    e4, e6 = b4 + b6, b4 - b6
    r2 = fma(b2, k["c6"], (b3 * k["s6"]).astype(f32))
    r6 = fma(b3, k["c6"], -(b2 * k["s6"]).astype(f32))
    return [b0 + b1, e7 + e4, r2, e5, b0 - b1, e6, r6, e7 - e4]


SCALE = np.array([1, 1, 1, SQRT2, 1, SQRT2, 1, 1])          # missing factor of output k
EXACT = np.zeros((8, 8), dtype=bool)                        # coefficients that are integers / 8, computed exactly
EXACT[np.ix_([0, 4], [0, 4])] = True


def transform_f32(blocks):
    """(n, 64) samples 0..255 -> (n, 8, 8) float32 'out' with v[x][u] ~= out[x][u] * SCALE[x] * SCALE[u]."""
    m = blocks.reshape(-1, 8, 8).astype(f32) - f32(128)
    rows = dct_1d_f32([m[:, :, y] for y in range(8)])
    m = np.stack(rows, axis=2)
    cols = dct_1d_f32([m[:, x, :] for x in range(8)])
    return np.stack(cols, axis=1)


def transform_f64(blocks):
    """The model's float64 v / 8 before int() (enc_frame.py:86-108)."""
    from oracle import model_np
    m = blocks.reshape(-1, 8, 8).astype(np.float64) - 128
    rows = model_np.dct_1d([m[:, :, y] for y in range(8)])
    m = np.stack(rows, axis=2)
    cols = model_np.dct_1d([m[:, x, :] for x in range(8)])
    return np.stack(cols, axis=1) / 8


def lane_constants(table, e_a=E_A):
    """Per coefficient (natural order x * 8 + u), in units of out: Ha = H - e, Hb = H + e (subtracted with the sign of
    out: the upper and the lower bound of the magnitude), and the multiplier R.  H = 0 for an even divisor, 1/2 for an
    odd one (int() truncates towards zero before the model rounds), less 1/16 for the four exactly computed
    coefficients (multiples of 1/8), which need no interval."""
    d = np.asarray(table, dtype=np.float64).reshape(8, 8)
    s = SCALE[:, None] * SCALE[None, :]
    h = np.where(d.astype(np.int64) % 2 == 1, 0.5, 0.0) - np.where(EXACT, 1.0 / 16, 0.0)     # in units of v / 8
    e = np.where(EXACT, 0.0, e_a)
    return ((h - e) * 8 / s).astype(f32), ((h + e) * 8 / s).astype(f32), (s / (8 * d)).astype(f32)


def table_supported(table):
    """d == 1 puts the +-1/2 shift of an odd divisor on a rounding tie of y."""
    return bool(np.all(np.asarray(table) >= 2))


def quantise_fast(out, table, e_a=E_A):
    """-> (q int64 (n, 8, 8), uncertain bool (n, 8, 8))."""
    Ha, Hb, R = lane_constants(table, e_a)
    neg = np.signbit(out)                                                   # bits: H ^ (out & 0x80000000)
    oa = (out - np.where(neg, -Ha[None], Ha[None])).astype(f32)
    ob = (out - np.where(neg, -Hb[None], Hb[None])).astype(f32)
    z1 = fma(oa, R[None], f32(MAGIC))
    z2 = fma(ob, R[None], f32(MAGIC))
    q = z1.astype(np.float64) - MAGIC
    return q.astype(np.int64), z1 != z2


def quantise_model(blocks, table):
    from oracle import model_np
    return model_np.quantize(model_np.dct(blocks), np.asarray(table).reshape(-1)).reshape(-1, 8, 8)
