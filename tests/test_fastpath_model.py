"""K1's single-precision path (csrc/b2j_fused_pair.cuh): where its error bound comes from, and that
"both ends of the interval agree" implies "equal to the model".  CPU only (NumPy restatement in
tests/fastpath_model.py; the GPU kernel itself is compared with the oracle in test_gpu_encode.py).

The model side is test/model/enc_frame.py:54-129 through oracle/model_np.py.
"""
import ctypes
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fastpath_model as fm                      # noqa: E402
from oracle import jpeg_ext as J                 # noqa: E402
from oracle import model_np                      # noqa: E402

U32 = 2.0 ** -24                                 # unit roundoff of float32


# ------------------------------------------------------------------ worst-case forward error bound
class _Node:
    """A value of the flow graph as a linear functional of the 64 samples (w) with a bound e on the
    distance between the float32 evaluation and the exact value of that functional."""

    def __init__(self, w, e):
        self.w, self.e = w, e

    def bound(self):                             # samples - 128 lie in [-128, 127]
        return 128 * np.abs(self.w).sum()


def _rounded(w, e_in, exact=False):
    n = _Node(w, e_in)
    if not exact:
        n.e = e_in + U32 * (n.bound() + e_in)
    return n


def _add(a, b, sign=1):
    w = a.w + sign * b.w
    exact = a.e == 0 and b.e == 0 and bool(np.all(w == np.round(w)))      # integers below 2^24
    return _rounded(w, a.e + b.e, exact)


def _mul(a, c):
    return _rounded(a.w * c, abs(c) * a.e)


def _fma(a, c, t, sign=1):
    return _rounded(a.w * c + sign * t.w, abs(c) * a.e + t.e)


def _dct1d(v, k):
    a3, a4 = _add(v[3], v[4]), _add(v[3], v[4], -1)
    a2, a5 = _add(v[2], v[5]), _add(v[2], v[5], -1)
    a1, a6 = _add(v[1], v[6]), _add(v[1], v[6], -1)
    a0, a7 = _add(v[0], v[7]), _add(v[0], v[7], -1)
    b0, b3 = _add(a0, a3), _add(a0, a3, -1)
    b1, b2 = _add(a1, a2), _add(a1, a2, -1)
    b4 = _fma(a4, k["c3"], _mul(a7, k["s3"]))
    b7 = _fma(a7, k["c3"], _mul(a4, k["s3"]), -1)
    b5 = _fma(a5, k["c1"], _mul(a6, k["s1"]))
    b6 = _fma(a6, k["c1"], _mul(a5, k["s1"]), -1)
    e7, e5 = _add(b7, b5), _add(b7, b5, -1)
    e4, e6 = _add(b4, b6), _add(b4, b6, -1)
    r2 = _fma(b2, k["c6"], _mul(b3, k["s6"]))
    r6 = _fma(b3, k["c6"], _mul(b2, k["s6"]), -1)
    return [_add(b0, b1), _add(e7, e4), r2, e5, _add(b0, b1, -1), e6, r6, _add(e7, e4, -1)]


def _graph(k):
    inp = [[_Node(np.eye(64)[x * 8 + y], 0.0) for y in range(8)] for x in range(8)]
    rows = [_dct1d(inp[x], k) for x in range(8)]
    return [_dct1d([rows[x][u] for x in range(8)], k) for u in range(8)]       # [u][x]


def test_forward_error_bound_is_below_the_constant_the_kernel_uses():
    """Rounding errors of every float32 operation at the largest magnitude its operand can take, the distance
    between the float32 and the model's float64 constants, and the three roundings of the quantiser
    (|out| - Ha, - (Hb - Ha), the float32 multiplier) stay below kFastPathErr = 1/1024 in units of v / 8."""
    k32 = {n: float(v) for n, v in fm.consts32().items()}
    k64 = {"c6": fm.SQRT2 * math.cos(6 * fm.PI / 16), "s6": fm.SQRT2 * math.sin(6 * fm.PI / 16)}
    for n in (1, 3):
        k64["c%d" % n], k64["s%d" % n] = math.cos(n * fm.PI / 16), math.sin(n * fm.PI / 16)
    g32, g64 = _graph(k32), _graph(k64)
    worst = 0.0
    for u in range(8):
        for x in range(8):
            sc = fm.SCALE[x] * fm.SCALE[u] / 8
            n = g32[u][x]
            mismatch = 128 * np.abs(n.w - g64[u][x].w).sum()
            b_a = n.bound() * sc                                   # largest |v / 8| of this coefficient
            total = (n.e + mismatch) * sc + 3 * U32 * b_a
            if fm.EXACT[x, u]:
                assert n.e == 0 and mismatch == 0                  # sums of integers: exact
            worst = max(worst, total)
    assert 5e-4 < worst < 8e-4                                     # 7.5e-4: the derivation in DESIGN.md quotes it
    assert worst + 1e-9 < fm.E_A                                   # (1e-9: the model's own float64 roundings)


def _blocks(kind, n, rng):
    if kind == "noise":
        return rng.integers(0, 256, (n, 64))
    if kind == "extreme":
        return rng.integers(0, 2, (n, 64)) * 255
    if kind == "basis":                                            # +-full-scale cosine sign patterns: the largest coefficients
        xx, yy = np.meshgrid(np.arange(8), np.arange(8))
        out = np.empty((n, 64), dtype=np.int64)
        for i in range(n):
            p, q = rng.integers(0, 8, 2)
            s = np.cos((2 * xx + 1) * p * np.pi / 16) * np.cos((2 * yy + 1) * q * np.pi / 16)
            out[i] = np.where(s.reshape(64) >= 0, 255, 0) if rng.integers(0, 2) else np.where(s.reshape(64) >= 0, 0, 255)
        return out
    base = rng.integers(0, 256, (n, 1))
    gx, gy = rng.normal(0, 6, (n, 1)), rng.normal(0, 6, (n, 1))
    xx, yy = np.meshgrid(np.arange(8), np.arange(8))
    b = base + gx * xx.reshape(1, 64) + gy * yy.reshape(1, 64) + rng.integers(-12, 13, (n, 64))
    return np.clip(np.rint(b), 0, 255).astype(np.int64)


@pytest.mark.parametrize("kind", ["noise", "extreme", "basis", "smooth"])
def test_measured_distance_to_the_float64_graph(kind):
    rng = np.random.default_rng(20171019)
    b = _blocks(kind, 60000, rng)
    a32 = fm.transform_f32(b).astype(np.float64) * (fm.SCALE[:, None] * fm.SCALE[None, :]) / 8
    err = np.abs(a32 - fm.transform_f64(b))
    assert err.max() < fm.E_A / 4                                  # measured: 1.0e-4 at most
    assert err[:, fm.EXACT].max() == 0.0


@pytest.mark.parametrize("quality", [10, 50, 75, 90, 92])
@pytest.mark.parametrize("kind", ["noise", "smooth", "basis"])
def test_agreeing_interval_ends_give_the_models_value(quality, kind):
    rng = np.random.default_rng(quality * 7 + len(kind))
    b = _blocks(kind, 40000, rng)
    out = fm.transform_f32(b)
    for table in J.quant_tables(quality):
        if not fm.table_supported(table):
            continue
        q, uncertain = fm.quantise_fast(out, table)
        want = fm.quantise_model(b, table)
        assert np.array_equal(q[~uncertain], want[~uncertain])
        assert uncertain.any(axis=(1, 2)).mean() < 0.05            # blocks sent to the exact path


def test_model_tables_and_steps_exactly_on_a_threshold():
    """Flat blocks put the DC (an exact multiple of 1/8) on and next to every step of the quantiser: the four exactly
    computed coefficients never need the exact path and are always right."""
    levels = np.arange(256)
    b = np.repeat(levels[:, None], 64, axis=1)
    out = fm.transform_f32(b)
    for table in (model_np.y_quant_table, model_np.cbcr_quant_table) + tuple(J.quant_tables(75)):
        q, uncertain = fm.quantise_fast(out, table)
        assert not uncertain.any()
        assert np.array_equal(q, fm.quantise_model(b, table))
    # checkerboards of two levels: (0, 0), (0, 4), (4, 0), (4, 4) only
    rng = np.random.default_rng(5)
    lo, hi = rng.integers(0, 256, (2, 4000, 1))
    xx, yy = np.meshgrid(np.arange(8), np.arange(8))
    pat = (((xx // 1) + (yy // 1)) % 2).reshape(1, 64)
    b = np.where(pat == 1, hi, lo)
    out = fm.transform_f32(b)
    q, uncertain = fm.quantise_fast(out, J.quant_tables(75)[0])
    assert not uncertain[:, fm.EXACT].any()
    assert np.array_equal(q[:, fm.EXACT], fm.quantise_model(b, J.quant_tables(75)[0])[:, fm.EXACT])


def test_library_constants_match_the_restatement():
    """b2j_fastpath_constants (host only, no device needed) against fastpath_model.lane_constants."""
    from litejpeg_b200 import _lib
    lib = _lib.lib()
    for quality in (25, 75, 90, 100):
        qy, qc = J.quant_tables(quality)
        out = np.zeros((3, 2, 64), dtype=np.float32)
        ty = (ctypes.c_uint8 * 64)(*[int(x) for x in qy])
        tc = (ctypes.c_uint8 * 64)(*[int(x) for x in qc])
        rc = lib.b2j_fastpath_constants(ty, tc, out.ctypes.data_as(ctypes.c_void_p))
        assert rc == (1 if fm.table_supported(qy) and fm.table_supported(qc) else 0)
        for c, table in enumerate((qy, qc)):
            ha, hb, r = fm.lane_constants(table)
            assert np.array_equal(out[0, c].reshape(8, 8), r)
            assert np.array_equal(out[1, c].reshape(8, 8), ha)
            assert np.array_equal(out[2, c].reshape(8, 8), hb)
