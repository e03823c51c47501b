"""The plain-C oracle (oracle/c/b2j_oracle.c) against the NumPy oracle and the golden
fixtures.  CPU only."""
import numpy as np
import pytest

from oracle import jpeg_ext as J
from oracle import model_np as M


def test_c_stage_functions(oracle_c, golden):
    s = golden.stages
    lib = oracle_c.lib
    p = oracle_c._p
    rgb = np.ascontiguousarray(s["rgb"].astype(np.uint8))
    n = rgb.shape[0]
    y, cb, cr = (np.zeros(n, np.uint8) for _ in range(3))
    lib.b2jo_rgb2ycbcr(p(rgb), n, p(y), p(cb), p(cr))
    assert np.array_equal(np.stack([y, cb, cr], 1), s["ycbcr"])
    din = np.ascontiguousarray(s["dct_in"].astype(np.int32))
    dout = np.zeros_like(din)
    lib.b2jo_dct(p(din), din.shape[0], p(dout))
    assert np.array_equal(dout, s["dct_out"])
    q = np.zeros_like(dout)
    tab = np.ascontiguousarray(s["y_quant_table"].astype(np.int32))
    lib.b2jo_quantize(p(dout), dout.shape[0], p(tab), p(q))
    assert np.array_equal(q, s["quant_y"])
    z = np.zeros_like(q)
    lib.b2jo_zigzag(p(q), q.shape[0], p(z))
    assert np.array_equal(z, s["zigzag_out"])


def test_c_lena_coefficients_are_the_model(oracle_c, golden):
    rgb = golden.lena["rgb"]
    coefs = oracle_c.coefficients(rgb, 444, M.y_quant_table, M.cbcr_quant_table)
    zz = golden.lena["zigzag"]
    for c in range(3):
        assert np.array_equal(coefs[:, c], zz[c])


@pytest.mark.parametrize("sub", ["444", "422", "420"])
@pytest.mark.parametrize("zrl", [0, 1])
@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (37, 53), (96, 120)])
def test_c_bitstream_equals_numpy(oracle_c, golden, sub, zrl, shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    if h <= 37:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        rgb = np.ascontiguousarray(golden.lena["rgb"][100:100 + h, 60:60 + w])
    ref = J.encode_jpeg(rgb, sub, 75, "eager" if zrl == 0 else "standard")
    got = oracle_c.encode_jpeg(rgb, sub, 75, zrl)
    assert got == ref
    rc, coefs = oracle_c.decode_coefficients(got, h, w, sub)
    assert rc == 0
    exp, geo = J.encode_coefficients(rgb, sub, quality=75)
    assert np.array_equal(coefs, exp)
    bits = oracle_c.block_bits(exp, h, w, sub, zrl)
    assert np.array_equal(bits.reshape(exp.shape[:2]), J.block_bit_lengths(exp, geo, "eager" if zrl == 0 else "standard"))


def test_c_quality_tables(oracle_c):
    for q in (1, 10, 49, 50, 75, 85, 90, 100):
        qy, qc = oracle_c.quant_tables(q)
        ey, ec = J.quant_tables(q)
        assert np.array_equal(qy, ey) and np.array_equal(qc, ec)


def test_c_special_content(oracle_c):
    # constant grey (EOB-only blocks) and saturated noise, all modes
    for kind in ("grey", "noise", "extremes"):
        if kind == "grey":
            rgb = np.full((32, 48, 3), 128, np.uint8)
        elif kind == "noise":
            rgb = np.random.default_rng(5).integers(0, 256, (32, 48, 3), dtype=np.uint8)
        else:
            rgb = (np.indices((32, 48)).sum(0) % 2 * 255).astype(np.uint8)[..., None].repeat(3, 2)
        for sub in ("444", "422", "420"):
            for q in (50, 100):
                assert oracle_c.encode_jpeg(rgb, sub, q, 0) == J.encode_jpeg(rgb, sub, q, "eager")


def test_fixed_dct_mode_c_equals_numpy_and_stays_near_the_model(oracle_c, golden):
    """[ext] dct_mode="fixed" (SURVEY 8f-3): the gateware datapath's constants and shift points
    (new_dct6.py:171-396), wiring slips corrected.  The two restatements agree bit for bit; the
    mode is NOT parity with the float64 model: measured deviation <= 4 LSB per coefficient."""
    from oracle import gateware as G
    from oracle import model_np as M
    rng = np.random.default_rng(7)
    blocks = np.concatenate([rng.integers(0, 256, (3000, 64)), golden.stages["dct_in"],
                             np.full((1, 64), 0), np.full((1, 64), 255), np.full((1, 64), 128)])
    f = G.dct_fixed(blocks)
    assert np.array_equal(oracle_c.dct_fixed(blocks), f)
    dev = np.abs(f - M.dct(blocks))
    assert dev.max() <= 4 and dev.mean() < 1.0
    assert f[-3, 0] == -1024 and f[-2, 0] == 1016 and not f[-1].any()      # flat blocks: exact DC, no AC
    assert not f[-3, 1:].any() and not f[-2, 1:].any()


def test_restart_intervals_c_equals_numpy_and_libjpeg_decodes_them(oracle_c):
    """[ext] restart intervals (T.81 B.2.4.4, E.1.4).  No reference counterpart, so
    the pin is threefold: the two restatements agree byte for byte; both baseline decoders of
    oracle/ recover the same coefficients as from the file without restarts; and libjpeg (PIL)
    decodes every file to exactly the pixels of the file without restarts."""
    import io
    from PIL import Image
    from oracle import jpeg_ext as J
    rng = np.random.default_rng(5)
    img = np.clip(np.add.outer(np.arange(40) * 5, np.arange(72) * 2)[..., None] + rng.integers(-20, 20, (40, 72, 3)),
                  0, 255).astype(np.uint8)
    for sub in ("444", "422", "420"):
        plain = J.encode_jpeg(img, sub, 75, "standard")
        ref_px = np.asarray(Image.open(io.BytesIO(plain)).convert("RGB"))
        ref_coefs = J.decode_jpeg_coefficients(plain)["coefs"]
        n_mcu = J.geometry(40, 72, sub)["n_mcu"]
        for zi, zrl in enumerate(("eager", "standard")):
            for ri in (1, 2, 3, 9, n_mcu - 1, n_mcu, 1000):
                a = J.encode_jpeg(img, sub, 75, zrl, ri)
                assert a == oracle_c.encode_jpeg(img, sub, 75, zi | (ri << 16)), (sub, zrl, ri)
                n_markers = sum(a.count(bytes([0xFF, 0xD0 + m])) for m in range(8))
                assert n_markers >= (n_mcu - 1) // ri                      # (entropy bytes are stuffed, so >= is exact enough)
                d = J.decode_jpeg_coefficients(a)
                assert d["restart_interval"] == ri and np.array_equal(d["coefs"], ref_coefs)
                rc, co = oracle_c.decode_coefficients(a, 40, 72, sub)
                assert rc == 0 and np.array_equal(co, ref_coefs)
                if zrl == "standard":
                    assert np.array_equal(np.asarray(Image.open(io.BytesIO(a)).convert("RGB")), ref_px)


def test_random_parameter_space_c_equals_numpy(oracle_c):
    """Seeded sweep over shape x subsampling x quality x ZRL policy x restart interval: the two
    restatements of the [ext] parts agree byte for byte and the decoder recovers the coefficients."""
    from oracle import jpeg_ext as J
    rng = np.random.default_rng(20171019)
    for case in range(24):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        sub = ("444", "422", "420")[case % 3]
        q = int(rng.integers(1, 101))
        zi = int(rng.integers(0, 2))
        ri = int(rng.choice([0, 0, 1, 2, 5, 17]))
        kind = case % 4
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        elif kind == 1:
            img = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
        else:
            base = np.add.outer(np.arange(h) * int(rng.integers(0, 6)), np.arange(w) * int(rng.integers(0, 6)))[..., None]
            img = np.clip(base + rng.integers(-9, 10, (h, w, 3)) + rng.integers(0, 200, 3), 0, 255).astype(np.uint8)
        a = J.encode_jpeg(img, sub, q, ("eager", "standard")[zi], ri)
        assert a == oracle_c.encode_jpeg(img, sub, q, zi | (ri << 16)), (case, h, w, sub, q, zi, ri)
        rc, co = oracle_c.decode_coefficients(a, h, w, sub)
        qy, qc = oracle_c.quant_tables(q)
        assert rc == 0 and np.array_equal(co, oracle_c.coefficients(img, sub, qy, qc)), case
