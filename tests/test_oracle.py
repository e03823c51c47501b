"""The NumPy oracle against the golden fixtures generated from the REAL reference
model (tests/golden/make_golden.py), and the [ext] stages against independent
decoders.  CPU only."""
import io
import math

import numpy as np
import pytest

from oracle import gateware as GW
from oracle import jpeg_ext as J
from oracle import model_np as M


def test_constants_match_python_math():
    # enc_frame.py:38-44: math.cos((n*pi)/16) with the model's truncated pi
    k = M.dct_constants()
    hexes = {"c1": "0x1.f6297cff75cb0p-1", "s1": "0x1.8f8b83c69a60ap-3", "c3": "0x1.a9b66290ea1a3p-1",
             "s3": "0x1.1c73b39ae68c8p-1", "c6": "0x1.87de2a6aea964p-2", "s6": "0x1.d906bcf328d46p-1",
             "sqrt2": "0x1.6a09e667f3bcdp+0"}
    for name, hx in hexes.items():
        assert k[name] == float.fromhex(hx), name
    src = open("jpeg_encoder_gsoc2017_ishan-bansal_b200/csrc/b2j_math.cuh").read() + open("oracle/c/b2j_oracle.c").read()
    for hx in hexes.values():
        assert hx in src


def test_lena_every_block_matches_model(golden):
    rgb = golden.lena["rgb"]
    zz = golden.lena["zigzag"]
    y, cb, cr = M.encode_planes_444(rgb)
    assert np.array_equal(y, zz[0]) and np.array_equal(cb, zz[1]) and np.array_equal(cr, zz[2])
    assert list(zz[0][0][:10]) == [51, 2, 2, -1, 0, 1, 0, 0, 0, 1]


def test_stage_vectors(golden):
    s = golden.stages
    y, cb, cr = M.rgb2ycbcr(s["rgb"][:, 0], s["rgb"][:, 1], s["rgb"][:, 2])
    assert np.array_equal(np.stack([y, cb, cr], 1), s["ycbcr"])
    assert np.array_equal(M.dct(s["dct_in"]), s["dct_out"])
    assert np.array_equal(M.quantize(s["dct_out"], s["y_quant_table"]), s["quant_y"])
    assert np.array_equal(M.quantize(s["dct_out"], s["cbcr_quant_table"]), s["quant_c"])
    assert np.array_equal(M.zigzag(s["quant_y"]), s["zigzag_out"])
    assert np.array_equal(s["zig_zag_order"], M.zig_zag_order)
    assert np.array_equal(s["y_quant_table"], M.y_quant_table) and np.array_equal(s["cbcr_quant_table"], M.cbcr_quant_table)


def test_exhaustive_colour_hash(golden):
    import hashlib
    allv = np.arange(1 << 24, dtype=np.int64)
    y, cb, cr = M.rgb2ycbcr(allv >> 16, (allv >> 8) & 255, allv & 255)
    h = hashlib.sha256()
    for a in (y, cb, cr):
        h.update(a.astype(np.uint8).tobytes())
    assert h.hexdigest() == golden.meta["rgb2ycbcr_all_2p24_sha256"]
    assert golden.meta["rgb2ycbcr_np_vs_model_mismatches"] == 0
    assert golden.meta["quantize_integer_restatement_mismatches_d1_255_v_pm2048"] == 0


def test_reference_fixture_classes(golden):
    r = golden.ref
    # DCTData (test/common.py:202-222): the model output is the pin, the paper vector is +-1 off
    assert np.array_equal(M.dct(r["dct_input"]), r["dct_output_model"])
    assert np.abs(M.dct(r["dct_input"]) - r["dct_output_paper"]).max() <= 37   # [0][7]: -18 vs 19 in the paper
    # Quantizer (test/common.py:309-351): enc_frame.quantize reproduces the expected vector exactly
    assert np.array_equal(M.quantize(r["quantizer_input"], r["quantizer_table"]), r["quantize_output_ref"])
    assert np.array_equal(M.quantize(r["quantizer_input"], r["quantizer_table"]), r["quantizer_output"])
    assert np.array_equal(M.quantize(r["quantizer_input"], r["quantizer_cr"]), r["quantizer_output2_ref"])
    # ZZData (test/common.py:280-289)
    assert np.array_equal(M.zigzag(r["zigzag_input"]), r["zigzag_output"])
    # RLE fixture (test/common.py:439): generic groupby run-length of the model
    assert M.rle_code(r["rle_matrices"][0]) == list(r["rle_output_red_pixels_1"])
    # the two colour models of RAWImage (test/common.py:133-154) on lena 64x64
    rgb = r["raw64_rgb"]
    assert np.array_equal(np.stack(M.rgb2ycbcr(rgb[0], rgb[1], rgb[2])), r["raw64_ycbcr_jfif"])
    assert np.array_equal(np.stack(GW.rgb2ycbcr_gateware_model(rgb[0], rgb[1], rgb[2])), r["raw64_ycbcr_xapp930_model"])
    # the integer gateware datapath stays within 2 LSB of its float model
    gw = np.stack(GW.rgb2ycbcr_gateware(rgb[0], rgb[1], rgb[2]))
    assert np.abs(gw - np.clip(r["raw64_ycbcr_xapp930_model"], 0, 255)).max() <= 2


def test_decode_direction_models(golden):
    r = golden.ref
    ycc = r["raw64_ycbcr_jfif"]
    assert np.array_equal(np.stack(GW.ycbcr2rgb_jfif(*ycc)), r["raw64_rgb_from_jfif_via_wikipedia"])
    assert np.array_equal(np.stack(GW.ycbcr2rgb_gateware_model(*ycc)), r["raw64_rgb_from_jfif_via_xapp931_model"])
    gw = np.stack(GW.ycbcr2rgb_gateware(*ycc))
    assert np.abs(gw - np.clip(r["raw64_rgb_from_jfif_via_xapp931_model"], 0, 255)).max() <= 1
    assert GW.ycbcr2rgb_coefs(8, 8)["acoef"] == 116 and GW.ycbcr2rgb_coefs(8, 8)["ccoef"] == -27
    y, cb, cr = GW.ycbcr422to444(*GW.ycbcr444to422(*ycc))
    assert np.array_equal(y, ycc[0]) and (cb[0::2] == cb[1::2]).all()


def test_reference_py_vectors(golden):
    r = golden.ref
    # test/model/reference.py: quant->zigzag vectors are exact; the colour vectors are only within
    # +-4 and the DCT vectors within +-1 of the model (they were produced by a different encoder)
    for i in range(2):
        assert np.array_equal(M.zigzag(r["ref_y422_qwant_16x8_blocks"][i]), r["ref_y422_zigzag_16x8_blocks"][i])
    assert np.array_equal(M.zigzag(r["ref_cb422_qwant_16x8_block"]), r["ref_cb_zigzag_16x8_block"])
    assert np.array_equal(M.quantize_trunc(r["ref_cb422_dct_16x8_block"], M.cbcr_quant_table), r["ref_cb422_qwant_16x8_block"])
    assert np.array_equal(M.quantize_trunc(r["ref_cr422_dct_16x8_block"], M.cbcr_quant_table), r["ref_cr422_qwant_16x8_block"])
    y, cb, cr = M.rgb2ycbcr(r["ref_r444_16x8_blocks"], r["ref_g444_16x8_blocks"], r["ref_b444_16x8_blocks"])
    assert np.abs(y - r["ref_y444_16x8_blocks"]).max() <= 3
    assert np.abs(cb - r["ref_cb444_16x8_blocks"]).max() <= 2 and np.abs(cr - r["ref_cr444_16x8_blocks"]).max() <= 4
    assert np.abs(M.dct(r["ref_y422_16x8_blocks"]) - r["ref_y422_dct_16x8_blocks"]).max() <= 1
    assert np.abs(M.dct(r["ref_cb422_16x8_block"]) - r["ref_cb422_dct_16x8_block"]).max() <= 1


def test_lena_dc_diff_and_rle(golden):
    zz = golden.lena["zigzag"].astype(np.int64)
    k = 0
    for i in range(64):
        for c in range(3):
            blk = M.dc_diff(zz[c][:i + 1])[i]
            assert M.rle_code(blk) == golden.rle_first[k]
            k += 1


def test_gateware_rle_rules(golden):
    # red_pixels_1 through the RLEDatapath rules
    slots, dc = GW.run_length(golden.ref["rle_matrices"][0])
    got = [(run, amp) for valid, run, amp in slots if valid]
    assert got == [(0, 1), (0, 12), (15, 0), (1, 10), (0, 2), (0, 3), (0, 4), (15, 0), (15, 0), (5, 1), (0, 0)]
    assert dc == 1
    words, _ = GW.rle_main(golden.ref["rle_matrices"][0])
    assert len(words) == 64 and (words[1] >> 12) & 15 == 4          # size of 12
    # a ZRL that ends exactly at index 63 is not followed by an EOB (rlecore.py:101-114)
    blk = [5] + [0] * 47 + [0] * 16
    blk[47] = 3
    slots, _ = GW.run_length(blk)
    assert [(r, a) for v, r, a in slots if v][-1] == (15, 0)
    assert [GW.entropy_size(v) for v in (0, 1, -1, 2, 3, 4, -255, 1023, -1024)] == [0, 1, 1, 2, 2, 3, 8, 10, 11]


def test_gateware_quantizer_differs_from_model(golden):
    r = golden.ref
    gw = GW.quantize_gateware(r["quantizer_input"], r["quantizer_table"])
    assert (gw != r["quantizer_output"]).sum() == 2                 # misses 2 of 64


def test_quality_tables():
    qy, qc = J.quant_tables(85)
    assert np.array_equal(qy, M.y_quant_table) and np.array_equal(qc, M.cbcr_quant_table)
    assert np.array_equal(J.quant_tables(50)[0], J.BASE_LUMA) and np.array_equal(J.quant_tables(50)[0], GW.quant_values)
    matches = [q for q in range(1, 101) if np.array_equal(J.quant_tables(q)[0], M.y_quant_table)]
    assert matches == [85]


@pytest.mark.parametrize("sub", ["444", "422", "420"])
@pytest.mark.parametrize("zrl", ["eager", "standard"])
def test_ext_roundtrip_and_third_party_decoders(golden, sub, zrl):
    from PIL import Image
    rgb = golden.lena["rgb"][128:128 + 96, 200:200 + 120]              # 96 x 120: not a multiple of 16 wide
    jpg = J.encode_jpeg(rgb, sub, 85, zrl)
    d = J.decode_jpeg_coefficients(jpg)
    coefs, geo = J.encode_coefficients(rgb, sub, quality=85)
    assert np.array_equal(d["coefs"], coefs) and d["subsampling"] == sub
    assert np.array_equal(d["qy"], M.y_quant_table)
    pil = np.array(Image.open(io.BytesIO(jpg)).convert("RGB")).astype(np.float64)
    assert pil.shape == rgb.shape
    psnr = 10 * math.log10(255 ** 2 / np.mean((pil - rgb) ** 2))
    assert psnr > 30, psnr
    try:
        import cv2
    except ImportError:
        return
    dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
    assert dec is not None and dec.shape == rgb.shape


def test_ext_444_coefficients_are_the_model(golden):
    # with 4:4:4 and the model tables the [ext] encoder's coefficients are exactly enc_frame's
    rgb = golden.lena["rgb"]
    coefs, geo = J.encode_coefficients(rgb, "444")
    zz = golden.lena["zigzag"]
    assert np.array_equal(coefs[:, 0], zz[0]) and np.array_equal(coefs[:, 1], zz[1]) and np.array_equal(coefs[:, 2], zz[2])


def test_stuffing_and_padding():
    data = np.array([0x12, 0xFF, 0xFF, 0x00, 0xFF], np.uint8)
    assert J.stuff_bytes(data).tolist() == [0x12, 0xFF, 0, 0xFF, 0, 0x00, 0xFF, 0]
    packed, total = J.pack_bits(np.array([0b101, 0b1], np.uint64), np.array([3, 2]))
    assert total == 5 and packed.tolist() == [0b10101111]
