"""`litejpeg_b200.model`: the function-level mirror of test/model/enc_frame.py, run on the GPU,
against the fixtures generated from the real reference model (tests/golden/make_golden.py).
The test bodies follow the model's own main loop (enc_frame.py:200-238)."""
import numpy as np
import pytest

from oracle import model_np as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(lj):
    return lj.model


def lena_blocks(golden, n=None):
    rgb = golden.lena["rgb"]
    a = rgb.reshape(64, 8, 64, 8, 3).transpose(4, 0, 2, 1, 3).reshape(3, 4096, 64)
    return [a[c][:n].tolist() for c in range(3)]


def test_tables_and_order_are_the_models(model, golden):
    assert model.y_quant_table == golden.stages["y_quant_table"].tolist()
    assert model.cbcr_quant_table == golden.stages["cbcr_quant_table"].tolist()
    assert model.zig_zag_order == golden.stages["zig_zag_order"].tolist()
    assert model.huffman([1, 2, 3]) == [1, 2, 3]                    # identity stub, enc_frame.py:179-180


def test_rle_code_golden_vector(model, golden):
    """red_pixels_1 -> [1,1,12,1,0,17,10,1,2,1,3,1,4,1,0,37,1,1,0,3] (measured on the model)."""
    r = golden.ref
    assert model.rle_code(r["rle_matrices"][0].tolist()) == r["rle_output_red_pixels_1"].tolist()
    batch = model.rle_code(r["rle_matrices"].tolist())
    assert batch == [M.rle_code(b) for b in r["rle_matrices"]]


def test_rle_code_edge_blocks(model):
    blocks = [[7] * 64, list(range(64)), [0] * 63 + [5], [-3] + [0] * 63, [1, 1, -1, -1] * 16]
    assert model.rle_code(blocks) == [M.rle_code(b) for b in blocks]
    assert model.rle_code(blocks[0]) == [7, 64]
    assert len(model.rle_code(blocks[1])) == 128                    # 64 pairs: the widest output


def test_single_block_functions_read_like_the_model(model, golden):
    """One block through the model's loop body, one call per stage, lists of 64 in and out."""
    s = golden.stages
    blk = s["dct_in"][5].tolist()
    d = model.dct(blk)
    assert isinstance(d, list) and d == s["dct_out"][5].tolist()
    q = model.quantize(d, model.y_quant_table)
    assert q == s["quant_y"][5].tolist()
    assert model.quantize(d, model.cbcr_quant_table) == s["quant_c"][5].tolist()
    z = model.zigzag(q)
    assert z == M.zigzag(np.array(q)).tolist()
    rgb = s["rgb"][:64]
    y, cb, cr = model.rgb2ycbcr(rgb[:, 0].tolist(), rgb[:, 1].tolist(), rgb[:, 2].tolist())
    assert [y, cb, cr] == [s["ycbcr"][:64, c].tolist() for c in range(3)]


def test_main_loop_on_lena_first_blocks(model, golden):
    """enc_frame.py:200-238 written out with this module's functions: zig-zag blocks with DC
    differences and their rle_code equal the model's for the first 64 blocks of lena."""
    r_blocks, g_blocks, b_blocks = lena_blocks(golden, 64)
    dcy_prev = dccb_prev = dccr_prev = 0
    k = 0
    zz = golden.lena["zigzag"]
    for i in range(64):
        y_block, cb_block, cr_block = model.rgb2ycbcr(r_blocks[i], g_blocks[i], b_blocks[i])
        y_zig = model.zigzag(model.quantize(model.dct(y_block), model.y_quant_table))
        cb_zig = model.zigzag(model.quantize(model.dct(cb_block), model.cbcr_quant_table))
        cr_zig = model.zigzag(model.quantize(model.dct(cr_block), model.cbcr_quant_table))
        assert y_zig == zz[0][i].tolist() and cb_zig == zz[1][i].tolist() and cr_zig == zz[2][i].tolist()
        y_zig[0], dcy_prev = y_zig[0] - dcy_prev, y_zig[0]
        cb_zig[0], dccb_prev = cb_zig[0] - dccb_prev, cb_zig[0]
        cr_zig[0], dccr_prev = cr_zig[0] - dccr_prev, cr_zig[0]
        for blk in (y_zig, cb_zig, cr_zig):
            assert model.huffman(model.rle_code(blk)) == golden.rle_first[k]
            k += 1


def test_encode_blocks_whole_lena(model, golden):
    """The batched form of the same loop over all 3 x 4096 blocks, in a handful of launches."""
    r_blocks, g_blocks, b_blocks = lena_blocks(golden)
    zigs, rles = model.encode_blocks(r_blocks, g_blocks, b_blocks)
    zz = golden.lena["zigzag"].astype(np.int64)
    for c in range(3):
        want = M.dc_diff(zz[c])
        assert np.array_equal(np.array(zigs[c]), want)
        for i in (0, 1, 63, 2047, 4095):
            assert rles[c][i] == M.rle_code(want[i])
    k = 0
    for i in range(64):
        for c in range(3):
            assert rles[c][i] == golden.rle_first[k]
            k += 1


def test_read_rgb_image_blocks(model, golden, tmp_path):
    from PIL import Image
    rgb = golden.lena["rgb"][:40, :72]
    path = str(tmp_path / "crop.png")
    Image.fromarray(rgb).save(path)
    r, g, b = model.readRgbImageBlocks(path)
    assert len(r) == 5 * 9 and len(r[0]) == 64
    assert r[10] == rgb[8:16, 8:16, 0].reshape(-1).tolist()         # block 10 = block row 1, column 1
    assert b[44] == rgb[32:40, 64:72, 2].reshape(-1).tolist()


def test_argument_errors(model):
    with pytest.raises(ValueError):
        model.dct([1, 2, 3])
    with pytest.raises(ValueError):
        model.quantize([0] * 64, [0] * 64)
    with pytest.raises(ValueError):
        model.rgb2ycbcr([0] * 64, [0] * 64, [[0] * 64, [0] * 64])
