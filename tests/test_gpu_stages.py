"""Per-stage GPU operators against the golden fixtures of the reference model and the
CPU restatements of the gateware - one test per reference bench (test/*_tb.py): a Packet
goes in, a Packet comes out, and here it is asserted instead of printed.  All calls go
through the C ABI (libb2j.so)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import gateware as GW
from oracle import model_np as M

pytestmark = pytest.mark.gpu


def pack24(a, b, c):
    return [(int(x) << 16) | (int(y) << 8) | int(z) for x, y, z in zip(a, b, c)]


def unpack24(words):
    w = np.array(words, dtype=np.int64)
    return (w >> 16) & 255, (w >> 8) & 255, w & 255


def test_rgb2ycbcr_tb(lj, golden):
    """test/rgb2ycbcr_tb.py: lena 64x64 as 24-bit words through RGB2YCbCr()."""
    rgb = golden.ref["raw64_rgb"]
    pkt = lj.Packet(pack24(rgb[0], rgb[1], rgb[2]))
    out = lj.RGB2YCbCr().process(pkt)                               # gateware datapath (XAPP930)
    y, cb, cr = unpack24(out)
    ey, ecb, ecr = GW.rgb2ycbcr_gateware(rgb[0], rgb[1], rgb[2])
    assert np.array_equal(y, ey) and np.array_equal(cb, ecb) and np.array_equal(cr, ecr)
    out = lj.RGB2YCbCr(model="jfif").process(pkt)                   # software model, fixture of the reference
    assert np.array_equal(np.stack(unpack24(out)), golden.ref["raw64_ycbcr_jfif"])


def test_rgb2ycbcr_all_16777216_inputs(lj, golden):
    """Exhaustive: every (r, g, b) through the JFIF kernel, hashed against the model's table."""
    allv = torch.arange(1 << 24, dtype=torch.int64, device="cuda")
    rgb = torch.stack([(allv >> 16) & 255, (allv >> 8) & 255, allv & 255], dim=1).to(torch.uint8)
    out = lj.RGB2YCbCr(model="jfif").process(rgb).cpu().numpy()
    h = hashlib.sha256()
    for c in range(3):
        h.update(np.ascontiguousarray(out[:, c]).tobytes())
    assert h.hexdigest() == golden.meta["rgb2ycbcr_all_2p24_sha256"]
    # gateware model on a large random sample
    idx = torch.randint(0, 1 << 24, (200000,), generator=torch.Generator().manual_seed(1))
    sub = rgb[idx.cuda()]
    got = lj.RGB2YCbCr().process(sub).cpu().numpy().astype(np.int64)
    s = sub.cpu().numpy().astype(np.int64)
    exp = np.stack(GW.rgb2ycbcr_gateware(s[:, 0], s[:, 1], s[:, 2]), 1)
    assert np.array_equal(got, exp)


def test_ycbcr_resampling_tb(lj, golden):
    """test/ycbcr_resampling_tb.py: (y, cb, cr) words through YCbCr444to422()."""
    ycc = golden.ref["raw64_ycbcr_jfif"]
    out = lj.YCbCr444to422().process(lj.Packet(pack24(ycc[0], ycc[1], ycc[2])))
    w = np.array(out)
    ey, ecc = GW.ycbcr444to422(ycc[0], ycc[1], ycc[2])
    assert np.array_equal(w >> 8, ey) and np.array_equal(w & 255, ecc)
    # [ext] plane subsampler
    from oracle import jpeg_ext as J
    plane = torch.randint(0, 256, (48, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    for mode in ("420", "422"):
        got = lj.YCbCr444to420(mode=mode).process(plane.cuda()).cpu().numpy()
        exp = J.subsample(plane.numpy(), plane.numpy(), mode)[0]
        assert np.array_equal(got, exp)


def test_ycbcr2rgb_tb(lj, golden):
    """test/ycbcr2rgb_tb.py: lena 64x64 (y, cb, cr) words through YCbCr2RGB() (decode direction)."""
    ycc = golden.ref["raw64_ycbcr_jfif"]
    pkt = lj.Packet(pack24(ycc[0], ycc[1], ycc[2]))
    out = lj.YCbCr2RGB().process(pkt)                                # gateware datapath (XAPP931)
    exp = GW.ycbcr2rgb_gateware(ycc[0], ycc[1], ycc[2])
    assert np.array_equal(np.stack(unpack24(out)), np.stack(exp))
    dev = torch.from_numpy(np.ascontiguousarray(ycc.T)).to(torch.uint8).cuda()
    got = lj.YCbCr2RGB(model="jfif").process(dev).cpu().numpy().T     # float model, fixture of the reference, unclamped
    assert np.array_equal(got, golden.ref["raw64_rgb_from_jfif_via_wikipedia"])
    # every (y, cb, cr) triple, both models
    allv = torch.arange(1 << 24, dtype=torch.int64, device="cuda")
    trip = torch.stack([(allv >> 16) & 255, (allv >> 8) & 255, allv & 255], dim=1).to(torch.uint8)
    a = trip.cpu().numpy().astype(np.int64)
    for model, fn in (("xapp931", GW.ycbcr2rgb_gateware), ("jfif", GW.ycbcr2rgb_jfif)):
        got = lj.YCbCr2RGB(model=model).process(trip).cpu().numpy().astype(np.int64)
        exp = np.stack(fn(a[:, 0], a[:, 1], a[:, 2]), 1)
        assert np.array_equal(got, exp), model


def test_ycbcr_resampling_roundtrip_tb(lj, golden):
    """test/ycbcr_resampling_tb.py chains YCbCr444to422 -> YCbCr422to444."""
    ycc = golden.ref["raw64_ycbcr_jfif"]
    mid = lj.YCbCr444to422().process(lj.Packet(pack24(ycc[0], ycc[1], ycc[2])))
    out = lj.YCbCr422to444().process(mid)
    ey, ecc = GW.ycbcr444to422(ycc[0], ycc[1], ycc[2])
    exp = GW.ycbcr422to444(ey, ecc)
    assert np.array_equal(np.stack(unpack24(out)), np.stack(exp))
    assert np.array_equal(np.stack(unpack24(out))[0], ycc[0])        # luma is untouched


def test_new_dct_tb2(lj, golden):
    """test/new_dct_tb2.py: DCTData.input_dct through DCT(); the model output is the pin."""
    out = lj.DCT().process(lj.Packet(golden.ref["dct_input"].tolist()))
    assert lj.to_signed(out, 12) == golden.ref["dct_output_model"].tolist()
    assert lj.to_signed(out, 12)[:8] == [186, -17, 14, -8, 23, -9, -13, -18]
    s = golden.stages
    got = lj.DCT().process(torch.from_numpy(s["dct_in"]).to(torch.int16).cuda()).cpu().numpy()
    assert np.array_equal(got, s["dct_out"])


def test_dct_random_blocks_bit_exact(lj):
    g = torch.Generator().manual_seed(7)
    blocks = torch.randint(0, 256, (20000, 64), generator=g, dtype=torch.int16)
    smooth = (torch.randint(0, 256, (20000, 1), generator=g) + torch.randint(-5, 6, (20000, 64), generator=g)).clamp(0, 255).to(torch.int16)
    x = torch.cat([blocks, smooth])
    got = lj.DCT().process(x.cuda()).cpu().numpy()
    assert np.array_equal(got, M.dct(x.numpy()))


def test_new_quantize_tb(lj, golden):
    """test/new_quantize_tb.py: Quantizer.quantizer_input through Quantization()."""
    r = golden.ref
    pkt = lj.to_unsigned(r["quantizer_input"].tolist(), 12)
    out = lj.Quantization(table=r["quantizer_table"].tolist()).process(pkt)
    assert lj.to_signed(out, 12) == r["quantizer_output"].tolist()                 # the Wikipedia vector
    out = lj.Quantization(table=r["quantizer_cr"].tolist()).process(pkt)
    assert lj.to_signed(out, 12) == r["quantizer_output2_ref"].tolist()
    out = lj.Quantization(mode="reciprocal").process(pkt)                           # gateware datapath
    assert lj.to_signed(out, 12) == GW.quantize_gateware(r["quantizer_input"]).tolist()
    s = golden.stages
    x = torch.from_numpy(s["dct_out"]).to(torch.int16).cuda()
    assert np.array_equal(lj.Quantization(table=s["y_quant_table"].tolist()).process(x).cpu().numpy(), s["quant_y"])
    assert np.array_equal(lj.Quantization(table=s["cbcr_quant_table"].tolist()).process(x).cpu().numpy(), s["quant_c"])
    # every divisor, every 12-bit value, all three rounding modes
    v = torch.arange(-2048, 2048, dtype=torch.int16).repeat_interleave(1)
    for d in (1, 2, 3, 5, 16, 99, 121, 255):
        xx = v.reshape(-1, 64).cuda()
        assert np.array_equal(lj.Quantization(table=[d] * 64).process(xx).cpu().numpy().reshape(-1), M.quantize(v.numpy(), d))
        assert np.array_equal(lj.Quantization(table=[d] * 64, mode="trunc").process(xx).cpu().numpy().reshape(-1),
                              M.quantize_trunc(v.numpy(), d))
        assert np.array_equal(lj.Quantization(table=[d] * 64, mode="reciprocal").process(xx).cpu().numpy().reshape(-1),
                              GW.quantize_gateware(v.numpy(), np.full(1, d)))


def test_new_zigzag_tb(lj, golden):
    """test/new_zigzag_tb.py / zigzag_tb.py: ZZData.zigzag_input through ZigZag()."""
    out = lj.ZigZag().process(lj.Packet(golden.ref["zigzag_input"].tolist()))
    assert list(out) == golden.ref["zigzag_output"].tolist()
    s = golden.stages
    got = lj.ZigZag().process(torch.from_numpy(s["quant_y"]).to(torch.int16).cuda()).cpu().numpy()
    assert np.array_equal(got, s["zigzag_out"])


def test_new_rlecore_tb(lj, golden):
    """test/new_rlecore_tb.py: RLE.red_pixels_1 through RunLength()."""
    mats = golden.ref["rle_matrices"]
    rl = lj.RunLength()
    out = rl.process(lj.to_unsigned(mats[0].tolist(), 12))
    assert lj.RunLength.decode(out) == [(1, 0), (12, 0), (0, 15), (10, 1), (2, 0), (3, 0), (4, 0), (0, 15), (0, 15), (1, 5), (0, 0)]
    # all six fixture matrices as one component stream: the DC predictor carries over
    rl.reset()
    prev = 0
    for m in mats:
        out = rl.process(lj.to_unsigned(m.tolist(), 12))
        slots, prev = GW.run_length(m, prev)
        assert list(out) == [GW.pack_rle_word(v, r, a) for v, r, a in slots]
    # batch form
    x = torch.from_numpy(mats).to(torch.int16).cuda()
    rl.reset()
    words = rl.process(x).cpu().numpy()
    prev = 0
    for i, m in enumerate(mats):
        slots, prev = GW.run_length(m, prev)
        assert words[i].tolist() == [GW.pack_rle_word(v, r, a) for v, r, a in slots]


def test_new_entropy_tb(lj, golden):
    """test/new_entropy_tb.py: sizes of RLE.red_pixels_1 through EntropyCoder()."""
    m = golden.ref["rle_matrices"][0]
    out = lj.EntropyCoder().process(lj.to_unsigned(m.tolist(), 12))
    assert list(out) == [GW.entropy_size(v) for v in m]
    v = torch.arange(-2047, 2048, dtype=torch.int16)
    assert lj.EntropyCoder().process(v.cuda()).cpu().tolist() == [GW.entropy_size(int(x)) for x in v]


def test_new_rlemain_tb(lj, golden):
    """test/new_rlemain_tb.py: RLE.red_pixels_1 through RLEMain()."""
    mats = golden.ref["rle_matrices"]
    top = lj.RLEMain()
    out = top.process(lj.to_unsigned(mats[0].tolist(), 12))
    words, _ = GW.rle_main(mats[0])
    assert list(out) == words
    assert lj.RLEMain.decode(out)[:4] == [(1, 0, 1), (12, 0, 4), (0, 15, 0), (10, 1, 4)]
    signed = np.array([[-5, 3, 0, 0, -1] + [0] * 59, [7, 0, 0, -300] + [0] * 60])
    top.reset()
    got = top.process(torch.from_numpy(signed).to(torch.int16).cuda()).cpu().numpy()
    prev = 0
    for i in range(2):
        words, prev = GW.rle_main(signed[i], prev)
        assert got[i].tolist() == words


def test_empty_and_ragged_inputs(lj):
    with pytest.raises(ValueError):
        lj.DCT().process(lj.Packet([1, 2, 3]))
    with pytest.raises(ValueError):
        lj.YCbCr444to422().process(lj.Packet([1, 2, 3]))
    with pytest.raises(ValueError):
        lj.RunLength().process(lj.Packet([]))
    assert lj.RGB2YCbCr().process(lj.Packet([])) == []


def test_dct_fixed_mode_operator(lj, oracle_c, golden):
    """[ext] DCT(mode="fixed"): the gateware datapath's arithmetic (new_dct6.py:171-396, corrected)."""
    from oracle import gateware as G
    rng = np.random.default_rng(3)
    blocks = np.concatenate([rng.integers(0, 256, (500, 64)), golden.stages["dct_in"][:100],
                             np.full((1, 64), 0), np.full((1, 64), 255)])
    op = lj.DCT(mode="fixed")
    got = op.process(torch.tensor(blocks, dtype=torch.int16).cuda()).cpu().numpy()
    assert np.array_equal(got, G.dct_fixed(blocks))
    pk = op.process(lj.Packet(blocks[0].tolist()))
    assert lj.to_signed(pk, 12) == G.dct_fixed(blocks[0]).tolist()
    with pytest.raises(ValueError):
        lj.DCT(mode="gateware")


def test_ingest_layouts_on_device(lj, tmp_path):
    """[ext] b2j_ingest / lj.to_rgb: BGR, RGBA, BGRA, planar and grey buffers (with row, plane and frame
    strides) become the packed RGB the path reads; read_image() follows the model's file reading
    (enc_frame.py:7-10: channels 0..2 of the decoded array)."""
    from PIL import Image
    g = torch.Generator().manual_seed(9)
    for (n, h, w) in ((1, 5, 7), (3, 16, 33), (2, 31, 64)):
        rgb = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
        a = torch.randint(0, 256, (n, h, w, 1), generator=g, dtype=torch.uint8)
        dev = rgb.cuda()
        assert torch.equal(lj.to_rgb(dev, "rgb"), dev)
        assert torch.equal(lj.to_rgb(dev.flip(-1).contiguous(), "bgr"), dev)
        assert torch.equal(lj.to_rgb(torch.cat([rgb, a], -1).cuda(), "rgba"), dev)
        assert torch.equal(lj.to_rgb(torch.cat([rgb.flip(-1), a], -1).cuda(), "bgra"), dev)
        assert torch.equal(lj.to_rgb(dev.permute(0, 3, 1, 2).contiguous(), "planar"), dev)
        grey = dev[..., 0].contiguous()
        assert torch.equal(lj.to_rgb(grey, "gray"), grey[..., None].expand(-1, -1, -1, 3))
        # strided views: a crop of a wider / taller buffer, every second frame
        big = torch.randint(0, 256, (2 * n, h + 3, w + 5, 3), generator=g, dtype=torch.uint8).cuda()
        view = big[::2, 1:h + 1, 2:w + 2]
        assert torch.equal(lj.to_rgb(view, "bgr"), view.flip(-1).contiguous())
        pl = torch.randint(0, 256, (n, 3, h + 2, w + 4), generator=g, dtype=torch.uint8).cuda()[:, :, 1:h + 1, 3:w + 3]
        assert torch.equal(lj.to_rgb(pl, "planar"), pl.permute(0, 2, 3, 1).contiguous())
        assert torch.equal(lj.to_rgb(dev[0], "rgb"), dev[0])
    arr = torch.randint(0, 256, (24, 40, 4), generator=g, dtype=torch.uint8).numpy()
    Image.fromarray(arr, "RGBA").save(str(tmp_path / "a.png"))
    Image.fromarray(arr[..., 0], "L").save(str(tmp_path / "l.png"))
    assert torch.equal(lj.read_image(str(tmp_path / "a.png")).cpu(), torch.from_numpy(arr[..., :3].copy()))
    assert torch.equal(lj.read_image(str(tmp_path / "l.png")).cpu(), torch.from_numpy(arr[..., :1].repeat(3, axis=2)))
    with pytest.raises(ValueError):
        lj.to_rgb(torch.zeros((4, 4, 3), dtype=torch.uint8).cuda(), "yuv")


def test_integration_md_ctypes_stub_runs_as_written(lj, oracle_c, golden):
    """INTEGRATION.md section 2 is a ctypes stub a reference maintainer would drop under test/: the code
    blocks are executed verbatim (only the library path is resolved) and checked against the oracle."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    sec = text[text.index("## 2. Bind the C ABI directly"):text.index("## 3. From C")]
    code = "\n".join(re.findall(r"```python\n(.*?)```", sec, flags=re.S))
    code = code.replace('ctypes.CDLL("libb2j.so")', 'ctypes.CDLL(%r)' % lj._lib.LIB_PATH)
    cudart = [p for p in os.popen("ldd %s" % lj._lib.LIB_PATH).read().split() if "libcudart" in p and "/" in p]
    if cudart:
        code = code.replace('ctypes.CDLL("libcudart.so")', 'ctypes.CDLL(%r)' % cudart[0])
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    blk = golden.stages["dct_in"][3].tolist()
    out = ns["dct_packet"](blk)
    assert lj.to_signed(out, 12) == golden.stages["dct_out"][3].tolist()
    zz = ns["zigzag_packet"](list(range(64)))
    assert zz == M.zigzag(np.arange(64)).tolist()
    rgb = golden.lena["rgb"][None, :64, :96].copy()
    files = ns["encode_frames"](rgb, 444, 85, 0)
    assert files[0] == oracle_c.encode_jpeg(rgb[0], "444", 85, 0)
