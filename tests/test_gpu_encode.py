"""The fused encode path K1..K4 on the GPU against the oracle: coefficients, bit counts,
offsets and the final JFIF bytes must be identical.  All calls go through the C ABI."""
import ctypes
import io

import numpy as np
import pytest
import torch

from oracle import jpeg_ext as J
from oracle import model_np as M

pytestmark = pytest.mark.gpu

SUBS = ["444", "422", "420"]
VARIANTS = ("simple", "tiled", "stream", "pair", "pair_wide")      # every K1 kernel of b2j_fused_block (b2j.h B2J_FUSED_*)


def synth(lj):
    import importlib
    return importlib.import_module("litejpeg_b200.synth")


def test_lena_config1_coefficients_are_the_models(lj, golden):
    """BASELINE config 1: lena 512x512, 4:4:4, model tables: every block == enc_frame.py output."""
    rgb = torch.from_numpy(golden.lena["rgb"]).cuda()
    enc = lj.JpegEncoder("444", tables=(M.y_quant_table.tolist(), M.cbcr_quant_table.tolist()))
    zz = golden.lena["zigzag"]
    for variant in VARIANTS:
        coefs = enc.coefficients(rgb, variant=variant).cpu().numpy()[0].reshape(4096, 3, 64)
        for c in range(3):
            assert np.array_equal(coefs[:, c], zz[c]), (variant, c)
    enc85 = lj.JpegEncoder("444", quality=85)                        # q85 == the model's tables
    assert np.array_equal(enc85.coefficients(rgb).cpu().numpy(), enc.coefficients(rgb).cpu().numpy())


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (37, 53), (96, 120), (64, 1040), (40, 2064), (17, 1025), (33, 272), (24, 31),
                                   (50, 544)])
def test_coefficients_any_geometry(lj, oracle_c, sub, shape):
    """Edge replication, ragged sides, several tiles / regions per row, unaligned row strides (widths that are
    not a multiple of 16 bytes take the byte-wise loaders), widths around the 32- and 256-pixel regions."""
    h, w = shape
    g = torch.Generator().manual_seed(h * 10007 + w)
    rgb = torch.randint(0, 256, (2, h, w, 3), generator=g, dtype=torch.uint8)
    qy, qc = J.quant_tables(75)
    enc = lj.JpegEncoder(sub, quality=75)
    exp = np.stack([oracle_c.coefficients(rgb[i].numpy(), sub, qy, qc).reshape(-1, 64) for i in range(2)])
    for variant in VARIANTS:
        got = enc.coefficients(rgb.cuda(), variant=variant).cpu().numpy()
        assert np.array_equal(got, exp), variant


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("zrl", ["eager", "standard"])
def test_stagewise_entropy_path(lj, oracle_c, golden, sub, zrl):
    """K2 bits, K3 offsets, K4 bytes, each against the oracle, on a lena crop (2 frames)."""
    crop = golden.lena["rgb"][64:64 + 200, 32:32 + 328]
    rgb = torch.from_numpy(np.stack([crop, crop[::-1].copy()])).cuda()
    enc = lj.JpegEncoder(sub, quality=85, zrl=zrl)
    desc = enc.desc_for(rgb)
    coefs = enc.coefficients(rgb)
    bits = lj.Huffman(zrl).process(coefs, desc)
    offsets, total = lj.BitPack(zrl).scan(bits, desc)
    out, sizes = lj.BitPack(zrl).process(coefs, bits, desc, enc.qy, enc.qc)
    zi = 0 if zrl == "eager" else 1
    for f in range(2):
        frame = rgb[f].cpu().numpy()
        c = coefs[f].cpu().numpy()
        ebits = oracle_c.block_bits(c, 200, 328, sub, zi)
        assert np.array_equal(bits[f].cpu().numpy(), ebits)
        eoff = np.cumsum(ebits) - ebits
        assert np.array_equal(offsets[f].cpu().numpy().astype(np.uint32), eoff.astype(np.uint32))
        assert int(total[f]) == int(ebits.sum())
        ref = oracle_c.encode_jpeg(frame, sub, 85, zi)
        n = int(sizes[f])
        assert n == len(ref)
        assert out[f, :n].cpu().numpy().tobytes() == ref
    # the fused single-pass entropy kernel (what b2j_encode uses) gives the same files
    L = lj.lib()
    cap = out.shape[1]
    ws = torch.empty((L.b2j_entropy_workspace_bytes(ctypes.byref(desc), cap) + 256,), dtype=torch.uint8, device="cuda")
    out2 = torch.zeros_like(out)
    sizes2 = torch.zeros_like(sizes)
    lj._lib.check(L.b2j_entropy_encode(coefs.data_ptr(), ctypes.byref(desc), lj._lib.table_arg(enc.qy), lj._lib.table_arg(enc.qc),
                                       zi, (ws.data_ptr() + 255) & ~255, out2.data_ptr(), cap, sizes2.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    assert torch.equal(sizes2, sizes)
    for f in range(2):
        assert torch.equal(out2[f, :int(sizes[f])], out[f, :int(sizes[f])])


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("kind", ["photo", "noise", "grey"])
def test_full_encode_bytes_equal_oracle(lj, oracle_c, sub, kind):
    s = synth(lj)
    frames = torch.stack([s.frame(120, 200, 9, i, kind) for i in range(3)])
    for q, zrl in ((75, "eager"), (90, "standard"), (100, "eager"), (10, "eager")):
        enc = lj.JpegEncoder(sub, quality=q, zrl=zrl, out_cap=120 * 200 * 8)
        files = enc.encode(frames.cuda())
        for i, f in enumerate(files):
            assert f == oracle_c.encode_jpeg(frames[i].numpy(), sub, q, 0 if zrl == "eager" else 1), (q, zrl, i)


@pytest.mark.parametrize("sub", SUBS)
def test_flat_frames_short_final_blocks(lj, oracle_c, sub):
    """Flat frames: every block after the first is 4-6 bits long, so the frame's final 32-bit word
    (with the 1-padding) is composed by a thread that owns none of the last blocks."""
    for (h, w) in ((8, 8), (16, 16), (24, 40), (136, 200), (264, 1032)):
        frames = torch.empty(5, h, w, 3, dtype=torch.uint8)
        for i, rgb in enumerate(((0, 0, 0), (255, 255, 255), (128, 128, 128), (200, 30, 90), (1, 254, 7))):
            frames[i] = torch.tensor(rgb, dtype=torch.uint8)
        frames[3, h - 1, w - 1, 0] = 17          # one busy block at the very end of one frame
        for zrl in ("eager", "standard"):
            enc = lj.JpegEncoder(sub, quality=75, zrl=zrl, out_cap=h * w * 8 + 4096)
            files = enc.encode(frames.cuda())
            for i, f in enumerate(files):
                assert f == oracle_c.encode_jpeg(frames[i].numpy(), sub, 75, 0 if zrl == "eager" else 1), (h, w, zrl, i)


def test_colour_exception_inputs_through_the_fused_kernel(lj, oracle_c):
    """Every input where the float64 colour model differs from the exact rational value (3464 for
    Y, 688 for Cr) plus look-alike candidates, laid out as an image: the fused kernel's patch-up
    path must reproduce the model for all of them, in every subsampling mode."""
    allv = np.arange(1 << 24, dtype=np.int64)
    R, G, B = allv >> 16, (allv >> 8) & 255, allv & 255
    Y, CB, CR = M.rgb2ycbcr(R, G, B)
    ny = 299 * R + 587 * G + 114 * B
    ncr = 5000 * R - 4187 * G - 813 * B + 1280000
    sel = np.nonzero((ny % 1000 == 0) | ((ncr % 10000 == 0) & (G >= 150)))[0]
    assert (Y[sel] != ny[sel] // 1000).sum() == 3464 and (CR[sel] != ncr[sel] // 10000).sum() == 688
    rng = np.random.default_rng(11)
    sel = np.concatenate([sel, sel[rng.permutation(sel.size)[:4096]]])        # some twice, in another order
    w = 256
    h = (sel.size + w - 1) // w
    h += (-h) % 16
    img = np.zeros((h * w, 3), np.uint8)
    img[:sel.size, 0], img[:sel.size, 1], img[:sel.size, 2] = R[sel], G[sel], B[sel]
    img[sel.size:] = img[:h * w - sel.size]
    img = img.reshape(h, w, 3)
    qy, qc = J.quant_tables(100)                                              # divisors 1: nothing is hidden by rounding
    for sub in SUBS:
        enc = lj.JpegEncoder(sub, quality=100)
        exp = oracle_c.coefficients(img, sub, qy, qc).reshape(-1, 64)
        for variant in ("tiled", "simple"):
            got = enc.coefficients(torch.from_numpy(img).cuda(), variant=variant).cpu().numpy()[0]
            assert np.array_equal(got, exp), (sub, variant)


def test_colour_exception_inputs_as_flat_tiles_through_every_variant(lj, oracle_c):
    """The same inputs as 16x16 tiles of one colour each, at quality 90 (DC divisors 3: a tile whose Y or Cr is off by one
    changes its DC coefficient): reaches the patch-up path of the stream and pair kernels, which only run for tables
    without a divisor 1, and the grey levels (every grey pixel is a candidate of both components)."""
    allv = np.arange(1 << 24, dtype=np.int64)
    R, G, B = allv >> 16, (allv >> 8) & 255, allv & 255
    ny = 299 * R + 587 * G + 114 * B
    ncr = 5000 * R - 4187 * G - 813 * B + 1280000
    sel = np.nonzero((ny % 1000 == 0) | ((ncr % 10000 == 0) & (G >= 150)))[0]
    greys = np.arange(256, dtype=np.int64) * 0x010101
    sel = np.concatenate([sel, greys])
    tiles_x = 128
    tiles_y = (sel.size + tiles_x - 1) // tiles_x
    colours = np.zeros((tiles_y * tiles_x, 3), np.uint8)
    colours[:sel.size, 0], colours[:sel.size, 1], colours[:sel.size, 2] = R[sel], G[sel], B[sel]
    colours[sel.size:] = colours[:tiles_y * tiles_x - sel.size]
    img = np.repeat(np.repeat(colours.reshape(tiles_y, tiles_x, 3), 16, axis=0), 16, axis=1)
    qy, qc = J.quant_tables(90)
    assert qy[0] == 3 and qc[0] == 3
    dev = torch.from_numpy(img).cuda()
    for sub in SUBS:
        enc = lj.JpegEncoder(sub, quality=90)
        exp = oracle_c.coefficients(img, sub, qy, qc).reshape(-1, 64)
        for variant in ("stream", "pair", "pair_wide", "auto"):
            got = enc.coefficients(dev, variant=variant).cpu().numpy()[0]
            assert np.array_equal(got, exp), (sub, variant)


@pytest.mark.parametrize("layout", ["bgr", "rgba", "bgra"])
def test_source_layouts_are_read_in_place(lj, oracle_c, layout):
    """[ext] SURVEY 8f-4: BGR / RGBA / BGRA frames go straight into K1 (B2J_DESC_LAYOUT): coefficients and files equal
    those of the RGB frames, for every subsampling, ragged sizes (edge regions, unaligned rows), a strided crop, the
    kernels that read other layouts (pair, simple; AUTO) and the host-buffer encoder."""
    s = synth(lj)
    rng = np.random.default_rng(3)

    def convert(rgb):                                   # (N, H, W, 3) uint8 tensor -> the layout under test
        a = rgb.numpy()
        if layout in ("bgr", "bgra"):
            a = a[..., ::-1]
        if layout in ("rgba", "bgra"):
            a = np.concatenate([a, rng.integers(0, 256, a.shape[:-1] + (1,), dtype=np.uint8)], axis=-1)
        return torch.from_numpy(np.ascontiguousarray(a))

    for sub in SUBS:
        for (h, w) in ((64, 128), (73, 200), (160, 1000)):
            rgb = s.batch(2, h, w, 9, "photo")
            src = convert(rgb)
            exp = np.stack([oracle_c.coefficients(rgb[i].numpy(), sub, *J.quant_tables(75)) for i in range(2)])
            enc = lj.JpegEncoder(sub, quality=75, layout=layout)
            for variant in ("auto", "pair", "simple"):
                got = enc.coefficients(src.cuda(), variant=variant).cpu().numpy()
                assert np.array_equal(got.reshape(exp.shape), exp), (sub, h, w, variant)
            files = enc.encode(src.cuda())
            ref = lj.JpegEncoder(sub, quality=75).encode(rgb.cuda())
            assert files == ref, (sub, h, w)
    # gateware DCT and a divisor 1 are outside the pair variant: the one-thread-per-block kernel reads the layout
    rgb = s.batch(1, 48, 80, 10, "noise")
    src = convert(rgb).cuda()
    assert lj.JpegEncoder("420", quality=100, layout=layout).encode(src) == lj.JpegEncoder("420", quality=100).encode(rgb.cuda())
    assert lj.JpegEncoder("422", quality=60, dct="fixed", layout=layout).encode(src) == \
        lj.JpegEncoder("422", quality=60, dct="fixed").encode(rgb.cuda())
    # host-buffer encoder: frames of the layout in, the same files out
    rgb = s.batch(3, 96, 160, 11, "photo")
    out, sizes = lj.JpegEncoder("420", quality=75, layout=layout).encode_host(convert(rgb))
    ref = lj.JpegEncoder("420", quality=75).encode(rgb.cuda())
    for i in range(3):
        assert bytes(out[i, :int(sizes[i])].numpy()) == ref[i]
    # a crop of a larger buffer through the strides of the descriptor
    big = convert(s.batch(1, 128, 256, 12, "photo"))[0].cuda()
    bpp = big.shape[-1]
    desc = lj.make_desc(72, 104, "420", 1, row_stride=256 * bpp, frame_stride=256 * bpp * 128, layout=layout)
    g = lj.geometry(desc)
    qy, qc = lj.quant_tables(75)
    coefs = torch.empty((1, g.blocks_per_frame, 64), dtype=torch.int16, device="cuda")
    lj._lib.check(lj.lib().b2j_fused_block(big.data_ptr() + (16 * 256 + 32) * bpp, ctypes.byref(desc), lj._lib.table_arg(qy),
                                           lj._lib.table_arg(qc), coefs.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
    crop = big[16:88, 32:136].cpu().numpy()
    crop = crop[..., 2::-1] if layout in ("bgr", "bgra") else crop[..., :3]
    exp = oracle_c.coefficients(np.ascontiguousarray(crop), "420", *J.quant_tables(75))
    assert np.array_equal(coefs.cpu().numpy()[0], exp.reshape(-1, 64))


def test_lena_files_decode(lj, oracle_c, golden):
    """The emitted files open in PIL / OpenCV and the oracle decoder recovers the model's coefficients."""
    from PIL import Image
    rgb = golden.lena["rgb"]
    for sub in SUBS:
        enc = lj.JpegEncoder(sub, quality=85)
        jpg = enc.encode(torch.from_numpy(rgb).cuda())[0]
        assert jpg[:2] == b"\xFF\xD8" and jpg[-2:] == b"\xFF\xD9"
        pil = np.array(Image.open(io.BytesIO(jpg)).convert("RGB")).astype(np.float64)
        psnr = 10 * np.log10(255 ** 2 / np.mean((pil - rgb) ** 2))
        assert psnr > 32, (sub, psnr)
        rc, dec = oracle_c.decode_coefficients(jpg, 512, 512, sub)
        assert rc == 0
        assert np.array_equal(dec.reshape(-1, 64), enc.coefficients(torch.from_numpy(rgb).cuda()).cpu().numpy()[0])
        if sub == "444":
            zz = golden.lena["zigzag"]
            for c in range(3):
                assert np.array_equal(dec[:, c], zz[c])
        try:
            import cv2
            assert cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR).shape == (512, 512, 3)
        except ImportError:
            pass


@pytest.mark.parametrize("cfg", [(1080, 1920, "420", 75, 2), (2160, 3840, "444", 90, 4), (2160, 3840, "420", 90, 4),
                                 (4320, 7680, "420", 75, 5)])
def test_full_size_frames_bytes_equal_oracle(lj, oracle_c, cfg):
    """BASELINE configs 2-5 at full frame size (one frame each): bytes identical to the C oracle,
    and the oracle decoder recovers the kernel's coefficients from the file."""
    h, w, sub, q, idx = cfg
    s = synth(lj)
    frame = s.frame(h, w, idx, 0, "photo")
    enc = lj.JpegEncoder(sub, quality=q)
    dev = frame.cuda()[None]
    jpg = enc.encode(dev)[0]
    ref = oracle_c.encode_jpeg(frame.numpy(), sub, q, 0)
    assert len(jpg) == len(ref)
    assert jpg == ref
    rc, dec = oracle_c.decode_coefficients(jpg, h, w, sub)
    assert rc == 0 and np.array_equal(dec.reshape(-1, 64), enc.coefficients(dev).cpu().numpy()[0])


def test_stress_8k_noise_and_grey(lj, oracle_c):
    """SURVEY 8d stress distributions on one 8K frame: worst-case scan / bit packer, and EOB-only."""
    s = synth(lj)
    for kind in ("noise", "grey"):
        frame = s.frame(4320, 7680, 5, 0, kind)
        enc = lj.JpegEncoder("420", quality=75, out_cap=4320 * 7680 * 3)
        jpg = enc.encode(frame.cuda()[None])[0]
        assert jpg == oracle_c.encode_jpeg(frame.numpy(), "420", 75, 0), kind


def test_batch_properties(lj, oracle_c):
    """Size-independent properties on a batch: frames are independent (same frame -> same bytes
    wherever it sits in the batch), and a checksum of checksums equals the per-frame oracle."""
    import hashlib
    s = synth(lj)
    uniq = torch.stack([s.frame(1080, 1920, 3, i) for i in range(4)])
    order = [0, 1, 2, 3, 3, 1, 0, 2, 2, 2, 0, 1]
    batch = uniq[order].cuda()
    enc = lj.JpegEncoder("420", quality=75)
    files = enc.encode(batch)
    ref = [oracle_c.encode_jpeg(uniq[i].numpy(), "420", 75, 0) for i in range(4)]
    for pos, i in enumerate(order):
        assert files[pos] == ref[i], pos
    top = hashlib.sha256(b"".join(hashlib.sha256(f).digest() for f in files)).hexdigest()
    assert top == hashlib.sha256(b"".join(hashlib.sha256(ref[i]).digest() for i in order)).hexdigest()


def test_capacity_error_is_reported(lj):
    s = synth(lj)
    frame = s.frame(64, 64, 1, 0, "noise").cuda()[None]
    enc = lj.JpegEncoder("444", quality=95, out_cap=1024)
    out, sizes = enc.encode_device(frame)
    assert int(sizes[0]) < 0 and -int(sizes[0]) > 1024
    with pytest.raises(lj.B2JError):
        enc.encode(frame)
    big = lj.JpegEncoder("444", quality=95, out_cap=-int(sizes[0]) + 64)
    assert len(big.encode(frame)[0]) <= -int(sizes[0])


def test_host_buffer_encoder_matches_device_path(lj, oracle_c):
    """The end-to-end entry point (host buffers, chunked copies) gives the same files."""
    s = synth(lj)
    frames = torch.stack([s.frame(360, 640, 6, i) for i in range(7)])
    enc = lj.JpegEncoder("420", quality=75)
    ref = [oracle_c.encode_jpeg(frames[i].numpy(), "420", 75, 0) for i in range(7)]
    for pinned in (False, True):
        src = frames.pin_memory() if pinned else frames
        out, sizes = enc.encode_host(src, frames_per_chunk=3)
        for i in range(7):
            assert out[i, :int(sizes[i])].numpy().tobytes() == ref[i], (pinned, i)
    enc.close()


def test_multi_device_encoder_one_call(lj, oracle_c):
    """b2j_multi_encoder_*: one host batch sharded over a device list inside ONE call.  A 1-GPU box lists
    device 0 several times (every entry gets its own worker thread, streams and buffers), which exercises
    the same sharding, threading and joining code as N distinct GPUs."""
    s = synth(lj)
    n = 11
    frames = torch.stack([s.frame(200, 328, 8, i) for i in range(n)]).pin_memory()
    ref = [oracle_c.encode_jpeg(frames[i].numpy(), "420", 75, 0) for i in range(n)]
    ndev = torch.cuda.device_count()
    for devices in ([0], [0, 0, 0], list(range(ndev)) * 2):
        enc = lj.MultiGpuJpegEncoder(devices, 200, 328, "420", quality=75, frames_per_chunk=2)
        for _ in range(2):
            out, sizes = enc.encode_host(frames)
            for i in range(n):
                assert out[i, :int(sizes[i])].numpy().tobytes() == ref[i], (devices, i)
        few, fsz = enc.encode_host(frames[:2].contiguous())                 # fewer frames than workers: empty shards
        assert few[1, :int(fsz[1])].numpy().tobytes() == ref[1]
        assert enc.h2d_seconds(frames, repeats=2) > 0.0
        enc.close()
    with pytest.raises(lj.B2JError):
        lj.MultiGpuJpegEncoder([ndev + 7], 200, 328)


def test_one_encoder_on_two_streams(lj, oracle_c):
    """one JpegEncoder used on two streams at once must not share a workspace between the calls."""
    s = synth(lj)
    a = torch.stack([s.frame(360, 640, 71, i) for i in range(6)]).cuda()
    b = torch.stack([s.frame(360, 640, 72, i) for i in range(6)]).cuda()
    ref_a = [oracle_c.encode_jpeg(a[i].cpu().numpy(), "420", 75, 0) for i in range(6)]
    ref_b = [oracle_c.encode_jpeg(b[i].cpu().numpy(), "420", 75, 0) for i in range(6)]
    enc = lj.JpegEncoder("420", quality=75)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):
        oa, za = enc.encode_device(a, stream=s1)
        ob, zb = enc.encode_device(b, stream=s2)
        torch.cuda.synchronize()
        for i in range(6):
            assert oa[i, :int(za[i])].cpu().numpy().tobytes() == ref_a[i]
            assert ob[i, :int(zb[i])].cpu().numpy().tobytes() == ref_b[i]
    with pytest.raises(ValueError):
        enc.encode_device(a, out=torch.empty((6, 100), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        enc.encode_device(a, sizes=torch.empty((5,), dtype=torch.int64, device="cuda"))


def test_strided_input_through_the_c_abi(lj, oracle_c):
    """row_stride / frame_stride of b2j_image_desc: a crop of a larger device buffer, no copy."""
    s = synth(lj)
    big = s.frame(128, 256, 4, 0).cuda()
    crop_h, crop_w, y0, x0 = 72, 104, 16, 32
    desc = lj.make_desc(crop_h, crop_w, "420", 1, row_stride=256 * 3, frame_stride=256 * 3 * 128)
    g = lj.geometry(desc)
    qy, qc = lj.quant_tables(75)
    coefs = torch.empty((1, g.blocks_per_frame, 64), dtype=torch.int16, device="cuda")
    base = big.data_ptr() + (y0 * 256 + x0) * 3
    for variant in (1, 2, 3, 4):
        coefs.fill_(0x5A5A)
        lj._lib.check(lj.lib().b2j_fused_block(base, ctypes.byref(desc), lj._lib.table_arg(qy), lj._lib.table_arg(qc),
                                               coefs.data_ptr(), variant, torch.cuda.current_stream().cuda_stream))
        exp = oracle_c.coefficients(big[y0:y0 + crop_h, x0:x0 + crop_w].cpu().numpy(), "420", *J.quant_tables(75))
        assert np.array_equal(coefs.cpu().numpy()[0], exp.reshape(-1, 64)), variant


@pytest.mark.parametrize("sub", SUBS)
def test_fixed_dct_mode_equals_its_oracle(lj, oracle_c, sub):
    """[ext] dct="fixed" (gateware fixed-point arithmetic): bit-exact against its own
    restatement through both K1 variants and the whole encode; a valid file; near the model's."""
    from PIL import Image
    s = synth(lj)
    frames = torch.stack([s.frame(136, 1048, 11, i) for i in range(2)])
    enc = lj.JpegEncoder(sub, quality=75, dct="fixed", out_cap=136 * 1048 * 6)
    qy, qc = oracle_c.quant_tables(75)
    for variant in ("tiled", "simple"):
        got = enc.coefficients(frames.cuda(), variant=variant).cpu().numpy()
        for i in range(2):
            want = oracle_c.coefficients(frames[i].numpy(), sub, qy, qc, dct_fixed=True).reshape(-1, 64)
            assert np.array_equal(got[i], want), (variant, i)
    files = enc.encode(frames.cuda())
    model_files = lj.JpegEncoder(sub, quality=75, out_cap=136 * 1048 * 6).encode(frames.cuda())
    for i in range(2):
        assert files[i] == oracle_c.encode_jpeg(frames[i].numpy(), sub, 75, 0 | 0x100)
        a = np.asarray(Image.open(io.BytesIO(files[i])).convert("RGB")).astype(np.float64)
        b = np.asarray(Image.open(io.BytesIO(model_files[i])).convert("RGB")).astype(np.float64)
        assert a.shape == (136, 1048, 3)
        psnr = 10 * np.log10(255 ** 2 / np.mean((a - b) ** 2))
        assert psnr > 35, psnr                         # the two arithmetic modes decode to nearly the same picture
    host = lj.JpegEncoder(sub, quality=75, dct="fixed", out_cap=136 * 1048 * 6)
    out_h, sizes_h = host.encode_host(frames)
    assert bytes(out_h[1, :int(sizes_h[1])].numpy().tobytes()) == files[1]
    host.close()


@pytest.mark.parametrize("sub", SUBS)
def test_restart_intervals_bytes_equal_oracle(lj, oracle_c, sub):
    """[ext] restart intervals: DRI + RSTm, predictors restart, every interval an independent
    byte-aligned stream.  Bytes equal the oracle for short, ragged, row-sized and oversized intervals."""
    s = synth(lj)
    h, w = 72, 200
    frames = torch.stack([s.frame(h, w, 13, i, kind) for i, kind in enumerate(("photo", "noise", "grey"))])
    g = oracle_c.geometry(h, w, sub)
    n_mcu, mcus_x = g["n_mcu"], g["mcus_x"]
    for zi, zrl in enumerate(("eager", "standard")):
        for ri in (1, 5, mcus_x, 2 * mcus_x + 3, n_mcu - 1, n_mcu, 30000):
            enc = lj.JpegEncoder(sub, quality=80, zrl=zrl, restart_interval=ri, out_cap=h * w * 8)
            files = enc.encode(frames.cuda())
            for i, f in enumerate(files):
                assert f == oracle_c.encode_jpeg(frames[i].numpy(), sub, 80, zi | (ri << 16)), (zrl, ri, i)


def test_restart_intervals_full_size_and_host_path(lj, oracle_c):
    """1080p 4:2:0 with one interval per MCU row (120 MCUs): 68 segments per frame, through the
    device path, the host-buffer encoder and together with dct="fixed"; libjpeg decodes the file
    to the same pixels as the file without restart markers."""
    import io as _io
    from PIL import Image
    s = synth(lj)
    frames = torch.stack([s.frame(1080, 1920, 2, i) for i in range(3)])
    enc = lj.JpegEncoder("420", quality=75, zrl="standard", restart_interval=120)
    files = enc.encode(frames.cuda())
    plain = lj.JpegEncoder("420", quality=75, zrl="standard").encode(frames[:1].cuda())[0]
    for i in range(3):
        assert files[i] == oracle_c.encode_jpeg(frames[i].numpy(), "420", 75, 1 | (120 << 16)), i
    a = np.asarray(Image.open(_io.BytesIO(files[0])).convert("RGB"))
    b = np.asarray(Image.open(_io.BytesIO(plain)).convert("RGB"))
    assert np.array_equal(a, b)
    out_h, sizes_h = enc.encode_host(frames, frames_per_chunk=2)
    for i in range(3):
        assert bytes(out_h[i, :int(sizes_h[i])].numpy().tobytes()) == files[i]
    enc.close()
    fx = lj.JpegEncoder("420", quality=75, zrl="eager", dct="fixed", restart_interval=7)
    assert fx.encode(frames[:1].cuda())[0] == oracle_c.encode_jpeg(frames[0].numpy(), "420", 75, 0 | 0x100 | (7 << 16))
    with pytest.raises(ValueError):
        lj.JpegEncoder("420", restart_interval=40000)


@pytest.mark.parametrize("sub", SUBS)
def test_one_frame_encoded_as_bands_equals_the_whole_frame(lj, oracle_c, sub):
    """[ext] a frame split into bands of whole MCU rows (the unit one GPU would take), each band
    encoded on its own with the right restart-marker phase, joined into one file: byte-identical
    to the whole frame with one restart interval per MCU row."""
    s = synth(lj)
    for (h, w) in ((203, 328), (1080, 1920)):
        frame = s.frame(h, w, 17, 0)
        mcus_x = oracle_c.geometry(h, w, sub)["mcus_x"]
        whole = oracle_c.encode_jpeg(frame.numpy(), sub, 75, 0 | (mcus_x << 16))
        enc = lj.JpegEncoder(sub, quality=75)
        for n_bands in ((1, 2, 3, 5, 8) if h < 1000 else (8,)):
            assert enc.encode_frame_bands(frame.cuda(), n_bands) == whole, (h, w, n_bands)


def test_encode_is_cuda_graph_capturable(lj, oracle_c):
    """The whole path (memset + 4 kernels) only enqueues work on the caller's stream, so it can be
    captured once into a CUDA graph and replayed per frame (the launch-bound single-frame case)."""
    s = synth(lj)
    frames = [s.frame(360, 640, 19, i) for i in range(3)]
    enc = lj.JpegEncoder("420", quality=75)
    rgb = frames[0].cuda()[None].clone()
    out, sizes = enc.encode_device(rgb)                     # warm-up outside the capture (workspace allocation)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        enc.encode_device(rgb, out=out, sizes=sizes, stream=torch.cuda.current_stream())
    for f in frames:
        rgb.copy_(f.cuda()[None])
        g.replay()
        torch.cuda.synchronize()
        n = int(sizes[0])
        assert bytes(out[0, :n].cpu().numpy().tobytes()) == oracle_c.encode_jpeg(f.numpy(), "420", 75, 0)


def test_out_of_range_coefficients_stay_memory_safe(lj):
    """b2j_entropy_encode is a public entry point: coefficients no baseline table can code (|v| up to
    32767, outside the contract stated in include/b2j.h) give an undefined stream or an error verdict,
    but never writes past the per-block scratch columns, and the workspace stays usable for the next,
    valid call."""
    L = lj.lib()
    desc = lj.make_desc(64, 96, "444", 2)
    g = lj.geometry(desc)
    cap = 64 * 96 * 8
    gen = torch.Generator().manual_seed(5)
    bad = torch.randint(-32768, 32768, (2, g.blocks_per_frame, 64), generator=gen, dtype=torch.int32).to(torch.int16).cuda()
    ws = torch.empty((L.b2j_entropy_workspace_bytes(ctypes.byref(desc), cap) + 256,), dtype=torch.uint8, device="cuda")
    guard = torch.full((1 << 20,), 0x5A, dtype=torch.uint8, device="cuda")           # canary right behind the workspace arena
    wp = (ws.data_ptr() + 255) & ~255
    out = torch.zeros((2, cap), dtype=torch.uint8, device="cuda")
    sizes = torch.zeros((2,), dtype=torch.int64, device="cuda")
    qy, qc = lj.quant_tables(75)
    ta, tb = lj._lib.table_arg(qy), lj._lib.table_arg(qc)
    for zrl in (0, 1):
        lj._lib.check(L.b2j_entropy_encode(bad.data_ptr(), ctypes.byref(desc), ta, tb, zrl, wp, out.data_ptr(), cap, sizes.data_ptr(), None))
        torch.cuda.synchronize()
        assert bool(((sizes < 0) | (sizes <= cap)).all())
    assert bool((guard == 0x5A).all())
    enc = lj.JpegEncoder("444", quality=75, out_cap=cap)
    good = enc.coefficients(synth(lj).frame(64, 96, 23, 0).cuda()[None].repeat(2, 1, 1, 1))
    lj._lib.check(L.b2j_entropy_encode(good.data_ptr(), ctypes.byref(desc), ta, tb, 0, wp, out.data_ptr(), cap, sizes.data_ptr(), None))
    torch.cuda.synchronize()
    assert (sizes > 0).all() and int(sizes[0]) == int(sizes[1])


@pytest.mark.parametrize("shape", [(1, 1), (1, 65535), (65535, 1), (3, 2)])
def test_extreme_dimensions_bytes_equal_oracle(lj, oracle_c, shape):
    """Smallest and largest baseline-JPEG dimensions (SOF0 holds 16-bit sizes): one pixel, one row of
    65535 pixels, one column of 65535 pixels; everything else is edge replication."""
    h, w = shape
    g = torch.Generator().manual_seed(h * 7 + w)
    frame = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
    for sub in SUBS:
        enc = lj.JpegEncoder(sub, quality=90, zrl="standard", out_cap=max(h, w) * 16 * 16 * 3 + 4096)
        jpg = enc.encode(frame.cuda()[None])[0]
        assert jpg == oracle_c.encode_jpeg(frame.numpy(), sub, 90, 1), (shape, sub)
    with pytest.raises(lj.B2JError):
        lj.geometry(lj.make_desc(65536, 8, "444", 1))


def test_random_parameter_space_bytes_equal_oracle(lj, oracle_c):
    """Seeded sweep over shape x subsampling x quality x ZRL policy x restart interval x DCT mode x
    content: every file byte-identical to the oracle's (batches of 1-3 frames)."""
    rng = np.random.default_rng(20171020)
    s = synth(lj)
    for case in range(48):
        h, w = int(rng.integers(1, 300)), int(rng.integers(1, 1300))
        sub = SUBS[case % 3]
        q = int(rng.integers(1, 101))
        zi = int(rng.integers(0, 2))
        ri = int(rng.choice([0, 0, 0, 1, 3, 11, 64, 500]))
        fixed = bool(case % 5 == 4)
        n = 1 + case % 3
        kind = ("photo", "noise", "grey", "photo")[case % 4]
        frames = torch.stack([s.frame(h, w, 31 + case, i, kind) for i in range(n)])
        if kind == "grey":
            frames[..., :] = int(rng.integers(0, 256))
        enc = lj.JpegEncoder(sub, quality=q, zrl=("eager", "standard")[zi], restart_interval=ri,
                             dct="fixed" if fixed else "model", out_cap=((h + 15) // 16 * 16) * ((w + 15) // 16 * 16) * 6 + 4096)
        files = enc.encode(frames.cuda())
        for i in range(n):
            want = oracle_c.encode_jpeg(frames[i].numpy(), sub, q, zi | (0x100 if fixed else 0) | (ri << 16))
            assert files[i] == want, (case, h, w, sub, q, zi, ri, fixed, kind, i)


def test_pure_c_client_of_the_host_buffer_encoder(lj, oracle_c, tmp_path):
    """tests/c/encode_raw.c: a C99 program that only knows include/b2j.h and libb2j.so (no Python, no
    torch in the process) encodes raw RGB frames from a file; the files equal the oracle's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "jpeg_encoder_gsoc2017_ishan-bansal_b200")
    exe = str(tmp_path / "encode_raw")
    res = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-I",
                          os.path.join(root, "include"), os.path.join(root, "tests", "c", "encode_raw.c"), "-o", exe, "-L", libdir,
                          "-l:libb2j.so", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    s = synth(lj)
    frames = torch.stack([s.frame(270, 484, 41, i) for i in range(6)])
    raw = str(tmp_path / "in.rgb")
    frames.numpy().tofile(raw)
    for sub, q, zrl, ri in (("420", 75, 0, 0), ("444", 92, 1, 0), ("422", 60, 1, 31)):
        prefix = str(tmp_path / ("out_%s_" % sub))
        run = subprocess.run([exe, raw, "484", "270", "6", sub, str(q), str(zrl), str(ri), prefix], capture_output=True, text=True)
        assert run.returncode == 0, (run.stdout, run.stderr)
        for i in range(6):
            got = open("%s%d.jpg" % (prefix, i), "rb").read()
            assert got == oracle_c.encode_jpeg(frames[i].numpy(), sub, q, zrl | (ri << 16)), (sub, i)


def test_concurrent_callers_on_separate_streams(lj, oracle_c):
    """The ABI contract says re-entrant, no global state, one stream per caller thread: four threads
    encode different batches concurrently, each on its own stream and workspace, five times over."""
    import threading
    s = synth(lj)
    jobs = [("420", 75, "eager", 0), ("444", 90, "standard", 0), ("422", 50, "eager", 9), ("420", 30, "standard", 0)]
    batches = [torch.stack([s.frame(200 + 8 * t, 328 - 16 * t, 51 + t, i) for i in range(3)]) for t in range(4)]
    want = [[oracle_c.encode_jpeg(batches[t][i].numpy(), jobs[t][0], jobs[t][1], (0 if jobs[t][2] == "eager" else 1) | (jobs[t][3] << 16))
             for i in range(3)] for t in range(4)]
    errors = []

    def worker(t):
        try:
            sub, q, zrl, ri = jobs[t]
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            enc = lj.JpegEncoder(sub, quality=q, zrl=zrl, restart_interval=ri, out_cap=300 * 400 * 6)
            dev = batches[t].cuda()
            for _ in range(5):
                with torch.cuda.stream(stream):
                    out, sizes = enc.encode_device(dev, stream=stream)
                stream.synchronize()
                host, sz = out.cpu(), sizes.cpu()
                for i in range(3):
                    if bytes(host[i, :int(sz[i])].numpy().tobytes()) != want[t][i]:
                        errors.append((t, i))
        except Exception as e:          # surfaced below
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("kind", ["noise", "photo", "grey"])
def test_no_byte_is_written_outside_the_stated_ranges(lj, kind):
    """Own bounds check (no sanitizer on this pool): guard bytes behind the workspace the library asked
    for, behind every output slot's used bytes, behind the last slot and behind `sizes` stay intact,
    for worst-case (noise) and best-case (grey) content, with and without restart intervals."""
    L = lj.lib()
    s = synth(lj)
    n, h, w = 3, 136, 1048
    frames = torch.stack([s.frame(h, w, 61, i, kind) for i in range(n)]).cuda()
    desc = lj.make_desc(h, w, "420", n)
    cap = 136 * 1048 * 3 + 4096
    qy, qc = lj.quant_tables(95)
    ta, tb = lj._lib.table_arg(qy), lj._lib.table_arg(qc)
    G = 1 << 16
    for opts in (0, 1, lj._lib.opt_restart(7), 1 | lj._lib.opt_restart(66) | 0x100):
        need = int(L.b2j_encode_workspace_bytes_opts(ctypes.byref(desc), cap, opts))
        ws = torch.full((need + 256 + G,), 0xA5, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        pad = wp - ws.data_ptr()
        out = torch.full((n * cap + G,), 0xC3, dtype=torch.uint8, device="cuda")
        sizes = torch.full((n + 8,), -12345, dtype=torch.int64, device="cuda")
        lj._lib.check(L.b2j_encode(frames.data_ptr(), ctypes.byref(desc), 0, ta, tb, opts, wp, need, out.data_ptr(), cap,
                                   sizes.data_ptr(), None))
        torch.cuda.synchronize()
        assert bool((ws[pad + need:] == 0xA5).all()), "write behind the workspace"
        assert bool((out[n * cap:] == 0xC3).all()), "write behind the last output slot"
        assert bool((sizes[n:] == -12345).all()), "write behind sizes[]"
        sz = sizes[:n].cpu()
        assert int(sz.min()) > 0
        for i in range(n):
            assert bool((out[i * cap + int(sz[i]):(i + 1) * cap] == 0xC3).all()), "write behind the file inside slot %d" % i
            assert bytes(out[i * cap:i * cap + 2].cpu().numpy().tobytes()) == b"\xFF\xD8"
            assert bytes(out[i * cap + int(sz[i]) - 2:i * cap + int(sz[i])].cpu().numpy().tobytes()) == b"\xFF\xD9"


def test_example_script_encodes_a_png(lj, oracle_c, golden, tmp_path):
    """examples/encode_image.py: PNG in, JPEG out, equal to the oracle's file and decodable by libjpeg."""
    import os
    import subprocess
    import sys as _sys
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rgb = golden.lena["rgb"][:200, :312]
    src, dst = str(tmp_path / "in.png"), str(tmp_path / "out.jpg")
    Image.fromarray(rgb).save(src)
    run = subprocess.run([_sys.executable, os.path.join(root, "examples", "encode_image.py"), src, dst, "--sub", "444",
                          "--quality", "85", "--zrl", "eager"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    got = open(dst, "rb").read()
    assert got == oracle_c.encode_jpeg(rgb, "444", 85, 0)
    assert Image.open(io.BytesIO(got)).size == (312, 200)
