"""[ext] PNG ingest on the device (b2j_png_decode / lj.decode_png / lj.read_image).
The parity anchor is what the reference does with the file: PIL decodes it (test/model/enc_frame.py:7,
test/common.py:47-51); every frame decoded on the GPU must equal np.array(Image.open(file)) bit for bit."""
import io
import os
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import png_util as U                                                   # noqa: E402

pytestmark = pytest.mark.gpu


def pil_rgb(data):
    from PIL import Image
    a = np.array(Image.open(io.BytesIO(data)))
    if a.ndim == 2:
        a = np.repeat(a[..., None], 3, axis=2)
    return a[..., :3]


def photo(h, w, c, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(w + h - 2, 1)), 255 - (xx % 256)], axis=2)
    a = base[..., :c] + rng.integers(-6, 7, (h, w, c))
    return np.clip(a, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY])
def test_every_filter_and_block_type(lj, c, strategy):
    """rows cycle through the five filter types; the zlib strategy picks the block types (dynamic, fixed, run-length
    matches at distance 1..4 = overlapping copies, literals only)."""
    a = photo(67, 45, c, 1 + c)
    files = [U.write_png(a, filters=[(r + k) % 5 for r in range(67)], strategy=strategy, level=lvl, idat_split=split)
             for k, (lvl, split) in enumerate([(6, 0), (9, 100), (1, 4096)])]
    got = lj.decode_png(files).cpu().numpy()
    for i, f in enumerate(files):
        assert np.array_equal(got[i], pil_rgb(f)), (c, strategy, i)
        assert np.array_equal(got[i], a[..., :3] if c > 1 else np.repeat(a, 3, axis=2))


def test_stored_blocks_and_flat_rows(lj):
    a = photo(40, 33, 3, 7)
    f0 = U.write_png(a, filters=[4] * 40, level=0)                     # stored blocks only
    flat = np.full((40, 33, 3), 200, np.uint8)                         # long matches at distance 1 and 3
    f1 = U.write_png(flat, filters=[1] * 40, level=9)
    got = lj.decode_png([f0, f1]).cpu().numpy()
    assert np.array_equal(got[0], a) and np.array_equal(got[1], flat)


def test_image_larger_than_the_window_pil_files(lj, tmp_path):
    """512 x 512 RGB (768 KB of scanlines: the 32 KB ring wraps 24 times), written by PIL with its adaptive filters
    at several compression levels, photo-like and noise content; several files in one call."""
    from PIL import Image
    import importlib
    synth = importlib.import_module("litejpeg_b200.synth")
    imgs = [synth.frame(512, 512, 9, 0).numpy(), synth.frame(512, 512, 9, 1, kind="noise").numpy(), photo(512, 512, 3, 5)]
    files = []
    for i, a in enumerate(imgs):
        for lvl in (1, 6, 9):
            buf = io.BytesIO()
            Image.fromarray(a).save(buf, format="PNG", compress_level=lvl, optimize=(lvl == 9))
            files.append(buf.getvalue())
    got = lj.decode_png(files).cpu().numpy()
    for i, f in enumerate(files):
        assert np.array_equal(got[i], imgs[i // 3]), i
    p = tmp_path / "x.png"
    p.write_bytes(files[0])
    assert np.array_equal(lj.read_image(str(p), decoder="device").cpu().numpy(), imgs[0])
    assert np.array_equal(lj.read_image(str(p), decoder="pil").cpu().numpy(), imgs[0])


def test_odd_sizes(lj):
    for (h, w, c) in [(1, 1, 3), (1, 300, 4), (300, 1, 1), (33, 31, 3), (32, 64, 4), (65, 2, 3)]:
        a = photo(h, w, c, h * 7 + w)
        f = U.write_png(a, filters=[(3 * r + 1) % 5 for r in range(h)])
        assert np.array_equal(lj.decode_png([f]).cpu().numpy()[0], pil_rgb(f)), (h, w, c)


def test_damaged_streams_are_reported(lj):
    a = photo(50, 40, 3, 3)
    good = U.write_png(a, filters=[4] * 50)
    raw = U.filter_rows(a, [4] * 50)
    short = U.write_png(a, raw_override=raw[:-100])                    # fewer bytes than the image holds
    long_ = U.write_png(a, raw_override=raw + b"\0" * 50)              # more
    badft = U.write_png(a, raw_override=bytes([9]) + raw[1:])          # filter type 9
    for bad, what in ((short, "less data"), (long_, "more data"), (badft, "filter")):
        with pytest.raises(ValueError, match=what):
            lj.decode_png([good, bad])
    # garbage in the middle of the zlib stream: any verdict but "decoded" (or, if the damage happens to decode, other pixels)
    z = bytearray(good)
    i0 = good.index(b"IDAT") + 4 + 40
    z[i0:i0 + 8] = b"\xff" * 8
    try:
        got = lj.decode_png([bytes(z)]).cpu().numpy()[0]
        assert not np.array_equal(got, a)
    except ValueError:
        pass
    assert np.array_equal(lj.decode_png([good]).cpu().numpy()[0], a)  # and the library is still healthy


def test_other_png_flavours_go_to_pil(lj, tmp_path):
    from PIL import Image
    a = photo(20, 24, 3, 2)
    pal = Image.fromarray(a).convert("P")
    p = tmp_path / "p.png"
    pal.save(str(p))
    with pytest.raises(RuntimeError):
        lj.png_info(p.read_bytes())
    with pytest.raises(RuntimeError):
        lj.read_image(str(p), decoder="device")
    assert np.array_equal(lj.read_image(str(p)).cpu().numpy(), np.array(pal.convert("RGB")))
    inter = U.write_png(a, interlace=1)
    with pytest.raises(RuntimeError):
        lj.decode_png([inter])
    with pytest.raises(RuntimeError):
        lj.decode_png([U.write_png(a), U.write_png(photo(21, 24, 3, 2))])   # one size per call


def test_png_to_jpeg_on_the_device(lj):
    """file bytes -> decode_png -> JpegEncoder: the whole path of examples/encode_image.py without a host decode."""
    from oracle import jpeg_ext as J
    a = photo(72, 88, 3, 11)
    f = U.write_png(a, filters=[r % 5 for r in range(72)])
    frames = lj.decode_png([f])
    out = lj.JpegEncoder("420", quality=75).encode(frames)
    assert out[0] == J.encode_jpeg(a, "420", 75, "eager")


def test_batch_transcode_example(lj, tmp_path):
    """examples/transcode_png_batch.py: a directory of PNG files in, JPEG files out, equal to the oracle's files."""
    import subprocess
    from oracle import jpeg_ext as J
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    imgs = [photo(48, 80, 3, 20 + i) for i in range(5)]
    for i, a in enumerate(imgs):
        (src / ("f%02d.png" % i)).write_bytes(U.write_png(a, filters=[(r + i) % 5 for r in range(48)]))
    run = subprocess.run([sys.executable, os.path.join(root, "examples", "transcode_png_batch.py"), str(src), str(dst), "--sub", "422",
                          "--quality", "80", "--batch", "3"], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    for i, a in enumerate(imgs):
        assert (dst / ("f%02d.jpg" % i)).read_bytes() == J.encode_jpeg(a, "422", 80, "standard"), i
