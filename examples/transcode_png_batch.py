#!/usr/bin/env python
"""A directory of equally sized PNG files -> baseline JPEG files, without a host pass over the pixels:
the files' zlib streams are inflated and unfiltered on the GPU (lj.decode_png), the frames go straight
into the encode path (lj.JpegEncoder), only file bytes cross the PCIe bus in either direction.

    python examples/transcode_png_batch.py in_dir out_dir [--sub 420] [--quality 75] [--batch 256]

(8-bit grey / RGB / RGBA, non-interlaced PNG files of one size; see examples/encode_image.py for anything else)"""
import argparse
import glob
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import litejpeg_b200 as lj          # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src_dir")
    ap.add_argument("dst_dir")
    ap.add_argument("--sub", default="420", choices=["444", "422", "420"])
    ap.add_argument("--quality", type=int, default=75)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    paths = sorted(glob.glob(os.path.join(a.src_dir, "*.png")))
    if not paths:
        sys.exit("no *.png files in " + a.src_dir)
    os.makedirs(a.dst_dir, exist_ok=True)
    enc = lj.JpegEncoder(a.sub, quality=a.quality, zrl="standard")
    t0 = time.perf_counter()
    n_in = n_out = 0
    for i in range(0, len(paths), a.batch):
        chunk = paths[i:i + a.batch]
        blobs = [open(p, "rb").read() for p in chunk]
        frames = lj.decode_png(blobs)                              # (N, H, W, 3) uint8 on the GPU
        for p, jpg in zip(chunk, enc.encode(frames)):
            with open(os.path.join(a.dst_dir, os.path.splitext(os.path.basename(p))[0] + ".jpg"), "wb") as f:
                f.write(jpg)
            n_out += len(jpg)
        n_in += sum(len(b) for b in blobs)
    dt = time.perf_counter() - t0
    print("%d files, %.1f MB of PNG -> %.1f MB of JPEG in %.2f s" % (len(paths), n_in / 1e6, n_out / 1e6, dt))


if __name__ == "__main__":
    main()
