#!/usr/bin/env python
"""Encode an image file to baseline JPEG on the GPU with the model-exact path.

    python examples/encode_image.py in.png out.jpg [--sub 420] [--quality 85] [--zrl standard] [--restart 0] [--dct model]

quality 85 + sub 444 is exactly the configuration of the reference's software model
(test/model/enc_frame.py: its tables are Annex-K at IJG quality 85); `--zrl eager` reproduces the
gateware's run-length rule (rlecore.py:101-105) instead of the T.81 one."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import litejpeg_b200 as lj          # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--sub", default="420", choices=["444", "422", "420"])
    ap.add_argument("--quality", type=int, default=85)
    ap.add_argument("--zrl", default="standard", choices=["eager", "standard"])
    ap.add_argument("--restart", type=int, default=0, help="restart interval in MCUs (0 = none)")
    ap.add_argument("--dct", default="model", choices=["model", "fixed"])
    a = ap.parse_args()
    rgb = lj.read_image(a.src)              # 8-bit grey / RGB / RGBA PNG: decoded on the GPU; other files: PIL on the host
    enc = lj.JpegEncoder(a.sub, quality=a.quality, zrl=a.zrl, restart_interval=a.restart, dct=a.dct,
                         out_cap=rgb.shape[0] * rgb.shape[1] * 4 + 65536)
    jpg = enc.encode(rgb[None])[0]
    with open(a.dst, "wb") as f:
        f.write(jpg)
    print("%s: %d x %d -> %d bytes (%.2f bits per pixel)" % (a.dst, rgb.shape[1], rgb.shape[0], len(jpg),
                                                            8.0 * len(jpg) / (rgb.shape[0] * rgb.shape[1])))


if __name__ == "__main__":
    main()
