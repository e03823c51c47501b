#!/usr/bin/env python
"""[ext] PNG ingest on the device: time of lj.decode_png for 1 / 64 / 1024 photographic 512 x 512 RGB files (PIL-written,
compress_level 6) against PIL decoding the same files on one host core.  usage: png_bench.py"""
import io
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402
from litejpeg_b200 import synth                 # noqa: E402

uniq = []
for i in range(8):
    buf = io.BytesIO()
    Image.fromarray(synth.frame(512, 512, 9, i).numpy()).save(buf, format="PNG", compress_level=6)
    uniq.append(buf.getvalue())
print("file size %.0f KB (768 KB of pixels)" % (len(uniq[0]) / 1e3))
t0 = time.perf_counter()
for f in uniq:
    np.array(Image.open(io.BytesIO(f)))
pil = (time.perf_counter() - t0) / len(uniq)
print("PIL on one host core: %.2f ms per file" % (pil * 1e3))
for n in (1, 64, 1024):
    files = [uniq[i % 8] for i in range(n)]
    lj.decode_png(files)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = lj.decode_png(files)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ref = np.array(Image.open(io.BytesIO(files[-1])))
    assert np.array_equal(out[-1].cpu().numpy(), ref)
    print("decode_png %4d files: %8.2f ms  = %7.3f ms per file  (%.1f x one host core; %.0f Mpixel/s)" % (n, dt * 1e3, dt * 1e3 / n, pil * n / dt, n * 0.262144 / dt))
