#!/usr/bin/env python
"""Wider seeded sweep than tests/test_gpu_encode.py::test_random_parameter_space_bytes_equal_oracle: shape x subsampling x
quality x ZRL policy x restart interval x DCT mode x content x layout, every file compared with the C oracle's bytes.
usage: fuzz_encode.py [seed] [cases]   (oracle/_build/libb2j_oracle.so must exist: make -C oracle)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import litejpeg_b200 as lj                      # noqa: E402
from litejpeg_b200 import synth                 # noqa: E402
from conftest import OracleC                    # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 300
oracle = OracleC(os.path.join(ROOT, "oracle", "_build", "libb2j_oracle.so"))
rng = np.random.default_rng(seed)
SUBS = ("444", "422", "420")
bad = 0
nbytes = 0
for case in range(cases):
    big = case % 10 == 9
    h = int(rng.integers(1, 1200 if big else 300))
    w = int(rng.integers(1, 2000 if big else 700))
    sub = SUBS[int(rng.integers(0, 3))]
    q = int(rng.choice([1, 5, 25, 50, 75, 90, 95, 100, int(rng.integers(1, 101))]))
    zi = int(rng.integers(0, 2))
    ri = int(rng.choice([0, 0, 0, 1, 2, 7, 64, 999]))
    fixed = bool(rng.integers(0, 6) == 0)
    n = int(rng.integers(1, 5))
    kind = ("photo", "noise", "grey", "extreme", "photo")[int(rng.integers(0, 5))]
    if kind == "extreme":                           # black / white checker of random period: the largest amplitudes, many 0xFF bytes
        p = int(rng.integers(1, 9))
        yy, xx = np.mgrid[0:h, 0:w]
        a = ((((yy // p) + (xx // p)) & 1) * 255).astype(np.uint8)
        frames = torch.from_numpy(np.repeat(a[None, :, :, None], 3, axis=3).repeat(n, axis=0).copy())
    else:
        frames = torch.stack([synth.frame(h, w, 100 + case, i, kind) for i in range(n)])
        if kind == "grey":
            frames[..., :] = int(rng.integers(0, 256))
    cap = ((h + 15) // 16 * 16) * ((w + 15) // 16 * 16) * 6 + 4096
    enc = lj.JpegEncoder(sub, quality=q, zrl=("eager", "standard")[zi], restart_interval=ri, dct="fixed" if fixed else "model", out_cap=cap)
    files = enc.encode(frames.cuda())
    for i in range(n):
        want = oracle.encode_jpeg(frames[i].numpy(), sub, q, zi | (0x100 if fixed else 0) | (ri << 16))
        nbytes += len(want)
        if files[i] != want:
            bad += 1
            print("MISMATCH", dict(case=case, h=h, w=w, sub=sub, q=q, zrl=zi, ri=ri, fixed=fixed, kind=kind, frame=i, got=len(files[i]), want=len(want)), flush=True)
print("seed %d: %d cases, %.1f MB of JPEG compared, %d mismatches" % (seed, cases, nbytes / 1e6, bad))
sys.exit(1 if bad else 0)
