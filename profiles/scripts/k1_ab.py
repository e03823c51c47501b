#!/usr/bin/env python
"""A/B of K1 (b2j_fused_block) builds/variants on one box: bit-compare + timing.
usage: k1_ab.py [--sub 420] [--n 256] [--hw 1080x1920] [--kind photo] spec ...   spec = path.so[:variant]  (variant 0 auto, 2 tiled, 3 stream)
The first spec is the reference all others are compared with (torch.equal on the whole coefficient array)."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402,F401
from litejpeg_b200 import _lib, synth           # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sub", default="420")
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--hw", default="1080x1920")
ap.add_argument("--kind", default="photo")
ap.add_argument("--q", type=int, default=75)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("specs", nargs="+")
a = ap.parse_args()
H, W = map(int, a.hw.split("x"))
N = a.n
dev = torch.device("cuda:0")
frames = synth.batch(N, H, W, 2, a.kind, unique=min(N, 16)).to(dev)
desc = _lib.make_desc(H, W, a.sub, N)
geo = lj.geometry(desc)
nblk = N * geo.blocks_per_frame
P = ctypes.c_void_p
ref = None
px_padded = N * geo.padded_width * geo.padded_height
bpp = {"444": 9, "422": 7, "420": 6}[a.sub]
for rnd in range(2):
    for spec in a.specs:
        path, _, var = spec.partition(":")
        var = int(var or 0)
        L = ctypes.CDLL(os.path.abspath(path))
        L.b2j_quant_tables.argtypes = [ctypes.c_int, P, P]
        L.b2j_fused_block.argtypes = [P, P, P, P, P, ctypes.c_int, P]
        qy = (ctypes.c_uint8 * 64)()
        qc = (ctypes.c_uint8 * 64)()
        assert L.b2j_quant_tables(a.q, qy, qc) == 0
        coefs = torch.full((nblk, 64), 0x5A5A, dtype=torch.int16, device=dev)
        sp = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            rc = L.b2j_fused_block(frames.data_ptr(), ctypes.byref(desc), qy, qc, coefs.data_ptr(), var, sp)
            assert rc == 0, rc
        torch.cuda.synchronize()
        e0.record()
        for it in range(a.iters):
            L.b2j_fused_block(frames.data_ptr(), ctypes.byref(desc), qy, qc, coefs.data_ptr(), var, sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        if ref is None:
            ref = coefs.clone()
            same = "ref"
        else:
            same = "same" if torch.equal(ref, coefs) else "MISMATCH (%d blocks)" % int((ref != coefs).any(dim=1).sum())
        print("round %d  %-44s K1 %.4f ms  %.1f%% of 6533.5 GB/s  %s" %
              (rnd, os.path.basename(path) + ":" + str(var), ms, 100 * px_padded * bpp / (ms * 1e-3) / 6533.5e9, same), flush=True)
        del coefs
