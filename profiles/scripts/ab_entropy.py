#!/usr/bin/env python
"""A/B timing of b2j_fused_block + b2j_entropy_encode for several builds of the library on one box.
usage: ab_entropy.py [--n=256 --h=1080 --w=1920 --sub=420 --q=75 --kind=photo --unique=32] lib_a.so lib_b.so ...
(default 256 x 1080p 4:2:0 q75, eager; CUDA events, 10 steps after 3 warm-up)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402  (frame generator, descriptor struct)
from litejpeg_b200 import _lib, synth           # noqa: E402

opt = {"n": "256", "h": "1080", "w": "1920", "sub": "420", "q": "75", "kind": "photo", "unique": "32"}
libs = [a for a in sys.argv[1:] if not a.startswith("--")]
for a in sys.argv[1:]:
    if a.startswith("--"):
        k, v = a[2:].split("=")
        opt[k] = v
N, H, W, SUB, Q = int(opt["n"]), int(opt["h"]), int(opt["w"]), opt["sub"], int(opt["q"])
dev = torch.device("cuda:0")
frames = synth.batch(N, H, W, 2, kind=opt["kind"], unique=min(N, int(opt["unique"]))).to(dev)
desc = _lib.make_desc(H, W, SUB, N)
hs, vs = {"444": (1, 1), "422": (2, 1), "420": (2, 2)}[SUB]
BLOCKS = ((W + 8 * hs - 1) // (8 * hs)) * ((H + 8 * vs - 1) // (8 * vs)) * (hs * vs + 2)
P = ctypes.c_void_p
for rnd in range(2):
    for path in libs:
        L = ctypes.CDLL(os.path.abspath(path))
        L.b2j_default_out_cap.restype = ctypes.c_int64
        L.b2j_entropy_workspace_bytes.restype = ctypes.c_size_t
        L.b2j_entropy_workspace_bytes.argtypes = [P, ctypes.c_int64]
        L.b2j_quant_tables.argtypes = [ctypes.c_int, P, P]
        L.b2j_fused_block.argtypes = [P, P, P, P, P, ctypes.c_int, P]
        L.b2j_entropy_encode.argtypes = [P, P, P, P, ctypes.c_int, P, P, ctypes.c_int64, P, P]
        qy = (ctypes.c_uint8 * 64)()
        qc = (ctypes.c_uint8 * 64)()
        assert L.b2j_quant_tables(Q, qy, qc) == 0
        cap = int(L.b2j_default_out_cap(ctypes.byref(desc)))
        coefs = torch.empty((N * BLOCKS, 64), dtype=torch.int16, device=dev)
        ws = torch.empty((L.b2j_entropy_workspace_bytes(ctypes.byref(desc), cap) + 256,), dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) & ~255
        out = torch.empty((N, cap), dtype=torch.uint8, device=dev)
        sizes = torch.empty((N,), dtype=torch.int64, device=dev)
        sp = torch.cuda.current_stream().cuda_stream
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t1 = t2 = 0.0
        iters = 13 if N > 8 else 203
        for it in range(iters):
            ev[0].record()
            assert L.b2j_fused_block(frames.data_ptr(), ctypes.byref(desc), qy, qc, coefs.data_ptr(), 0, sp) == 0
            ev[1].record()
            assert L.b2j_entropy_encode(coefs.data_ptr(), ctypes.byref(desc), qy, qc, 0, wp, out.data_ptr(), cap, sizes.data_ptr(), sp) == 0
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 3:
                t1 += ev[0].elapsed_time(ev[1])
                t2 += ev[1].elapsed_time(ev[2])
        print("round %d  %-40s K1 %.4f ms  entropy+stuffing %.4f ms  bytes %d" % (rnd, os.path.basename(path), t1 / (iters - 3), t2 / (iters - 3), int(sizes.sum())), flush=True)
        del coefs, ws, out, sizes
