#!/usr/bin/env python
"""K1 (b2j_fused_block) alone, for several builds of the library on one box.
usage: time_k1.py lib_a.so lib_b.so ...   (256 x 1080p 4:2:0 q75; CUDA events, 10 launches after 3 warm-up)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402,F401
from litejpeg_b200 import _lib, synth           # noqa: E402

N, H, W = 256, 1080, 1920
dev = torch.device("cuda:0")
frames = synth.batch(N, H, W, 2, unique=32).to(dev)
desc = _lib.make_desc(H, W, "420", N)
coefs = torch.zeros((N * 48960, 64), dtype=torch.int16, device=dev)
P = ctypes.c_void_p
for rnd in range(2):
    for path in sys.argv[1:]:
        L = ctypes.CDLL(os.path.abspath(path))
        L.b2j_quant_tables.argtypes = [ctypes.c_int, P, P]
        L.b2j_fused_block.argtypes = [P, P, P, P, P, ctypes.c_int, P]
        qy = (ctypes.c_uint8 * 64)()
        qc = (ctypes.c_uint8 * 64)()
        assert L.b2j_quant_tables(75, qy, qc) == 0
        sp = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            assert L.b2j_fused_block(frames.data_ptr(), ctypes.byref(desc), qy, qc, coefs.data_ptr(), 0, sp) == 0
        torch.cuda.synchronize()
        e0.record()
        for it in range(10):
            assert L.b2j_fused_block(frames.data_ptr(), ctypes.byref(desc), qy, qc, coefs.data_ptr(), 0, sp) == 0
        e1.record()
        torch.cuda.synchronize()
        print("round %d  %-28s K1 %.4f ms" % (rnd, os.path.basename(path), e0.elapsed_time(e1) / 10), flush=True)
