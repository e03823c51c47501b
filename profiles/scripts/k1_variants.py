#!/usr/bin/env python
"""Times b2j_fused_block (K1) of several builds of libb2j.so on the same device buffers.
usage: python profiles/scripts/k1_variants.py lib1.so lib2.so ...   (256 x 1080p, 4:2:0, q75)"""
import ctypes
import importlib
import sys

import torch

sys.path.insert(0, ".")
import litejpeg_b200 as lj

synth = importlib.import_module("litejpeg_b200.synth")
H, W, N = 1080, 1920, 256
host = synth.batch(N, H, W, 2, "photo", unique=16)
rgb = host.cuda()
desc = lj.make_desc(H, W, "420", N)
geo = lj.geometry(desc)
qy, qc = lj.quant_tables(75)
qa, qb = lj._lib.table_arg(qy), lj._lib.table_arg(qc)
coefs = torch.empty((N * geo.blocks_per_frame, 64), dtype=torch.int16, device="cuda")
ref = None
st = torch.cuda.current_stream().cuda_stream
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    fn = lib.b2j_fused_block
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(lj._lib.ImageDesc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int, ctypes.c_void_p]
    for _ in range(3):
        assert fn(rgb.data_ptr(), ctypes.byref(desc), qa, qb, coefs.data_ptr(), 0, st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn(rgb.data_ptr(), ctypes.byref(desc), qa, qb, coefs.data_ptr(), 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    chk = int(coefs[::997].to(torch.int64).sum().item())
    if ref is None:
        ref = chk
    print("%-70s %.3f ms  %.1f%% of HBM peak  checksum %s" % (path.split("/")[-1], ms, 100 * N * 1920 * 1088 * 6 / (ms * 1e-3) / 6533.5e9,
                                                              "ok" if chk == ref else "MISMATCH"))
