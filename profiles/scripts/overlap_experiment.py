#!/usr/bin/env python
"""Does running K1 (FP64-heavy) and the entropy kernel (ALU-heavy) of different sub-batches
concurrently on one GPU beat running them back to back?  Both are issue-bound at ~68 % issue
utilisation with complementary pipes, so co-resident CTAs could fill each other's stall slots.

usage: overlap_experiment.py [n_frames] [sub_batches ...]
Env knobs B2J_K1_CTAS_PER_SM / B2J_ENT_CTAS_PER_SM cap the persistent grids (read per launch).
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402
from litejpeg_b200 import _lib, synth           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SUBS = [int(v) for v in sys.argv[2:]] or [1, 2, 4, 8]
H, W, SUB, Q = 1080, 1920, "420", 75
L = _lib.lib()
dev = torch.device("cuda:0")
frames = synth.batch(N, H, W, 2, unique=32).to(dev)
qy, qc = _lib.quant_tables(Q)
tqy, tqc = _lib.table_arg(qy), _lib.table_arg(qc)
d1 = _lib.make_desc(H, W, SUB, 1)
g = _lib.geometry(d1)
cap = int(L.b2j_default_out_cap(ctypes.byref(d1)))
coefs = torch.empty((N, g.blocks_per_frame, 64), dtype=torch.int16, device=dev)
out = torch.empty((N, cap), dtype=torch.uint8, device=dev)
sizes = torch.empty((N,), dtype=torch.int64, device=dev)


def run(n_sub, k1_ctas, ent_ctas, steps=5, warm=3):
    if k1_ctas:
        os.environ["B2J_K1_CTAS_PER_SM"] = str(k1_ctas)
    else:
        os.environ.pop("B2J_K1_CTAS_PER_SM", None)
    if ent_ctas:
        os.environ["B2J_ENT_CTAS_PER_SM"] = str(ent_ctas)
    else:
        os.environ.pop("B2J_ENT_CTAS_PER_SM", None)
    per = N // n_sub
    dsub = _lib.make_desc(H, W, SUB, per)
    wsb = int(L.b2j_entropy_workspace_bytes(ctypes.byref(dsub), cap))
    wss = [torch.empty((wsb,), dtype=torch.uint8, device=dev) for _ in range(n_sub)]
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    evs = [torch.cuda.Event() for _ in range(n_sub)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        for i in range(n_sub):
            lo = i * per
            _lib.check(L.b2j_fused_block(frames[lo:lo + per].data_ptr(), ctypes.byref(dsub), tqy, tqc,
                                         coefs[lo:lo + per].data_ptr(), 0, sa.cuda_stream))
            evs[i].record(sa)
            sb.wait_event(evs[i])
            _lib.check(L.b2j_entropy_encode(coefs[lo:lo + per].data_ptr(), ctypes.byref(dsub), tqy, tqc, 0,
                                            wss[i].data_ptr(), out[lo:lo + per].data_ptr(), cap,
                                            sizes[lo:lo + per].data_ptr(), sb.cuda_stream))
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    t0.record(cur)
    sa.wait_event(t0)
    sb.wait_event(t0)
    for _ in range(steps):
        step()
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(sa)
    e2.record(sb)
    cur.wait_event(e1)
    cur.wait_event(e2)
    t1.record(cur)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    return ms, int(sizes.sum())


ref = None
for n_sub in SUBS:
    for k1c, entc in ((0, 0), (2, 0), (2, 3), (2, 4), (3, 2), (1, 0), (1, 6)):
        if n_sub == 1 and (k1c, entc) != (0, 0):
            continue
        ms, total = run(n_sub, k1c, entc)
        ref = ref or total
        print("sub_batches %2d  K1 ctas/SM %s  entropy ctas/SM %s  -> %.3f ms/step  (%.0f Mpixel/s)%s"
              % (n_sub, k1c or "max", entc or "max", ms, N * H * W / ms / 1e3, "" if total == ref else "  SIZE MISMATCH"), flush=True)
