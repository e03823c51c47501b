#!/usr/bin/env python
"""One 8K frame across the ranks of a torchrun job: every rank encodes its band of
whole MCU rows, rank 0 joins the bands and checks the file against its own whole-frame encode with
one restart interval per MCU row.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/scripts/split_frame_demo.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import litejpeg_b200 as lj                      # noqa: E402
from litejpeg_b200 import synth                 # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, SUB = 4320, 7680, "420"
frame = synth.frame(H, W, 4, 0)
enc = lj.ShardedJpegEncoder(world, rank, subsampling=SUB, quality=75)
plan = lj.band_plan(H, SUB, world)
y0, y1, before = plan[rank]
band = frame[y0:y1].cuda()
out = enc.encode_frame_split(band.new_empty(0) if False else frame.cuda())          # warm-up (tables, workspaces)
# device time of this rank's band alone
be = enc.encoder.band_encoder(W, before)
be.encode_device(band[None])
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
dist.barrier()
torch.cuda.synchronize()
t0.record()
for _ in range(10):
    be.encode_device(band[None])
t1.record()
torch.cuda.synchronize()
ms = torch.tensor([t0.elapsed_time(t1) / 10], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
w0 = time.perf_counter()
out = enc.encode_frame_split(frame.cuda())
wall = time.perf_counter() - w0
if rank == 0:
    mcus_x = lj.geometry(lj.make_desc(H, W, SUB, 1)).mcus_x
    whole_enc = lj.JpegEncoder(SUB, quality=75, restart_interval=mcus_x)
    whole = whole_enc.encode(frame.cuda()[None])[0]
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    dev = frame.cuda()[None]
    whole_enc.encode_device(dev)
    e0.record()
    for _ in range(10):
        whole_enc.encode_device(dev)
    e1.record()
    torch.cuda.synchronize()
    print("ranks %d  bands %s" % (world, [(a, b) for a, b, _ in plan]))
    print("joined file == whole-frame encode on one GPU: %s (%d bytes)" % (out == whole, len(out)))
    print("device time: slowest band %.3f ms  vs whole frame on one GPU %.3f ms  -> %.2fx" % (ms.item(), e0.elapsed_time(e1) / 10,
                                                                                            e0.elapsed_time(e1) / 10 / ms.item()))
    print("end to end incl. H2D of the frame, D2H of the bands, gather_object and the join: %.1f ms wall" % (wall * 1e3))
dist.barrier()
dist.destroy_process_group()
