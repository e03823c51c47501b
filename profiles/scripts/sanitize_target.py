#!/usr/bin/env python
"""What compute-sanitizer runs (profiles/scripts/sanitize.sh): smoke(), then the cases the inter-CTA
flag passing and the lookahead tail of the entropy kernel are most exposed to, each checked against the oracle:
  * restart intervals (many short streams, SEG instantiation), 4:2:0 / 4:2:2 / 4:4:4;
  * flat frames (4-6 bit blocks: the stream's final word is composed by a thread of the previous chunk,
    through the 8 lookahead columns);
  * several chunks per stream with a look-back over more than one predecessor (a 1080p frame);
  * every K1 variant (simple / tiled / stream) on a ragged geometry;
  * the host-buffer encoders (copies on two streams, worker threads)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__                     # noqa: E402
import litejpeg_b200 as lj                # noqa: E402
from oracle import jpeg_ext as J          # noqa: E402
import importlib                          # noqa: E402

synth = importlib.import_module("litejpeg_b200.synth")
__graft_entry__.smoke()
torch.cuda.set_device(0)

frames = torch.stack([synth.frame(136, 200, 5, i) for i in range(3)])
for sub, ri in (("420", 5), ("422", 13), ("444", 25)):
    enc = lj.JpegEncoder(sub, quality=75, restart_interval=ri)
    files = enc.encode(frames.cuda())
    for i, f in enumerate(files):
        assert f == J.encode_jpeg(frames[i].numpy(), sub, 75, "eager", restart_interval=ri), ("restart", sub, i)
print("restart intervals ok")

flat = torch.stack([torch.full((264, 520, 3), v, dtype=torch.uint8) for v in (128, 0, 255)])
for sub in ("420", "444"):
    files = lj.JpegEncoder(sub, quality=75).encode(flat.cuda())
    for i, f in enumerate(files):
        assert f == J.encode_jpeg(flat[i].numpy(), sub, 75, "eager"), ("flat", sub, i)
print("flat tails ok")

big = synth.frame(1080, 1920, 6, 0)[None]
f = lj.JpegEncoder("420", quality=75).encode(big.cuda())[0]
assert f == J.encode_jpeg(big[0].numpy(), "420", 75, "eager")
print("1080p look-back ok")

rag = torch.stack([synth.frame(77, 301, 7, i, "noise") for i in range(2)])
for sub in ("420", "444"):
    enc = lj.JpegEncoder(sub, quality=75)
    exp = np.stack([J.encode_coefficients(rag[i].numpy(), sub, quality=75)[0].reshape(-1, 64) for i in range(2)])
    for variant in ("simple", "tiled", "stream"):
        assert np.array_equal(enc.coefficients(rag.cuda(), variant=variant).cpu().numpy(), exp), (sub, variant)
print("K1 variants ok")

enc = lj.JpegEncoder("420", quality=75)
out, sizes = enc.encode_host(frames.pin_memory(), frames_per_chunk=2)
menc = lj.MultiGpuJpegEncoder([0, 0], 136, 200, "420", quality=75, frames_per_chunk=1)
out2, sizes2 = menc.encode_host(frames.pin_memory())
for i in range(3):
    ref = J.encode_jpeg(frames[i].numpy(), "420", 75, "eager")
    assert out[i, :int(sizes[i])].numpy().tobytes() == ref and out2[i, :int(sizes2[i])].numpy().tobytes() == ref
enc.close()
menc.close()
print("host encoders ok")
