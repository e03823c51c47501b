"""Generates the committed golden fixtures from the REAL reference model.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports /root/reference/test/model/enc_frame.py (and the older encoder.py,
reference.py and the fixture classes of test/common.py behind an empty `litex`
stub) unmodified, runs them, and stores inputs + outputs as .npz files next to
this script.  The tests then check oracle/ and the CUDA path against these
files without touching /root/reference.
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "test"))
sys.path.insert(0, os.path.join(REF, "test", "model"))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from PIL import Image  # noqa: E402

if not hasattr(Image, "ANTIALIAS"):          # removed in Pillow >= 10; test/common.py:50 uses it
    Image.ANTIALIAS = Image.LANCZOS

# empty litex stub so test/common.py imports (it only uses litex for `from ... import *`)
for name in ("litex", "litex.gen", "litex.soc", "litex.soc.interconnect",
             "litex.soc.interconnect.stream"):
    sys.modules.setdefault(name, types.ModuleType(name))

from model import enc_frame as ef  # noqa: E402
import common as refcommon  # noqa: E402  (test/common.py)
import reference as refvec  # noqa: E402  (test/model/reference.py)

from oracle import model_np as M  # noqa: E402

# litejpeg/core/csc.py cannot be imported (it needs LiteX), but its coefficient helper is plain
# Python: execute just that function (and `coef` from core/common.py) from the reference source.
_ns = {}
_src = open(os.path.join(REF, "litejpeg", "core", "common.py")).read()
exec(_src[_src.index("def coef("):_src.index("def rgb_layout(")], _ns)
_src = open(os.path.join(REF, "litejpeg", "core", "csc.py")).read()
exec(_src[_src.index("def ycbcr2rgb_coefs("):_src.index("@CEInserter()", _src.index("def ycbcr2rgb_coefs("))], _ns)
exec(_src[_src.index("def rgb2ycbcr_coefs("):_src.index("datapath_latency")], _ns)
_mod = types.ModuleType("litejpeg_csc_coefs")
_mod.ycbcr2rgb_coefs = _ns["ycbcr2rgb_coefs"]
_mod.rgb2ycbcr_coefs = _ns["rgb2ycbcr_coefs"]
sys.modules["litejpeg_csc_coefs"] = _mod

rng = np.random.default_rng(20171019)
meta = {}


def model_block_pipeline(r, g, b):
    """enc_frame.py:203-217 for one block."""
    y, cb, cr = ef.rgb2ycbcr(r, g, b)
    out = []
    for blk, tab in ((y, ef.y_quant_table), (cb, ef.cbcr_quant_table), (cr, ef.cbcr_quant_table)):
        out.append(ef.zigzag(ef.quantize(ef.dct(blk), tab)))
    return (y, cb, cr), out


# ---- 1. lena.png, every block, through the real model (config #1) -------------
lena = np.array(Image.open(os.path.join(REF, "test", "lena.png")))[..., :3]
rb, gb, bb = ef.readRgbImageBlocks(os.path.join(REF, "test", "lena.png"))
n = len(rb)
ycc = np.zeros((3, n, 64), dtype=np.uint8)
zz = np.zeros((3, n, 64), dtype=np.int16)
for i in range(n):
    planes, outs = model_block_pipeline(rb[i], gb[i], bb[i])
    for c in range(3):
        ycc[c, i] = planes[c]
        zz[c, i] = outs[c]
# DC differencing + rle_code exactly as enc_frame.py:219-234, first 64 blocks
prev = [0, 0, 0]
rle_first = []
for i in range(64):
    for c in range(3):
        blk = [int(x) for x in zz[c, i]]
        d = blk[0] - prev[c]
        prev[c] = blk[0]
        blk[0] = d
        rle_first.append(ef.rle_code(blk))
# (the colour planes are not stored: the colour stage is pinned exhaustively in section 3)
np.savez_compressed(os.path.join(HERE, "lena_model.npz"), rgb=lena, zigzag=zz)
with open(os.path.join(HERE, "lena_rle_first64.json"), "w") as f:
    json.dump(rle_first, f)
meta["lena_blocks"] = n

# ---- 2. per-stage vectors from the real model ----------------------------------
# colour: random + corners
px = rng.integers(0, 256, size=(8192, 3), dtype=np.int64)
corners = np.array([[a, b, c] for a in (0, 1, 127, 128, 254, 255) for b in (0, 1, 127, 128, 254, 255)
                    for c in (0, 1, 127, 128, 254, 255)], dtype=np.int64)
px = np.concatenate([px, corners])
col = np.zeros((len(px), 3), dtype=np.int64)
for s in range(0, len(px), 64):
    chunk = px[s:s + 64]
    y, cb, cr = ef.rgb2ycbcr(list(chunk[:, 0]), list(chunk[:, 1]), list(chunk[:, 2]))
    col[s:s + len(chunk), 0] = y[:len(chunk)]
    col[s:s + len(chunk), 1] = cb[:len(chunk)]
    col[s:s + len(chunk), 2] = cr[:len(chunk)]

# dct: random, smooth, extreme
blocks = [rng.integers(0, 256, 64) for _ in range(192)]
blocks += [np.clip(rng.integers(0, 256) + rng.integers(-6, 7, 64), 0, 255) for _ in range(192)]
blocks += [np.zeros(64, int), np.full(64, 255), np.full(64, 128), np.arange(64) * 4,
           np.array([255 if (i // 8 + i % 8) % 2 else 0 for i in range(64)]),
           np.array([255 if i % 2 else 0 for i in range(64)]),
           np.array([255 if (i // 8) % 2 else 0 for i in range(64)])]
blocks = np.array(blocks, dtype=np.int64)
dct_out = np.array([ef.dct([int(v) for v in b]) for b in blocks], dtype=np.int64)

# quantize: model on dct outputs with both tables; plus exhaustive check of the integer restatement
q_y = np.array([ef.quantize([int(v) for v in b], ef.y_quant_table) for b in dct_out], dtype=np.int64)
q_c = np.array([ef.quantize([int(v) for v in b], ef.cbcr_quant_table) for b in dct_out], dtype=np.int64)
bad = 0
for d in range(1, 256):
    vals = list(range(-2048, 2048))
    for s in range(0, 4096, 64):
        got = ef.quantize(vals[s:s + 64], [d] * 64)
        exp = M.quantize(np.array(vals[s:s + 64]), np.full(64, d))
        bad += int((np.array(got) != exp).sum())
meta["quantize_integer_restatement_mismatches_d1_255_v_pm2048"] = bad
assert bad == 0
zz_out = np.array([ef.zigzag([int(v) for v in b]) for b in q_y], dtype=np.int64)
np.savez_compressed(os.path.join(HERE, "stage_vectors.npz"), rgb=px, ycbcr=col, dct_in=blocks,
                    dct_out=dct_out, quant_y=q_y, quant_c=q_c, zigzag_out=zz_out,
                    y_quant_table=np.array(ef.y_quant_table), cbcr_quant_table=np.array(ef.cbcr_quant_table),
                    zig_zag_order=np.array(ef.zig_zag_order))

# ---- 3. exhaustive colour conversion: all 2^24 inputs ---------------------------
# The NumPy restatement is first compared with the scalar model on 200k random
# inputs (+ every input whose exact value is an integer for Y, the only place
# float rounding can matter), then its full table is hashed.
allv = np.arange(1 << 24, dtype=np.int64)
R, G, B = allv >> 16, (allv >> 8) & 255, allv & 255
Y, CB, CR = M.rgb2ycbcr(R, G, B)
assert Y.min() >= 0 and Y.max() <= 255 and CB.min() >= 0 and CB.max() <= 255 and CR.min() >= 0 and CR.max() <= 255
crit = np.nonzero(((299 * R + 587 * G + 114 * B) % 1000 == 0) |
                  ((5000 * R - 4187 * G - 813 * B) % 10000 == 0) |
                  ((-1687 * R - 3313 * G + 5000 * B) % 10000 == 0))[0]
sample = np.concatenate([rng.integers(0, 1 << 24, 200000), crit])
mism = 0
for s in range(0, len(sample), 64):
    idx = sample[s:s + 64]
    y, cb, cr = ef.rgb2ycbcr(list(R[idx]), list(G[idx]), list(B[idx]))
    k = len(idx)
    mism += int((np.array(y[:k]) != Y[idx]).sum() + (np.array(cb[:k]) != CB[idx]).sum() +
                (np.array(cr[:k]) != CR[idx]).sum())
meta["rgb2ycbcr_np_vs_model_mismatches"] = mism
meta["rgb2ycbcr_checked_inputs"] = int(len(sample))
assert mism == 0
h = hashlib.sha256()
for a in (Y, CB, CR):
    h.update(a.astype(np.uint8).tobytes())
meta["rgb2ycbcr_all_2p24_sha256"] = h.hexdigest()
# integer identities the CUDA kernel relies on (hard part 1)
ny = 299 * R + 587 * G + 114 * B
ncb = -1687 * R - 3313 * G + 5000 * B + 1280000
ncr = 5000 * R - 4187 * G - 813 * B + 1280000
ey = np.nonzero(Y != ny // 1000)[0]
ecb = np.nonzero(CB != ncb // 10000)[0]
ecr = np.nonzero(CR != ncr // 10000)[0]
meta["int_identity_exceptions"] = {"y": int(len(ey)), "cb": int(len(ecb)), "cr": int(len(ecr))}
assert (ny[ey] % 1000 == 0).all() and (Y[ey] == ny[ey] // 1000 - 1).all()
assert (ncr[ecr] % 10000 == 0).all() and (CR[ecr] == ncr[ecr] // 10000 - 1).all()
assert len(ecb) == 0

# ---- 4. fixtures of test/common.py and test/model/reference.py -----------------
d = refcommon.DCTData(64, 12)
q = refcommon.Quantizer()
z = refcommon.ZZData()
r = refcommon.RLE()
rle_names = ["red_pixels_1", "red_pixels_2", "green_pixels_1", "green_pixels_2", "blue_pixels_1", "blue_pixels_2"]
fx = dict(
    dct_input=np.array(d.input_dct), dct_output_paper=np.array(d.output_dct), dct_output_model=np.array(d.output_dct_model),
    quantizer_input=np.array(q.quantizer_input), quantizer_table=np.array(q.quantizer_table),
    quantizer_cr=np.array(q.quantizer_cr), quantize_output_ref=np.array(q.quantize_output_ref),
    quantizer_output2_ref=np.array(q.quantizer_output2_ref), quantizer_output=np.array(q.quantizer_output),
    zigzag_input=np.array(z.zigzag_input), zigzag_output=np.array(z.zigzag_output),
    rle_matrices=np.array([getattr(r, nme) for nme in rle_names]),
    rle_output_red_pixels_1=np.array(r.output_red_pixels_1),
)
for nme in ("r444_16x8_blocks", "g444_16x8_blocks", "b444_16x8_blocks", "y444_16x8_blocks", "cb444_16x8_blocks",
            "cr444_16x8_blocks", "y422_16x8_blocks", "y422_dct_16x8_blocks", "y422_qwant_16x8_blocks",
            "y422_zigzag_16x8_blocks"):
    fx["ref_" + nme] = np.array(getattr(refvec, nme))
for nme in ("cb422_16x8_block", "cr422_16x8_block", "cb422_dct_16x8_block", "cr422_dct_16x8_block",
            "cb422_qwant_16x8_block", "cr422_qwant_16x8_block", "cb_zigzag_16x8_block", "cr_zigzag_16x8_block"):
    fx["ref_" + nme] = np.array(getattr(refvec, nme))
# RAWImage (lena resized to 64x64) through both colour models of test/common.py:133-154
raw = refcommon.RAWImage({"ca": 0.1819, "cb": 0.0618, "cc": 0.6495, "cd": 0.5512, "yoffset": 16, "coffset": 128},
                         os.path.join(REF, "test", "lena.png"), 64)
fx["raw64_rgb"] = np.array([raw.r, raw.g, raw.b])
fx["raw64_ycbcr_xapp930_model"] = np.array(raw.rgb2ycbcr_model())
fx["raw64_ycbcr_jfif"] = np.array(raw.rgb2ycbcr())
# decode direction: both inverse colour models of RAWImage (test/common.py:158-178)
raw.rgb2ycbcr()
fx["raw64_rgb_from_jfif_via_wikipedia"] = np.array(raw.ycbcr2rgb())
from litejpeg_csc_coefs import ycbcr2rgb_coefs as _c2  # noqa: E402  (copied constants, see below)
raw2 = refcommon.RAWImage(_c2(8), os.path.join(REF, "test", "lena.png"), 64)
raw2.rgb2ycbcr()
fx["raw64_rgb_from_jfif_via_xapp931_model"] = np.array(raw2.ycbcr2rgb_model())
np.savez_compressed(os.path.join(HERE, "reference_fixtures.npz"), **fx)
meta["rle_names"] = rle_names

with open(os.path.join(HERE, "meta.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print(json.dumps(meta, indent=1, sort_keys=True))
