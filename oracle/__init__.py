"""CPU oracle for the B200 JPEG encode path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`jpeg_encoder_gsoc2017_ishan-bansal_b200/`)
imports this.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import, link or execute it, and there
only as the checker / CPU baseline, never as the thing measured or shipped.

What it restates (all citations are relative to /root/reference):

* `oracle.model_np`  - the reference software model `test/model/enc_frame.py`
  (rgb2ycbcr :27-35, dct :42-109, quantize :111-129, zigzag :161-165, DC
  differencing :219-230, rle_code :175-176) in NumPy float64, same operation
  order, no FMA.  PINNED: `tests/golden/make_golden.py` imports the real model
  in the build container and the committed fixtures hold its outputs.
* `oracle.gateware`  - integer restatements of the gateware operators the model
  has no counterpart for (litejpeg/core/csc.py, crs.py, quantization.py,
  zigzag.py, rle/rlecore.py, rle/entropycoder.py, rle/rlemain.py).  The gateware
  cannot be simulated here (LiteX/Migen absent, unpinned), so these are pinned
  only by the fixture matrices in test/common.py and by the rules in the source.
* `oracle.jpeg_ext`  - everything the reference never implemented ([ext]):
  4:2:2 / 4:2:0 cascade pair-mean, IJG quality scaling, Annex-K Huffman coding,
  bit packing, 0xFF stuffing, JFIF container, and a baseline decoder.
  PARITY UNPINNED against the reference (it has no such code); pinned instead
  by decode round trips (own decoder, PIL, cv2) and the q=85 table identity.
* `oracle/c/b2j_oracle.c` - plain-C restatement of the same pipeline (OpenMP),
  used for full-size checks and as the timed CPU baseline ("port").
"""
