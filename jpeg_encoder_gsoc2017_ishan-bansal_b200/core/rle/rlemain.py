"""RLE top operator.

Mirrors `RLEMain()` of litejpeg/core/rle/rlemain.py:24-80: the same input goes
to RunLength and EntropyCoder; output slots are 21-bit words amp[0:12]
size[12:16] run[16:20] valid[20] (decoder: test/common.py:455-471).
"""
from ...common import block_layout, to_signed
from .rlecore import RunLength

datapath_latency = 3      # rlemain.py:21


class RLEMain(RunLength):
    _entry = "b2j_rle_main"
    _width = 21

    def __init__(self):
        super().__init__()
        self.source_layout = block_layout(21)
        self.latency = datapath_latency

    @staticmethod
    def decode(words):
        """test/common.py:455-471 `RLE.set_rledata`: valid slots -> (amplitude, runlength, size)."""
        out = []
        for w in words:
            w = int(w)
            if (w >> 20) & 1:
                out.append((to_signed([w & 0xFFF], 12)[0], (w >> 16) & 15, (w >> 12) & 15))
        return out
