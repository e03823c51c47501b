"""Run-length operator.

Mirrors `RunLength()` of litejpeg/core/rle/rlecore.py:149-310 (datapath :30-146):
64 x 12-bit words in, 64 slots of 17-bit words out, amp[0:12] run[12:16]
valid[16].  The DC predictor `prev_dc_0` is a register in the gateware, so it
persists across `process` calls here too (reset() clears it).
"""
import torch

from ... import _lib
from ...common import Packet, block_layout, dev_tensor, require_cuda, stream_ptr, to_signed

datapath_latency = 3      # rlecore.py:21


class RunLength:
    _entry = "b2j_run_length"
    _width = 17

    def __init__(self):
        self.sink_layout = block_layout(12)
        self.source_layout = block_layout(self._width)
        self.latency = datapath_latency
        self.prev_dc = 0

    def reset(self):
        self.prev_dc = 0

    def process(self, x, stream=None):
        """Packet of 64*k 12-bit two's complement words (blocks of one component, zig-zag
        order, in scan order) -> Packet of 64*k output words; or int16 CUDA tensor
        (k, 64) -> int32 tensor (k, 64)."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        t = dev_tensor(to_signed(x, 12) if packet else x, torch.int16)
        if t.numel() % 64 or t.numel() == 0:
            raise ValueError("input must be a non-empty multiple of 64 words")
        t = t.reshape(-1, 64)
        words = torch.empty(t.shape, dtype=torch.int32, device=t.device)
        with torch.cuda.device(t.device):
            fn = getattr(_lib.lib(), self._entry)
            _lib.check(fn(t.data_ptr(), words.data_ptr(), t.shape[0], int(self.prev_dc), stream_ptr(stream)))
        self.prev_dc = int(t[-1, 0].item())
        if packet:
            return Packet(words.reshape(-1).cpu().tolist())
        return words

    @staticmethod
    def decode(words):
        """test/common.py:441-453 `RLE.setdata`: valid slots -> list of (amplitude, runlength)."""
        out = []
        for w in words:
            w = int(w)
            if (w >> 16) & 1:
                out.append((to_signed([w & 0xFFF], 12)[0], (w >> 12) & 15))
        return out
