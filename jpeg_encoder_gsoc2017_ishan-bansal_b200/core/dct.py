"""8x8 DCT operator.

Mirrors the surface of `DCT(dw=12, dct_block=64)` (litejpeg/core/new_dct6.py:400-525):
64 words in, 64 words out, level shift -128, final /8.  The arithmetic is the
software model's float64 Loeffler graph (test/model/enc_frame.py:42-109), bit
for bit - not the gateware's 12-bit fixed point, which has two wiring bugs
(new_dct6.py:262-263, :294) and is not a usable arithmetic spec.
"""
import torch

from .. import _lib
from ..common import Packet, block_layout, dev_tensor, require_cuda, stream_ptr, to_unsigned

datapath_latency = 8      # new_dct6.py:20


class DCT:
    def __init__(self, dw=12, dct_block=64):
        if dct_block != 64:
            raise ValueError("dct_block must be 64")
        self.dw = dw
        self.sink_layout = block_layout(dw)
        self.source_layout = block_layout(dw)
        self.latency = datapath_latency

    def process(self, x, stream=None):
        """Packet of 64*k samples (0..255, row-major per block) -> Packet of dw-bit two's
        complement words (decode with common.to_signed, as test/common.py:264-270);
        or int16 CUDA tensor (..., 64) -> int16 tensor of signed coefficients."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        t = dev_tensor(x, torch.int16)
        if t.numel() % 64:
            raise ValueError("DCT input must be a multiple of 64 words")
        out = torch.empty_like(t)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().b2j_dct(t.data_ptr(), out.data_ptr(), t.numel() // 64, stream_ptr(stream)))
        if packet:
            return to_unsigned(out.cpu().tolist(), self.dw)
        return out
