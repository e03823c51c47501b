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
    def __init__(self, dw=12, dct_block=64, mode="model"):
        """mode="model": the float64 software model (enc_frame.py:38-109), the parity path.
        mode="fixed": [ext] the gateware datapath's fixed-point arithmetic (new_dct6.py:171-396,
        its two wiring slips corrected, wide sums); <= 4 LSB from the model, not parity."""
        if dct_block != 64:
            raise ValueError("dct_block must be 64")
        if mode not in _lib.DCT_MODE:
            raise ValueError("mode must be 'model' or 'fixed'")
        self.mode = mode
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
            fn = _lib.lib().b2j_dct_fixed if self.mode == "fixed" else _lib.lib().b2j_dct
            _lib.check(fn(t.data_ptr(), out.data_ptr(), t.numel() // 64, stream_ptr(stream)))
        if packet:
            return to_unsigned(out.cpu().tolist(), self.dw)
        return out
