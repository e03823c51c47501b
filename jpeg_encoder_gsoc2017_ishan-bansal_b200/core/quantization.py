"""Quantisation operator.

Mirrors `Quantization()` of litejpeg/core/quantization.py:29-182 (64 x 12-bit in
and out, table `quant_values` :20-27).  mode="round" is the software model
(test/model/enc_frame.py:111-129, round half away from zero) and the default;
mode="reciprocal" is the gateware datapath (quantization.py:58, 124-168);
mode="trunc" the older model (test/model/encoder.py:84-88).
"""
import torch

from .. import _lib
from ..common import Packet, block_layout, dev_tensor, require_cuda, stream_ptr, to_signed, to_unsigned

datapath_latency = 3

# litejpeg/core/quantization.py:20-27 (Annex-K luminance, unscaled)
quant_values = [16, 11, 10, 16, 24, 40, 51, 61,
                12, 12, 14, 19, 26, 58, 60, 55,
                14, 13, 16, 24, 40, 57, 69, 56,
                14, 17, 22, 29, 51, 87, 80, 62,
                18, 22, 37, 56, 68, 109, 103, 77,
                24, 35, 55, 64, 81, 104, 113, 92,
                49, 64, 78, 87, 103, 121, 120, 101,
                72, 92, 95, 98, 112, 100, 103, 99]


class Quantization:
    def __init__(self, table=None, mode="round", dw=12):
        if mode not in _lib.QMODE:
            raise ValueError("mode must be 'round', 'trunc' or 'reciprocal'")
        self.table = list(quant_values if table is None else table)
        if len(self.table) != 64 or min(self.table) < 1 or max(self.table) > 255:
            raise ValueError("table must hold 64 divisors in 1..255")
        self.mode = mode
        self.dw = dw
        self.sink_layout = block_layout(dw)
        self.source_layout = block_layout(dw)
        self.latency = datapath_latency

    def process(self, x, stream=None):
        """Packet of 64*k dw-bit two's complement words -> same; or int16 CUDA tensor -> int16."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        t = dev_tensor(to_signed(x, self.dw) if packet else x, torch.int16)
        if t.numel() % 64:
            raise ValueError("Quantization input must be a multiple of 64 words")
        out = torch.empty_like(t)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().b2j_quantize(t.data_ptr(), out.data_ptr(), t.numel() // 64, _lib.table_arg(self.table),
                                                _lib.QMODE[self.mode], stream_ptr(stream)))
        if packet:
            return to_unsigned(out.cpu().tolist(), self.dw)
        return out
