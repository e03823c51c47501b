"""Zig-zag reordering operator.

Mirrors `ZigZag()` of litejpeg/core/zigzag.py:37-212 and `zigzag` of
test/model/enc_frame.py:161-165: out[zigzag_order[i]] = in[i].
"""
import torch

from .. import _lib
from ..common import Packet, block_layout, dev_tensor, require_cuda, stream_ptr

# litejpeg/core/zigzag.py:19-27 == test/model/enc_frame.py:131-139
zigzag_order = [
    0, 1, 5, 6, 14, 15, 27, 28,
    2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43,
    9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,
    20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,
    35, 36, 48, 49, 57, 58, 62, 63]

# litejpeg/core/zigzag.py:30-34: natural index read at zig-zag position i
zigzag_rom = [zigzag_order.index(i) for i in range(64)]


class ZigZag:
    latency = 0

    def __init__(self, dw=12):
        self.sink_layout = block_layout(dw)
        self.source_layout = block_layout(dw)

    def process(self, x, stream=None):
        """Packet of 64*k words -> Packet (words are moved, not interpreted); or int16 CUDA tensor."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        t = dev_tensor(x, torch.int16)
        if t.numel() % 64:
            raise ValueError("ZigZag input must be a multiple of 64 words")
        out = torch.empty_like(t)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().b2j_zigzag(t.data_ptr(), out.data_ptr(), t.numel() // 64, stream_ptr(stream)))
        if packet:
            return Packet(out.cpu().tolist())
        return out
