"""Size/category operator.

Mirrors `EntropyCoder()` of litejpeg/core/rle/entropycoder.py:69-225 (datapath
:19-65): 12-bit words in, 4-bit words out, size = number of bits of |x|.
(Despite its name the reference class is not a Huffman coder.)
"""
import torch

from ... import _lib
from ...common import Packet, block_layout, dev_tensor, require_cuda, stream_ptr, to_signed

datapath_latency = 3      # entropycoder.py:17


class EntropyCoder:
    def __init__(self):
        self.sink_layout = block_layout(12)
        self.source_layout = block_layout(4)
        self.latency = datapath_latency

    def process(self, x, stream=None):
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        t = dev_tensor(to_signed(x, 12) if packet else x, torch.int16)
        out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().b2j_entropy_size(t.data_ptr(), out.data_ptr(), t.numel(), stream_ptr(stream)))
        if packet:
            return Packet(out.reshape(-1).cpu().tolist())
        return out
