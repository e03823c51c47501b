"""[ext] Entropy-coding operators of the fused path (no counterpart class in the
reference: its `huffman` is an identity stub, test/model/enc_frame.py:179-180).

They work on the fused coefficient layout: int16 CUDA tensor
(n_frames, blocks_per_frame, 64), zig-zag order, blocks in scan order.

Huffman  : K2 `b2j_block_bits` - Huffman-coded size of every block (DC difference,
           gateware run-length rules rlecore.py:84-132, sizes entropycoder.py:55-65,
           Annex-K code lengths).
BitPack  : K3 `b2j_scan` + K4 `b2j_bitpack` - exclusive scan of the sizes, then the
           packed, 0xFF-stuffed scan with JFIF header and EOI.
"""
import ctypes

import torch

from .. import _lib
from ..common import require_cuda, stream_ptr


def _check_coefs(coefs, desc):
    g = _lib.geometry(desc)
    if coefs.dtype != torch.int16 or not coefs.is_cuda or not coefs.is_contiguous():
        raise ValueError("coefs must be a contiguous int16 CUDA tensor")
    if coefs.numel() != desc.n_frames * g.blocks_per_frame * 64:
        raise ValueError("coefs has %d elements, geometry needs %d" % (coefs.numel(), desc.n_frames * g.blocks_per_frame * 64))
    return g


class Huffman:
    latency = 0

    def __init__(self, zrl="eager"):
        if zrl not in _lib.ZRL:
            raise ValueError("zrl must be 'eager' or 'standard'")
        self.zrl = zrl

    def process(self, coefs, desc, stream=None):
        """-> int32 CUDA tensor (n_frames, blocks_per_frame): bits per block."""
        require_cuda()
        g = _check_coefs(coefs, desc)
        bits = torch.empty((desc.n_frames, g.blocks_per_frame), dtype=torch.int32, device=coefs.device)
        with torch.cuda.device(coefs.device):
            _lib.check(_lib.lib().b2j_block_bits(coefs.data_ptr(), ctypes.byref(desc), _lib.ZRL[self.zrl], bits.data_ptr(),
                                                  stream_ptr(stream)))
        return bits


class BitPack:
    latency = 0

    def __init__(self, zrl="eager"):
        if zrl not in _lib.ZRL:
            raise ValueError("zrl must be 'eager' or 'standard'")
        self.zrl = zrl

    def scan(self, bits, desc, stream=None):
        """-> (offsets uint32-as-int32 (n_frames, blocks_per_frame), total_bits int64 (n_frames,))."""
        require_cuda()
        g = _lib.geometry(desc)
        offsets = torch.empty((desc.n_frames, g.blocks_per_frame), dtype=torch.int32, device=bits.device)
        total = torch.empty((desc.n_frames,), dtype=torch.int64, device=bits.device)
        ws = torch.empty((_lib.lib().b2j_scan_workspace_bytes(ctypes.byref(desc)),), dtype=torch.uint8, device=bits.device)
        with torch.cuda.device(bits.device):
            _lib.check(_lib.lib().b2j_scan(bits.data_ptr(), ctypes.byref(desc), offsets.data_ptr(), total.data_ptr(),
                                            ws.data_ptr(), stream_ptr(stream)))
        return offsets, total

    def process(self, coefs, bits, desc, qy, qc, out_cap=None, stream=None):
        """-> (out uint8 (n_frames, out_cap), sizes int64 (n_frames,)): whole JFIF files."""
        require_cuda()
        _check_coefs(coefs, desc)
        L = _lib.lib()
        offsets, total = self.scan(bits, desc, stream)
        cap = int(out_cap or L.b2j_default_out_cap(ctypes.byref(desc)))
        ws = torch.empty((L.b2j_bitpack_workspace_bytes(ctypes.byref(desc), cap),), dtype=torch.uint8, device=coefs.device)
        out = torch.empty((desc.n_frames, cap), dtype=torch.uint8, device=coefs.device)
        sizes = torch.empty((desc.n_frames,), dtype=torch.int64, device=coefs.device)
        with torch.cuda.device(coefs.device):
            _lib.check(L.b2j_bitpack(coefs.data_ptr(), offsets.data_ptr(), total.data_ptr(), ctypes.byref(desc),
                                     _lib.table_arg(qy), _lib.table_arg(qc), _lib.ZRL[self.zrl], ws.data_ptr(),
                                     out.data_ptr(), cap, sizes.data_ptr(), stream_ptr(stream)))
        return out, sizes
