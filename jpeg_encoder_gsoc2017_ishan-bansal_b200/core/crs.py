"""Chroma resampling operators.

`YCbCr444to422(dw=8)` mirrors litejpeg/core/crs.py:68-84 (datapath :8-65): pair
mean (c0 + c1) >> 1, Cb on even and Cr on odd pixels of the (y, cb_cr) stream.
`YCbCr444to420` is [ext]: the same mean applied horizontally, then vertically,
on whole planes (what the fused encode path does for 4:2:0).
"""
import torch

from .. import _lib
from ..common import Packet, dev_tensor, require_cuda, stream_ptr, ycbcr422_layout, ycbcr444_layout


class YCbCr444to422:
    latency = 3           # crs.py:17

    def __init__(self, dw=8):
        if dw != 8:
            raise ValueError("the B200 path implements dw = 8 (the reference's default)")
        self.sink_layout = ycbcr444_layout(dw)
        self.source_layout = ycbcr422_layout(dw)

    def process(self, x, stream=None):
        """Packet of 24-bit words y[16:24] cb[8:16] cr[0:8] (test/ycbcr_resampling_tb.py:19-22)
        -> Packet of 16-bit words y[8:16] cb_cr[0:8]; or uint8 CUDA tensor (n, 3) -> (n, 2)."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        if packet:
            w = torch.tensor(list(x), dtype=torch.int64)
            ycc = torch.stack([(w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF], dim=-1).to(torch.uint8)
        else:
            ycc = x
        ycc = dev_tensor(ycc, torch.uint8)
        if ycc.dim() != 2 or ycc.shape[1] != 3 or ycc.shape[0] % 2:
            raise ValueError("expected an even number of (y, cb, cr) pixels, shape (n, 3)")
        n = ycc.shape[0]
        planes = ycc.t().contiguous()
        out = torch.empty((2, n), dtype=torch.uint8, device=ycc.device)
        with torch.cuda.device(ycc.device):
            _lib.check(_lib.lib().b2j_ycbcr444to422(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(),
                                                     out[0].data_ptr(), out[1].data_ptr(), n, stream_ptr(stream)))
        res = out.t().contiguous()
        if packet:
            o = res.to(torch.int64).cpu()
            return Packet(((o[:, 0] << 8) | o[:, 1]).tolist())
        return res


class YCbCr422to444:
    """Decode-direction resampler.  Mirrors `YCbCr422to444(dw=8)` of
    litejpeg/core/crs.py:130-146 (datapath :88-128): sample repeat."""
    latency = 2           # crs.py:97

    def __init__(self, dw=8):
        if dw != 8:
            raise ValueError("the B200 path implements dw = 8 (the reference's default)")
        self.sink_layout = ycbcr422_layout(dw)
        self.source_layout = ycbcr444_layout(dw)

    def process(self, x, stream=None):
        """Packet of 16-bit words y[8:16] cb_cr[0:8] -> Packet of 24-bit words y[16:24] cb[8:16] cr[0:8];
        or uint8 CUDA tensor (n, 2) -> (n, 3)."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        if packet:
            w = torch.tensor(list(x), dtype=torch.int64)
            ycc = torch.stack([(w >> 8) & 0xFF, w & 0xFF], dim=-1).to(torch.uint8)
        else:
            ycc = x
        ycc = dev_tensor(ycc, torch.uint8)
        if ycc.dim() != 2 or ycc.shape[1] != 2 or ycc.shape[0] % 2:
            raise ValueError("expected an even number of (y, cb_cr) pixels, shape (n, 2)")
        n = ycc.shape[0]
        planes = ycc.t().contiguous()
        out = torch.empty((3, n), dtype=torch.uint8, device=ycc.device)
        with torch.cuda.device(ycc.device):
            _lib.check(_lib.lib().b2j_ycbcr422to444(planes[0].data_ptr(), planes[1].data_ptr(), out[0].data_ptr(),
                                                     out[1].data_ptr(), out[2].data_ptr(), n, stream_ptr(stream)))
        res = out.t().contiguous()
        if packet:
            o = res.to(torch.int64).cpu()
            return Packet(((o[:, 0] << 16) | (o[:, 1] << 8) | o[:, 2]).tolist())
        return res


class YCbCr444to420:
    """[ext] plane subsampler: uint8 CUDA tensor (H, W) -> (H/2, W/2) (or (H, W/2) with mode='422')."""
    latency = 0

    def __init__(self, dw=8, mode="420"):
        if dw != 8 or mode not in ("420", "422"):
            raise ValueError("dw must be 8 and mode '420' or '422'")
        self.mode = mode

    def process(self, plane, stream=None):
        require_cuda()
        p = dev_tensor(plane, torch.uint8)
        h, w = p.shape
        out = torch.empty((h // 2 if self.mode == "420" else h, w // 2), dtype=torch.uint8, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().b2j_chroma_subsample(p.data_ptr(), out.data_ptr(), h, w, _lib.SUB[self.mode],
                                                        stream_ptr(stream)))
        return out
