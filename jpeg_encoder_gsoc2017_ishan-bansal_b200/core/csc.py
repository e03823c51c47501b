"""Colour space conversion operator.

Mirrors `RGB2YCbCr(rgb_w=8, ycbcr_w=8, coef_w=8)` of litejpeg/core/csc.py:163-190
(same constructor, `.latency`, sink/source layouts) and runs on the GPU through
`b2j_rgb2ycbcr`.  model="xapp930" is the gateware datapath (csc.py:85-160);
model="jfif" is the software model (test/model/enc_frame.py:27-35), which is
what the fused encode path uses.
"""
import torch

from .. import _lib
from ..common import Packet, coef, dev_tensor, require_cuda, rgb_layout, stream_ptr, ycbcr444_layout

datapath_latency = 8      # csc.py:38


def rgb2ycbcr_coefs(dw, cw=None):
    """litejpeg/core/csc.py:22-35."""
    return {
        "ca": coef(0.1819, cw), "cb": coef(0.0618, cw), "cc": coef(0.6495, cw), "cd": coef(0.5512, cw),
        "yoffset": 2 ** (dw - 4), "coffset": 2 ** (dw - 1),
        "ymax": 2 ** dw - 1, "cmax": 2 ** dw - 1, "ymin": 0, "cmin": 0,
    }


class RGB2YCbCr:
    def __init__(self, rgb_w=8, ycbcr_w=8, coef_w=8, model="xapp930"):
        if (rgb_w, ycbcr_w, coef_w) != (8, 8, 8):
            raise ValueError("the B200 path implements rgb_w = ycbcr_w = coef_w = 8 (the reference's defaults)")
        if model not in _lib.CSC:
            raise ValueError("model must be 'xapp930' or 'jfif'")
        self.sink_layout = rgb_layout(rgb_w)
        self.source_layout = ycbcr444_layout(ycbcr_w)
        self.latency = datapath_latency
        self.model = model

    def process(self, x, stream=None):
        """Packet of 24-bit words r[16:24] g[8:16] b[0:8] (test/rgb2ycbcr_tb.py:29-36)
        -> Packet of y[16:24] cb[8:16] cr[0:8]; or uint8 CUDA tensor (..., 3) -> same shape."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        if packet:
            w = torch.tensor(list(x), dtype=torch.int64)
            rgb = torch.stack([(w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF], dim=-1).to(torch.uint8)
        else:
            rgb = x
        rgb = dev_tensor(rgb, torch.uint8)
        if rgb.shape[-1] != 3:
            raise ValueError("last dimension must be 3 (r, g, b)")
        n = rgb.numel() // 3
        if n == 0:
            return Packet() if packet else rgb.clone()
        planes = torch.empty((3, n), dtype=torch.uint8, device=rgb.device)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().b2j_rgb2ycbcr(rgb.data_ptr(), planes[0].data_ptr(), planes[1].data_ptr(),
                                                 planes[2].data_ptr(), n, _lib.CSC[self.model], stream_ptr(stream)))
        out = planes.t().contiguous().reshape(rgb.shape)
        if packet:
            o = out.to(torch.int64).cpu()
            return Packet(((o[:, 0] << 16) | (o[:, 1] << 8) | o[:, 2]).tolist())
        return out


def ycbcr2rgb_coefs(dw, cw=None):
    """litejpeg/core/csc.py:194-216."""
    ca, cb, cc, cd = 0.1819, 0.0618, 0.6495, 0.5512
    xcoef_w = None if cw is None else cw - 2
    return {
        "ca": coef(ca, cw), "cb": coef(cb, cw), "cc": coef(cc, cw), "cd": coef(cd, cw),
        "yoffset": 2 ** (dw - 4), "coffset": 2 ** (dw - 1),
        "ymax": 2 ** dw - 1, "cmax": 2 ** dw - 1, "ymin": 0, "cmin": 0,
        "acoef": coef(1 / cd, xcoef_w), "bcoef": coef(-cb / (cc * (1 - ca - cb)), xcoef_w),
        "ccoef": coef(-ca / (cd * (1 - ca - cb)), xcoef_w), "dcoef": coef(1 / cc, xcoef_w),
    }


class YCbCr2RGB:
    """Decode-direction colour operator.  Mirrors `YCbCr2RGB(ycbcr_w=8, rgb_w=8, coef_w=8)`
    of litejpeg/core/csc.py:313-337.  model="xapp931" is the gateware datapath (:219-310, saturated);
    model="jfif" is `RAWImage.ycbcr2rgb` (test/common.py:172-178: float64, int(), not clamped)."""

    def __init__(self, ycbcr_w=8, rgb_w=8, coef_w=8, model="xapp931"):
        if (ycbcr_w, rgb_w, coef_w) != (8, 8, 8):
            raise ValueError("the B200 path implements ycbcr_w = rgb_w = coef_w = 8 (the reference's defaults)")
        if model not in ("xapp931", "jfif"):
            raise ValueError("model must be 'xapp931' or 'jfif'")
        self.sink_layout = ycbcr444_layout(ycbcr_w)
        self.source_layout = rgb_layout(rgb_w)
        self.latency = 4          # csc.py:219 datapath: 4 registered stages
        self.model = model

    def process(self, x, stream=None):
        """Packet of 24-bit words y[16:24] cb[8:16] cr[0:8] (test/ycbcr2rgb_tb.py:17-20) -> Packet of
        r[16:24] g[8:16] b[0:8] (fields masked to 8 bits like RAWImage.pack_rgb, test/common.py:78-89);
        or uint8 CUDA tensor (..., 3) -> int16 tensor (..., 3)."""
        require_cuda()
        packet = not isinstance(x, torch.Tensor)
        if packet:
            w = torch.tensor(list(x), dtype=torch.int64)
            ycc = torch.stack([(w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF], dim=-1).to(torch.uint8)
        else:
            ycc = x
        ycc = dev_tensor(ycc, torch.uint8)
        if ycc.shape[-1] != 3:
            raise ValueError("last dimension must be 3 (y, cb, cr)")
        n = ycc.numel() // 3
        if n == 0:
            return Packet() if packet else ycc.to(torch.int16)
        planes = ycc.reshape(-1, 3).t().contiguous()
        out = torch.empty((3, n), dtype=torch.int16, device=ycc.device)
        with torch.cuda.device(ycc.device):
            _lib.check(_lib.lib().b2j_ycbcr2rgb(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(),
                                                 out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n,
                                                 _lib.CSC["xapp930" if self.model == "xapp931" else "jfif"], stream_ptr(stream)))
        res = out.t().contiguous().reshape(ycc.shape)
        if packet:
            o = res.to(torch.int64).cpu()
            return Packet((((o[:, 0] & 0xFF) << 16) | ((o[:, 1] & 0xFF) << 8) | (o[:, 2] & 0xFF)).tolist())
        return res
