"""Layouts and helpers shared by the stage operators.

Mirrors litejpeg/core/common.py of the reference (layouts :95-110, `coef`
:91-92) plus the stream-simulation `Packet` the benches use
(test/new_zigzag_tb.py:46-51): a Packet is a flat list of unsigned integers.
"""
import torch

from . import _lib


def coef(value, cw=None):
    """litejpeg/core/common.py:91-92."""
    return int(value * 2 ** cw) if cw is not None else value


def rgb_layout(dw):
    return [("r", dw), ("g", dw), ("b", dw)]


def ycbcr444_layout(dw):
    return [("y", dw), ("cb", dw), ("cr", dw)]


def ycbcr422_layout(dw):
    return [("y", dw), ("cb_cr", dw)]


def dct_block_layout(dw, ds):
    return [("dct_" + str(i), dw) for i in range(ds)]


def block_layout(dw):
    return [("data", dw)]


class Packet(list):
    """A flat list of unsigned words, as litex.soc.interconnect.stream_sim.Packet."""


def to_unsigned(values, dw=12):
    """Signed ints -> dw-bit two's complement words (what a gateware source emits)."""
    mask = (1 << dw) - 1
    return Packet(int(v) & mask for v in values)


def to_signed(values, dw=12):
    """dw-bit two's complement words -> signed ints (test/common.py:353-361)."""
    sign = 1 << (dw - 1)
    mask = (1 << dw) - 1
    out = []
    for v in values:
        v = int(v) & mask
        out.append(v - (1 << dw) if v & sign else v)
    return out


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("the B200 JPEG path needs a CUDA device (no CPU fallback)")
    _lib.lib()


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def dev_tensor(values, dtype, device=None):
    """Packet / list / tensor -> contiguous CUDA tensor of `dtype`."""
    if isinstance(values, torch.Tensor):
        t = values.to(device=device or "cuda", dtype=dtype)
    else:
        t = torch.tensor(list(values), dtype=torch.int64).to(dtype).to(device or "cuda")
    return t.contiguous()
