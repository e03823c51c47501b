"""[ext] Image ingest: the step in front of the encode path.

The reference reads images on the host with PIL and unpacks them in Python
(test/model/enc_frame.py:7-10 takes `arr[..., 0..2]`; test/common.py:47-51 unzips
`img.getdata()`).  Here the file is still decoded by PIL on the host (`read_image`), but
every layout conversion runs on the GPU (`to_rgb`, C entry point `b2j_ingest`), so a
frame grabber's BGRA or planar buffers can be fed to `JpegEncoder` without a host pass.
"""
import torch

from . import _lib
from .common import require_cuda, stream_ptr


def to_rgb(x, layout="rgb", stream=None):
    """CUDA uint8 tensor in `layout` -> packed interleaved RGB (N, H, W, 3) (or (H, W, 3)).

    layout: "rgb" / "bgr" (..., H, W, 3), "rgba" / "bgra" (..., H, W, 4; alpha ignored),
    "planar" (..., 3, H, W), "gray" (..., H, W).  Any strides along rows / planes / frames
    are honoured as long as the pixels of a row are contiguous."""
    require_cuda()
    if layout not in _lib.LAYOUT:
        raise ValueError("layout must be one of %s" % sorted(_lib.LAYOUT))
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8:
        raise ValueError("to_rgb takes a CUDA uint8 tensor")
    core = {"rgb": 3, "bgr": 3, "rgba": 3, "bgra": 3, "planar": 3, "gray": 2}[layout]
    if x.dim() not in (core, core + 1):
        raise ValueError("unexpected rank %d for layout %r" % (x.dim(), layout))
    single = x.dim() == core
    t = x[None] if single else x
    if layout in ("rgb", "bgr", "rgba", "bgra"):
        ch = 3 if layout in ("rgb", "bgr") else 4
        if t.shape[-1] != ch:
            raise ValueError("last dimension must be %d for layout %r" % (ch, layout))
        n, h, w = t.shape[0], t.shape[1], t.shape[2]
        if t.stride(3) != 1 or t.stride(2) != ch:
            t = t.contiguous()
        row, plane, frame = t.stride(1), 0, t.stride(0)
    elif layout == "planar":
        if t.shape[1] != 3:
            raise ValueError("planar input must be (..., 3, H, W)")
        n, h, w = t.shape[0], t.shape[2], t.shape[3]
        if t.stride(3) != 1:
            t = t.contiguous()
        row, plane, frame = t.stride(2), t.stride(1), t.stride(0)
    else:
        n, h, w = t.shape
        if t.stride(2) != 1:
            t = t.contiguous()
        row, plane, frame = t.stride(1), 0, t.stride(0)
    if n == 1:
        frame = 0                                   # a single frame: let the library derive the packed value
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b2j_ingest(t.data_ptr(), _lib.LAYOUT[layout], w, h, n, row, plane, frame, out.data_ptr(),
                                         stream_ptr(stream)))
    return out[0] if single else out


def read_image(path, device="cuda"):
    """Image file -> CUDA uint8 RGB tensor (H, W, 3).  PIL decodes on the host (as the reference
    does); palette / grey / RGBA files are brought to RGB on the GPU."""
    from PIL import Image
    import numpy as np
    require_cuda()
    img = Image.open(path)
    if img.mode not in ("RGB", "RGBA", "L"):
        img = img.convert("RGBA" if "A" in img.getbands() else "RGB")
    arr = torch.from_numpy(np.array(img)).to(device)
    return to_rgb(arr, {"RGB": "rgb", "RGBA": "rgba", "L": "gray"}[img.mode])
