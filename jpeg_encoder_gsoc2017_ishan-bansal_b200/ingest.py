"""[ext] Image ingest: the step in front of the encode path.

The reference reads images on the host with PIL and unpacks them in Python
(test/model/enc_frame.py:7-10 takes `arr[..., 0..2]`; test/common.py:47-51 unzips
`img.getdata()`).  Here 8-bit grey / RGB / RGBA PNG files are decoded on the GPU (`decode_png`,
C entry point `b2j_png_decode`: inflate + scanline filters, one warp per file, a batch of files
per call); `read_image` uses it for such files and PIL for everything else, and every layout
conversion runs on the GPU (`to_rgb`, C entry point `b2j_ingest`), so a frame grabber's BGRA or
planar buffers can be fed to `JpegEncoder` without a host pass.
"""
import ctypes

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


PNG_STATUS = {1: "bad zlib header", 2: "bad Huffman code or distance", 3: "more data than the image holds",
              4: "stream ends early", 5: "less data than the image holds", 6: "unknown filter type"}


def png_info(data):
    """(width, height, channels, zlib_bytes) of a PNG file held in `data` (bytes); ValueError for files the
    device decoder does not take (anything but 8-bit grey / RGB / RGBA, non-interlaced)."""
    w, h, c, z = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
    data = bytes(data)
    _lib.check(_lib.lib().b2j_png_info(ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p), len(data), ctypes.byref(w), ctypes.byref(h),
                                       ctypes.byref(c), ctypes.byref(z)))
    return w.value, h.value, c.value, z.value


def decode_png(files, device="cuda", stream=None):
    """PNG files (a list of bytes objects or paths, all of one size and colour type) -> CUDA uint8 RGB tensor
    (N, H, W, 3), decoded on the GPU: the host only gathers the IDAT payloads.  Raises ValueError when a
    stream is damaged (the message names the file and what the decoder found)."""
    require_cuda()
    blobs = []
    for f in files:
        if isinstance(f, (bytes, bytearray, memoryview)):
            blobs.append(bytes(f))
        else:
            with open(f, "rb") as fh:
                blobs.append(fh.read())
    if not blobs:
        raise ValueError("decode_png: no files")
    infos = [png_info(b) for b in blobs]
    w, h, c, _ = infos[0]
    n = len(blobs)
    dev = torch.device(device)
    L = _lib.lib()
    ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p) for b in blobs])   # (no copies: `blobs` stays alive)
    lens = (ctypes.c_int64 * n)(*[len(b) for b in blobs])
    status = (ctypes.c_int32 * n)()
    with torch.cuda.device(dev):
        need = L.b2j_png_workspace_bytes(w, h, c, n, max(i[3] for i in infos))
        ws = torch.empty((need + 256,), dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) & ~255
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        _lib.check(L.b2j_png_decode(ptrs, lens, n, w, h, c, wp, need, out.data_ptr(), status, stream_ptr(stream)))
    for i, st in enumerate(status):
        if st:
            raise ValueError("decode_png: file %d: %s" % (i, PNG_STATUS.get(st, "status %d" % st)))
    return out


def read_image(path, device="cuda", decoder="auto"):
    """Image file -> CUDA uint8 RGB tensor (H, W, 3).  decoder "auto": 8-bit grey / RGB / RGBA non-interlaced
    PNG files are decoded on the GPU (`decode_png`), any other file by PIL on the host (as the reference
    does), its palette / grey / RGBA pixels then brought to RGB on the GPU; "device" insists on the GPU
    decoder (ValueError for other files), "pil" on the host one."""
    if decoder not in ("auto", "device", "pil"):
        raise ValueError("decoder must be auto, device or pil")
    require_cuda()
    if decoder != "pil":
        with open(path, "rb") as fh:
            data = fh.read()
        try:
            png_info(data)
            return decode_png([data], device)[0]
        except (ValueError, RuntimeError):
            if decoder == "device":
                raise
    from PIL import Image
    import numpy as np
    img = Image.open(path)
    if img.mode not in ("RGB", "RGBA", "L"):
        img = img.convert("RGBA" if "A" in img.getbands() else "RGB")
    arr = torch.from_numpy(np.array(img)).to(device)
    return to_rgb(arr, {"RGB": "rgb", "RGBA": "rgba", "L": "gray"}[img.mode])
