"""ctypes binding of libb2j.so (the C ABI declared in include/b2j.h).

The product path has no CPU fallback: if the CUDA library is missing or a call
fails, the error is raised.  `build()` compiles the library in-tree with nvcc
for sm_100a (cross-compiles without a GPU).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("B2J_LIB") or os.path.join(_HERE, "libb2j.so")   # B2J_LIB: kernel-variant experiments

c_void_p, c_int, c_int32, c_int64, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int32,
                                               ctypes.c_int64, ctypes.c_size_t)


class B2JError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libb2j error %d: %s" % (code, message))
        self.code = code


class ImageDesc(ctypes.Structure):
    """b2j_image_desc (include/b2j.h)."""
    _fields_ = [("width", c_int32), ("height", c_int32), ("subsampling", c_int32), ("n_frames", c_int32),
                ("row_stride", c_int64), ("frame_stride", c_int64)]


class Geometry(ctypes.Structure):
    """b2j_geometry (include/b2j.h)."""
    _fields_ = [("h", c_int32), ("v", c_int32), ("mcus_x", c_int32), ("mcus_y", c_int32),
                ("blocks_per_mcu", c_int32), ("padded_width", c_int32), ("padded_height", c_int32),
                ("blocks_per_frame", c_int64), ("header_bytes", c_int32), ("reserved", c_int32)]


_P = c_void_p
_DESC = ctypes.POINTER(ImageDesc)
# name -> (restype, argtypes); every symbol include/b2j.h declares
SIGNATURES = {
    "b2j_version": (ctypes.c_char_p, []),
    "b2j_last_error": (c_size_t, [ctypes.c_char_p, c_size_t]),
    "b2j_get_geometry": (c_int, [_DESC, ctypes.POINTER(Geometry)]),
    "b2j_quant_tables": (c_int, [c_int, _P, _P]),
    "b2j_fastpath_constants": (c_int, [_P, _P, _P]),
    "b2j_rgb2ycbcr": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "b2j_ycbcr444to422": (c_int, [_P, _P, _P, _P, _P, c_int64, _P]),
    "b2j_ycbcr2rgb": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "b2j_ycbcr422to444": (c_int, [_P, _P, _P, _P, _P, c_int64, _P]),
    "b2j_chroma_subsample": (c_int, [_P, _P, c_int32, c_int32, c_int32, _P]),
    "b2j_dct": (c_int, [_P, _P, c_int64, _P]),
    "b2j_quantize": (c_int, [_P, _P, c_int64, _P, c_int, _P]),
    "b2j_zigzag": (c_int, [_P, _P, c_int64, _P]),
    "b2j_run_length": (c_int, [_P, _P, c_int64, c_int32, _P]),
    "b2j_entropy_size": (c_int, [_P, _P, c_int64, _P]),
    "b2j_rle_main": (c_int, [_P, _P, c_int64, c_int32, _P]),
    "b2j_rle_code": (c_int, [_P, _P, _P, c_int64, _P]),
    "b2j_dct_fixed": (c_int, [_P, _P, c_int64, _P]),
    "b2j_ingest": (c_int, [_P, c_int, c_int32, c_int32, c_int32, c_int64, c_int64, c_int64, _P, _P]),
    "b2j_png_info": (c_int, [_P, c_int64, _P, _P, _P, _P]),
    "b2j_png_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32, c_int64]),
    "b2j_png_decode": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, _P, c_size_t, _P, _P, _P]),
    "b2j_fused_block": (c_int, [_P, _DESC, _P, _P, _P, c_int, _P]),
    "b2j_block_bits": (c_int, [_P, _DESC, c_int, _P, _P]),
    "b2j_scan_workspace_bytes": (c_size_t, [_DESC]),
    "b2j_scan": (c_int, [_P, _DESC, _P, _P, _P, _P]),
    "b2j_bitpack_workspace_bytes": (c_size_t, [_DESC, c_int64]),
    "b2j_bitpack": (c_int, [_P, _P, _P, _DESC, _P, _P, c_int, _P, _P, c_int64, _P, _P]),
    "b2j_entropy_workspace_bytes": (c_size_t, [_DESC, c_int64]),
    "b2j_entropy_workspace_bytes_opts": (c_size_t, [_DESC, c_int64, c_int]),
    "b2j_entropy_encode": (c_int, [_P, _DESC, _P, _P, c_int, _P, _P, c_int64, _P, _P]),
    "b2j_encode_workspace_bytes": (c_size_t, [_DESC, c_int64]),
    "b2j_encode_workspace_bytes_opts": (c_size_t, [_DESC, c_int64, c_int]),
    "b2j_default_out_cap": (c_int64, [_DESC]),
    "b2j_encode": (c_int, [_P, _DESC, c_int, _P, _P, c_int, _P, c_size_t, _P, c_int64, _P, _P]),
    "b2j_encode_launch_count": (c_int, [_DESC]),
    "b2j_encoder_create": (c_int, [ctypes.POINTER(_P), _DESC, c_int, c_int, c_int64, c_int32]),
    "b2j_encoder_encode_host": (c_int, [_P, _P, c_int32, _P, _P]),
    "b2j_encoder_destroy": (c_int, [_P]),
    "b2j_multi_encoder_create": (c_int, [ctypes.POINTER(_P), _P, c_int32, _DESC, c_int, c_int, c_int64, c_int32]),
    "b2j_multi_encoder_encode_host": (c_int, [_P, _P, c_int32, _P, _P]),
    "b2j_multi_encoder_h2d_probe": (c_int, [_P, _P, c_int32, c_int32, ctypes.POINTER(ctypes.c_double)]),
    "b2j_multi_encoder_destroy": (c_int, [_P]),
}

# enums of include/b2j.h
SUB = {"444": 444, "422": 422, "420": 420}
ZRL = {"eager": 0, "standard": 1}
CSC = {"jfif": 0, "xapp930": 1}
QMODE = {"round": 0, "trunc": 1, "reciprocal": 2}
LAYOUT = {"rgb": 0, "bgr": 1, "rgba": 2, "bgra": 3, "planar": 4, "gray": 5}
DCT_MODE = {"model": 0, "fixed": 0x100}      # B2J_FUSED_DCT_FIXED / B2J_OPT_DCT_FIXED


def opt_restart(n):
    """B2J_OPT_RESTART(n): restart interval in MCUs for the option word (0 = none)."""
    n = int(n)
    if not 0 <= n <= 32767:
        raise ValueError("restart_interval must be in 0..32767 MCUs")
    return n << 16
FUSED = {"auto": 0, "simple": 1, "tiled": 2, "stream": 3, "pair": 4, "pair_wide": 4 | 0x200}

_lib = None


def build(verbose=False):
    """Compile csrc/*.cu into libb2j.so with nvcc (sm_100a, -lineinfo, -fmad=false)."""
    res = subprocess.run(["make", "-C", _CSRC, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode:
        raise RuntimeError("building libb2j.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb2j.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the encode path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().b2j_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(code):
    if code != 0:
        raise B2JError(code, last_error())


def make_desc(height, width, subsampling="444", n_frames=1, row_stride=0, frame_stride=0, layout="rgb"):
    """b2j_image_desc; layout "bgr" | "rgba" | "bgra" is OR-ed into the subsampling word (B2J_DESC_LAYOUT)."""
    sub = SUB[str(subsampling)] if str(subsampling) in SUB else int(subsampling)
    if layout not in ("rgb", "bgr", "rgba", "bgra"):
        raise ValueError("the encode path reads rgb, bgr, rgba or bgra frames in place; use to_rgb() for other layouts")
    return ImageDesc(int(width), int(height), sub | (LAYOUT[layout] << 16), int(n_frames), int(row_stride), int(frame_stride))


def geometry(desc):
    g = Geometry()
    check(lib().b2j_get_geometry(ctypes.byref(desc), ctypes.byref(g)))
    return g


def quant_tables(quality):
    """[ext] IJG-scaled Annex-K tables as two lists of 64 ints (natural order)."""
    qy = (ctypes.c_uint8 * 64)()
    qc = (ctypes.c_uint8 * 64)()
    check(lib().b2j_quant_tables(int(quality), qy, qc))
    return list(qy), list(qc)


def table_arg(table):
    arr = (ctypes.c_uint8 * 64)(*[int(x) for x in table])
    return arr
