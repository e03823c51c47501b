"""B200-native baseline-JPEG encode path behind LiteJPEG's stage-operator surface.

Import name: `litejpeg_b200` (see litejpeg_b200.py at the repository root; the
directory name contains a '-' and cannot be imported with a plain statement).
"""
from . import _lib                                                     # noqa: F401
from ._lib import B2JError, build, lib, make_desc, geometry, quant_tables   # noqa: F401
from .common import (Packet, coef, rgb_layout, ycbcr444_layout, ycbcr422_layout, dct_block_layout,   # noqa: F401
                     block_layout, to_signed, to_unsigned)
from .core import (RGB2YCbCr, YCbCr2RGB, rgb2ycbcr_coefs, ycbcr2rgb_coefs, YCbCr444to422, YCbCr422to444, YCbCr444to420, DCT, Quantization, quant_values,  # noqa: F401
                   ZigZag, zigzag_order, zigzag_rom, RunLength, EntropyCoder, RLEMain, Huffman, BitPack)
from . import model                                                    # noqa: F401
from .ingest import to_rgb, read_image, decode_png, png_info                                 # noqa: F401
from .encoder import (JpegEncoder, MultiGpuJpegEncoder, ShardedJpegEncoder, shard_range, gather_sizes, gather_streams,   # noqa: F401
                      band_plan, assemble_bands)

__version__ = "0.1.0"
