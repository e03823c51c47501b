"""Stage operators, one class per stage of litejpeg/core of the reference."""
from .csc import RGB2YCbCr, YCbCr2RGB, rgb2ycbcr_coefs, ycbcr2rgb_coefs   # noqa: F401
from .crs import YCbCr444to422, YCbCr422to444, YCbCr444to420   # noqa: F401
from .dct import DCT                                   # noqa: F401
from .quantization import Quantization, quant_values  # noqa: F401
from .zigzag import ZigZag, zigzag_order, zigzag_rom   # noqa: F401
from .rle import RunLength, EntropyCoder, RLEMain      # noqa: F401
from .huffman import Huffman, BitPack                  # noqa: F401
