"""Host side of the PNG ingest (no GPU): the chunk walk of b2j_png_info and the workspace query."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import png_util as U                                                   # noqa: E402


def test_png_info_reads_the_header_and_sums_the_idat_payloads(lj):
    a = np.zeros((13, 21, 4), np.uint8)
    f = U.write_png(a, idat_split=7)
    w, h, c, z = lj.png_info(f)
    import zlib
    assert (w, h, c) == (21, 13, 4)
    assert z == len(zlib.compress(U.filter_rows(a, [0] * 13), 6))
    assert lj.png_info(U.write_png(a[..., 0]))[:3] == (21, 13, 1)


def test_png_info_refuses_what_the_device_decoder_does_not_take(lj):
    a = np.zeros((4, 4, 3), np.uint8)
    for bad in (U.write_png(a, depth=16), U.write_png(a, interlace=1), b"not a png at all" * 4, U.write_png(a)[:40]):
        with pytest.raises(RuntimeError):
            lj.png_info(bad)


def test_png_workspace_query(lj):
    L = lj.lib()
    assert L.b2j_png_workspace_bytes(512, 512, 3, 4, 500000) >= 4 * (500000 + 512 * (1 + 512 * 3))
    assert L.b2j_png_workspace_bytes(0, 512, 3, 4, 1) == 0
    assert L.b2j_png_workspace_bytes(512, 512, 2, 4, 1) == 0
