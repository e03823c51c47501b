"""The C-ABI library: it loads, exports every symbol include/b2j.h declares, and its
host-only entry points and argument checks behave.  No GPU needed (no kernel is launched)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import jpeg_ext as J

HEADER = open("include/b2j.h").read()


def declared_symbols():
    return sorted(set(re.findall(r"\b(b2j_[a-z0-9_]+)\s*\(", HEADER)))


def test_library_exports_every_declared_symbol(lj):
    lib = lj.lib()
    names = declared_symbols()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), "libb2j.so does not export %s" % n
    assert set(names) == set(lj._lib.SIGNATURES), set(names) ^ set(lj._lib.SIGNATURES)


def test_header_cites_the_reference_for_every_stage():
    for cite in ("csc.py", "crs.py", "new_dct6.py", "quantization.py", "zigzag.py", "rlecore.py", "entropycoder.py",
                 "rlemain.py", "enc_frame.py"):
        assert cite in HEADER


def test_version_and_geometry(lj):
    assert b"sm_100a" in lj.lib().b2j_version()
    for (h, w, sub) in ((1080, 1920, "420"), (512, 512, "444"), (2160, 3840, "422"), (37, 53, "420"), (1, 1, "444")):
        g = lj.geometry(lj.make_desc(h, w, sub))
        e = J.geometry(h, w, sub)
        assert (g.mcus_x, g.mcus_y, g.blocks_per_mcu, g.padded_width, g.padded_height) == \
               (e["mcus_x"], e["mcus_y"], e["blocks_per_mcu"], e["padded_w"], e["padded_h"])
        assert g.blocks_per_frame == e["n_mcu"] * e["blocks_per_mcu"]
        assert g.header_bytes == len(J.jfif_header(h, w, sub, *J.quant_tables(75)))
    g = lj.geometry(lj.make_desc(1080, 1920, "420"))
    assert (g.mcus_x, g.mcus_y, g.blocks_per_frame) == (120, 68, 48960)


def test_quant_tables_match_oracle(lj):
    for q in (1, 25, 50, 75, 85, 90, 100):
        qy, qc = lj.quant_tables(q)
        ey, ec = J.quant_tables(q)
        assert qy == ey.tolist() and qc == ec.tolist()


def test_argument_errors_are_reported_not_raised_in_c(lj):
    lib = lj.lib()
    d = lj.make_desc(16, 16, "444")
    bad = lj.make_desc(16, 16, 411)
    g = lj._lib.Geometry()
    assert lib.b2j_get_geometry(ctypes.byref(bad), ctypes.byref(g)) == -1
    assert "subsampling" in lj._lib.last_error()
    assert lib.b2j_get_geometry(ctypes.byref(lj.make_desc(0, 16, "444")), ctypes.byref(g)) == -2
    assert lib.b2j_quant_tables(0, (ctypes.c_uint8 * 64)(), (ctypes.c_uint8 * 64)()) == -1
    # null pointers and bad enums are refused before any CUDA call
    assert lib.b2j_dct(None, None, 1, None) == -1
    assert lib.b2j_rgb2ycbcr(None, None, None, None, 4, 0, None) == -1
    assert lib.b2j_fused_block(None, ctypes.byref(d), None, None, None, 0, None) == -1
    assert lib.b2j_block_bits(None, ctypes.byref(d), 0, None, None) == -1
    assert lib.b2j_encode(None, ctypes.byref(d), 75, None, None, 0, None, 0, None, 0, None, None) == -1
    with pytest.raises(lj.B2JError):
        lj._lib.check(lib.b2j_quant_tables(101, (ctypes.c_uint8 * 64)(), (ctypes.c_uint8 * 64)()))
    # workspace queries
    cap = lib.b2j_default_out_cap(ctypes.byref(d))
    assert cap >= 623 + 2 and cap % 16 == 0
    assert lib.b2j_encode_workspace_bytes(ctypes.byref(d), cap) > 3 * 4 * 128
    assert lib.b2j_encode_workspace_bytes(ctypes.byref(d), 10) == 0
    assert lib.b2j_encode_launch_count(ctypes.byref(d)) == 3


def test_operator_surface_matches_the_reference(lj):
    # names, constructor kwargs, layouts and latencies copied from litejpeg/core
    assert lj.RGB2YCbCr(rgb_w=8, ycbcr_w=8, coef_w=8).latency == 8
    assert lj.RGB2YCbCr().sink_layout == [("r", 8), ("g", 8), ("b", 8)]
    assert lj.RGB2YCbCr().source_layout == [("y", 8), ("cb", 8), ("cr", 8)]
    assert lj.YCbCr2RGB(ycbcr_w=8, rgb_w=8, coef_w=8).sink_layout == [("y", 8), ("cb", 8), ("cr", 8)]
    assert lj.YCbCr422to444(dw=8).latency == 2 and lj.ycbcr2rgb_coefs(8, 8)["dcoef"] == 98
    assert lj.YCbCr444to422(dw=8).latency == 3 and lj.YCbCr444to422().source_layout == [("y", 8), ("cb_cr", 8)]
    assert lj.DCT(dw=12, dct_block=64).latency == 8 and lj.DCT().sink_layout == [("data", 12)]
    assert lj.Quantization().table == lj.quant_values and lj.quant_values[:3] == [16, 11, 10]
    assert lj.RunLength().source_layout == [("data", 17)] and lj.RunLength().latency == 3
    assert lj.EntropyCoder().source_layout == [("data", 4)]
    assert lj.RLEMain().source_layout == [("data", 21)]
    assert lj.zigzag_rom[lj.zigzag_order[5]] == 5
    assert lj.rgb2ycbcr_coefs(8, 8) == {"ca": 46, "cb": 15, "cc": 166, "cd": 141, "yoffset": 16, "coffset": 128,
                                        "ymax": 255, "cmax": 255, "ymin": 0, "cmin": 0}
    assert lj.to_signed(lj.to_unsigned([-415, 35, 0, -1]), 12) == [-415, 35, 0, -1]


def test_product_does_not_import_the_oracle():
    # the product package must not route through oracle/ (or any CPU fallback)
    root = "jpeg_encoder_gsoc2017_ishan-bansal_b200"
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle.", "oracle.").lower() or "import oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f


def test_no_cuda_means_a_loud_error(lj):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        lj.JpegEncoder("420", 75).encode(torch.zeros((16, 16, 3), dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        lj.DCT().process([0] * 64)


def test_header_is_valid_c99_and_host_entry_points_work(tmp_path):
    """include/b2j.h compiles as strict C99 (the language a cgo / JNI / N-API stub is written in) and a
    plain C client linked against libb2j.so can use every host-only entry point without a GPU."""
    import os
    import subprocess
    import litejpeg_b200 as lj
    lj.lib()                                            # builds the library if needed
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "jpeg_encoder_gsoc2017_ishan-bansal_b200")
    exe = str(tmp_path / "abi_host_only")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "c", "abi_host_only.c"), "-o", exe, "-L", libdir, "-l:libb2j.so", "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "120x68" in run.stdout
