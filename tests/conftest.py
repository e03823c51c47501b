import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible and `-m gpu` was not asked for."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class OracleC:
    """ctypes view of oracle/_build/libb2j_oracle.so (the plain-C CPU oracle)."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.b2jo_encode_jpeg.restype = ctypes.c_int64
        self.lib.b2jo_entropy_encode.restype = ctypes.c_int64

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def geometry(self, h, w, sub):
        from oracle import jpeg_ext as J
        return J.geometry(h, w, str(sub))

    def quant_tables(self, quality):
        qy = np.zeros(64, np.uint8)
        qc = np.zeros(64, np.uint8)
        self.lib.b2jo_quant_tables(int(quality), self._p(qy), self._p(qc))
        return qy, qc

    def dct_fixed(self, blocks):
        inp = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 64)
        out = np.zeros_like(inp)
        assert self.lib.b2jo_dct_fixed(self._p(inp), ctypes.c_int64(inp.shape[0]), self._p(out)) == 0
        return out

    def coefficients(self, rgb, sub, qy, qc, dct_fixed=False):
        rgb = np.ascontiguousarray(rgb)
        h, w = rgb.shape[:2]
        g = self.geometry(h, w, sub)
        qy = np.ascontiguousarray(qy, dtype=np.uint8)
        qc = np.ascontiguousarray(qc, dtype=np.uint8)
        coefs = np.zeros((g["n_mcu"], g["blocks_per_mcu"], 64), np.int16)
        rc = self.lib.b2jo_encode_coefficients_ex(self._p(rgb), h, w, int(sub), self._p(qy), self._p(qc), int(bool(dct_fixed)),
                                                  self._p(coefs))
        assert rc == 0
        return coefs

    def block_bits(self, coefs, h, w, sub, zrl):
        coefs = np.ascontiguousarray(coefs, dtype=np.int16)
        bits = np.zeros(coefs.size // 64, np.int32)
        rc = self.lib.b2jo_block_bits(self._p(coefs), h, w, int(sub), int(zrl), self._p(bits))
        assert rc == 0
        return bits

    def encode_jpeg(self, rgb, sub, quality, zrl):
        rgb = np.ascontiguousarray(rgb)
        h, w = rgb.shape[:2]
        cap = ((h + 15) // 16 * 16) * ((w + 15) // 16 * 16) * 6 + 4096      # padded to whole MCUs
        out = np.zeros(cap, np.uint8)
        n = self.lib.b2jo_encode_jpeg(self._p(rgb), h, w, int(sub), int(quality), int(zrl), self._p(out), ctypes.c_int64(cap))
        assert 0 < n <= cap
        return out[:n].tobytes()

    def decode_coefficients(self, jpg, h, w, sub):
        g = self.geometry(h, w, sub)
        buf = np.frombuffer(jpg, dtype=np.uint8).copy()
        coefs = np.zeros((g["n_mcu"], g["blocks_per_mcu"], 64), np.int16)
        rc = self.lib.b2jo_decode_coefficients(self._p(buf), ctypes.c_int64(buf.size), h, w, int(sub), self._p(coefs))
        return rc, coefs


@pytest.fixture(scope="session")
def oracle_c():
    path = os.path.join(ROOT, "oracle", "_build", "libb2j_oracle.so")
    src = os.path.join(ROOT, "oracle", "c", "b2j_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
    return OracleC(path)


@pytest.fixture(scope="session")
def golden():
    class G:
        lena = np.load(os.path.join(GOLDEN, "lena_model.npz"))
        stages = np.load(os.path.join(GOLDEN, "stage_vectors.npz"))
        ref = np.load(os.path.join(GOLDEN, "reference_fixtures.npz"))
        import json
        meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
        rle_first = json.load(open(os.path.join(GOLDEN, "lena_rle_first64.json")))
    return G


@pytest.fixture(scope="session")
def lj():
    import litejpeg_b200
    return litejpeg_b200
