"""Integer identities the CUDA kernels rely on, re-derived on the CPU from the constants
in csrc/b2j_math.cuh and checked exhaustively against the oracle (which is pinned to the
reference model).  CPU only; the kernels themselves are checked by the -m gpu tests."""
import re

import numpy as np

from oracle import model_np as M

SRC = open("jpeg_encoder_gsoc2017_ishan-bansal_b200/csrc/b2j_math.cuh").read()


def const(name):
    m = re.search(r"#define\s+%s\s+(0x[0-9A-Fa-f]+|\d+)u" % name, SRC)
    assert m, name
    return int(m.group(1), 0)


def all_rgb():
    allv = np.arange(1 << 24, dtype=np.int64)
    return allv >> 16, (allv >> 8) & 255, allv & 255


def cr_exception_rule(r, g):
    """Python port of cr_exception() in b2j_math.cuh."""
    full = np.isin(g, [201, 213, 225, 237, 249])
    part = np.isin(g, [164, 176, 188]) & ((r + 128 <= g) | (r + 30 >= g))
    return (g >= 164) & (r <= g) & (full | part)


def test_fixed_point_colour_is_the_model_for_all_inputs():
    R, G, B = all_rgb()
    Y, CB, CR = M.rgb2ycbcr(R, G, B)
    mask = const("B2J_CONST_EXC_MASK")
    m32 = (1 << 32) - 1
    sy = (R * const("B2J_CONST_Y_R") + G * const("B2J_CONST_Y_G") + B * const("B2J_CONST_Y_B"))
    assert sy.max() <= m32                                  # no 32-bit overflow
    ycand = (sy & mask) == 0
    assert np.array_equal((sy >> 24)[~ycand], Y[~ycand])    # candidates are re-evaluated in FP64 by the kernel
    assert ycand.sum() == 16774
    scb = (B * const("B2J_CONST_CX_HALF") - R * const("B2J_CONST_CB_R") - G * const("B2J_CONST_CB_G") + const("B2J_CONST_C_OFFSET")) & m32
    assert np.array_equal(scb >> 24, CB)
    scr = (R * const("B2J_CONST_CX_HALF") - G * const("B2J_CONST_CR_G") - B * const("B2J_CONST_CR_B") + const("B2J_CONST_C_OFFSET")) & m32
    ccand = (scr & mask) == 0
    assert np.array_equal(ccand, (G == B) & ((R - G) % 2 == 0))
    exc = (scr >> 24) != CR
    assert exc.sum() == 688 and not (exc & ~ccand).any()
    assert np.array_equal(((scr >> 24) - exc)[ccand], CR[ccand])
    # the closed-form exception rule used on the device lists exactly those 688 inputs
    assert np.array_equal(cr_exception_rule(R[ccand], G[ccand]), exc[ccand])


def test_quant_multiplier_identity():
    # q = (a * ceil(2^20 / d) + 2^19) >> 20 == round-half-away(a / d) for a <= 1024 (fused kernel),
    # and equals the model's quantize (oracle.model_np.quantize is pinned exhaustively in make_golden.py)
    a = np.arange(0, 1025, dtype=np.int64)
    for d in range(1, 256):
        mul = ((1 << 20) + d - 1) // d
        assert (a * mul + (1 << 19)).max() < (1 << 32)
        assert np.array_equal((a * mul + (1 << 19)) >> 20, M.quantize(a, np.full_like(a, d)))


def test_dct_output_range_fits_the_fused_quantiser():
    # |int(v/8)| <= 1024 for any 8-bit block, so the 2^20 multiplier identity applies
    rng = np.random.default_rng(0)
    blocks = [np.zeros(64, int), np.full(64, 255)]
    for _ in range(200):
        blocks.append(rng.choice([0, 255], 64))
    for u in range(8):
        for v in range(8):   # sign patterns of every basis function
            blk = np.array([255 if np.cos((2 * (i // 8) + 1) * v * np.pi / 16) * np.cos((2 * (i % 8) + 1) * u * np.pi / 16) > 0 else 0
                            for i in range(64)])
            blocks += [blk, 255 - blk]
    out = M.dct(np.array(blocks))
    assert np.abs(out).max() <= 1024


def test_magic_number_conversions():
    # i2d: (2^52 + 2^31 + x) - (2^52 + 2^31) == x exactly; abs_div8_trunc: low word of RZ(|v| + 2^55)
    import struct
    for x in (-2 ** 31, -2049, -1, 0, 1, 255, 2040, 2 ** 31 - 1):
        bits = (0x43300000 << 32) | ((x ^ 0x80000000) & 0xFFFFFFFF)
        d = struct.unpack("<d", struct.pack("<Q", bits))[0]
        assert d - 4503601774854144.0 == float(x)
    # the RZ add is exact here whenever |v| is a multiple of 8 or we truncate by hand:
    for v in (0.0, 7.999, 8.0, -15.99, 8191.5, -8192.0, 1e-300):
        t = 36028797018963968.0 + float(int(abs(v) / 8)) * 8        # what RZ(|v| + 2^55) evaluates to
        lo = struct.unpack("<Q", struct.pack("<d", t))[0] & 0xFFFFFFFF
        assert lo == int(abs(v) / 8)


def _entropy_walk_permutation(counts, nown, warps=8):
    """NumPy restatement of the block -> thread assignment of k_entropy_fused (csrc/b2j_entropy_fused.cu, "permutation"):
    warp w of the CTA has loaded the pieces of blocks (256 / 8 or 128 / 8) * a + 4 w + k, a = 0..7, k = 0..3 (lane = 4 a + k);
    the blocks of one residue modulo 32 are ranked by non-zero count (ties: lower lane first, absent blocks last) and the
    block of rank r goes to thread 32 (warps - 1 - r) + residue.  Returns perm[thread] = block."""
    chunk = 32 * warps
    step = 32 // warps                       # lane distance between two blocks of one group
    perm = np.full(chunk, -1, dtype=np.int64)
    for w in range(warps):
        lanes = np.arange(32)
        blk = (lanes >> 2) * (chunk // 8) + w * 4 + (lanes & 3)
        key = np.where(blk < nown, 1 + counts[np.minimum(blk, chunk - 1)], 0) * 32 + (31 - lanes)
        for lane in range(32):
            r = sum(int(key[lane ^ (step * d)] > key[lane]) for d in range(1, warps))
            t = (warps - 1 - r) * 32 + (blk[lane] & 31)
            assert perm[t] == -1                      # one block per thread
            perm[t] = blk[lane]
    return perm


def test_entropy_walk_permutation_properties():
    """What the kernel's comments claim about the permutation: it is a bijection of the chunk's columns, the walking
    thread's lane is the block's index modulo 32 (same shared-memory bank pattern as thread t on block t), warp 7 gets
    the longest and warp 0 the shortest block of every residue class, absent blocks of a short chunk go to the low
    warps, and the busiest-lane sum over the warps never exceeds that of consecutive blocks."""
    rng = np.random.default_rng(20171019)
    for warps in (8, 4):
        chunk = 32 * warps
        for trial in range(200):
            kind = trial % 4
            if kind == 0:
                counts = rng.integers(0, 64, chunk)
            elif kind == 1:
                counts = rng.integers(0, 4, chunk)                      # many ties
            elif kind == 2:
                counts = np.full(chunk, int(rng.integers(0, 64)))       # all equal (grey frames)
            else:
                counts = np.minimum(63, rng.geometric(0.12, chunk))     # photo-like: long tail
            nown = chunk if trial % 3 else int(rng.integers(1, chunk + 1))
            perm = _entropy_walk_permutation(counts, nown, warps)
            assert sorted(perm.tolist()) == list(range(chunk))
            t = np.arange(chunk)
            assert np.array_equal(perm & 31, t & 31)
            c = np.where(perm < nown, counts[perm], -1).reshape(warps, 32)  # [warp][lane], -1 = absent
            assert np.all(np.diff(c, axis=0) >= 0)                      # per residue: counts ascend with the warp index
            if nown == chunk:
                new = c.max(axis=1).sum()
                old = counts.reshape(warps, 32).max(axis=1).sum()
                assert c[warps - 1].max() == counts.max() and new <= old
