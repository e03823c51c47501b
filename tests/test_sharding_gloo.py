"""Multi-GPU host logic on CPU: frame sharding and the output-only gathers, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_partition_the_batch(lj):
    for n in (1, 2, 7, 16, 64, 256, 257):
        for world in (1, 2, 3, 4, 8):
            ranges = [lj.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert [lj.shard_range(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]
    with pytest.raises(ValueError):
        lj.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import sys
    sys.path.insert(0, os.getcwd())
    import litejpeg_b200 as lj
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = lj.shard_range(n_frames, world, rank)
    # stand-ins for what each rank's GPU would produce: frame i -> (i+1)*3 bytes of value i
    streams = [bytes([i % 251]) * ((i + 1) * 3) for i in range(lo, hi)]
    sizes = torch.tensor([len(s) for s in streams], dtype=torch.int64)
    all_sizes = lj.gather_sizes(sizes, n_frames)
    gathered = lj.gather_streams(streams, n_frames, dst=0)
    ok = all_sizes.tolist() == [(i + 1) * 3 for i in range(n_frames)]
    if rank == 0:
        ok = ok and gathered == [bytes([i % 251]) * ((i + 1) * 3) for i in range(n_frames)]
    else:
        ok = ok and gathered is None
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_gathers_world_size_2_gloo(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


# ---------------------------------------------------------------- one frame split into bands
def _test_frame(h, w):
    rng = np.random.default_rng(11)
    return np.clip(np.add.outer(np.arange(h) * 2, np.arange(w))[..., None] % 256 + rng.integers(-25, 25, (h, w, 3)), 0, 255).astype(np.uint8)


class _OracleBandEncoder:
    """CPU stand-in for JpegEncoder.band_encoder(): the oracle encodes the band (restart markers
    counted from 0, so the tests below use bands of 8 MCU rows: phase 0 everywhere)."""

    def __init__(self, sub, ri):
        self.sub, self.ri = sub, ri

    def encode(self, band):
        from oracle import jpeg_ext as J
        return [J.encode_jpeg(np.asarray(band[0]), self.sub, 75, "standard", self.ri)]


def test_band_plan_and_assembly_equal_the_whole_frame(lj):
    """[ext] band_plan / assemble_bands (host logic of the one-frame-on-several-GPUs path): bands
    encoded independently and joined are byte-identical to the whole frame encoded with one
    restart interval per MCU row."""
    from oracle import jpeg_ext as J
    assert lj.band_plan(1080, "420", 8)[0] == (0, 128, 0) and lj.band_plan(1080, "420", 8)[-1] == (944, 1080, 59)
    assert lj.band_plan(16, "420", 4) == [(0, 16, 0)]                 # never more bands than MCU rows
    for sub, h, w in (("444", 128, 72), ("420", 250, 40), ("422", 190, 56)):
        img = _test_frame(h, w)
        mcus_x = J.geometry(h, w, sub)["mcus_x"]
        whole = J.encode_jpeg(img, sub, 75, "standard", mcus_x)
        mh = 16 if sub == "420" else 8
        rows = (h + mh - 1) // mh
        plan = [(r0 * mh, min((r0 + 8) * mh, h), r0) for r0 in range(0, rows, 8)]     # 8 MCU rows per band: phase 0
        files = [_OracleBandEncoder(sub, mcus_x).encode(img[None, y0:y1])[0] for (y0, y1, _) in plan]
        assert lj.assemble_bands(files, plan, h) == whole, sub
    with pytest.raises(ValueError):
        lj.assemble_bands([b"x"], [(0, 8, 0), (8, 16, 1)], 16)


def _split_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.getcwd())
    import litejpeg_b200 as lj
    from oracle import jpeg_ext as J
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w, sub = 256, 48, "420"                                        # 16 MCU rows -> 2 bands of 8 rows
    img = _test_frame(h, w)
    mcus_x = J.geometry(h, w, sub)["mcus_x"]
    enc = lj.ShardedJpegEncoder(world, rank, subsampling=sub, quality=75, zrl="standard")
    enc.encoder.band_encoder = lambda width, before: _OracleBandEncoder(sub, mcus_x)    # no GPU in this test
    out = enc.encode_frame_split(torch.from_numpy(img), dst=0)
    ok = (out == J.encode_jpeg(img, sub, 75, "standard", mcus_x)) if rank == 0 else out is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_split_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
