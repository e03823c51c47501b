"""Multi-GPU host logic on CPU: frame sharding and the output-only gathers, world_size 2, gloo."""
import os
import socket

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
