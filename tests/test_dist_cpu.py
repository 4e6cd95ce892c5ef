"""CPU, world_size 2 over gloo: the clip sharding and the final frame gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rib.dist import gather_frames, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.uint8).view(-1, 1, 1, 1).repeat(1, 4, 6, 3)
    out = gather_frames(local, n_total, dst=0)
    # the asynchronous form into a caller-owned buffer (what bench.py uses): same frames, same buffer
    buf = torch.full((n_total, 4, 6, 3), 255, dtype=torch.uint8) if rank == 0 else None
    pending = gather_frames(local + 100, n_total, dst=0, out=buf, async_op=True)
    out2 = pending.wait()
    if rank == 0:
        assert out2 is buf and out.shape == (n_total, 4, 6, 3)
        assert torch.equal(out2, out + 100)
        q.put(out[:, 0, 0, 0].tolist())
    else:
        assert out is None and out2 is None
    dist.destroy_process_group()


def test_gather_frames_gloo_world2_ragged():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n_total = 5                      # ragged: shards of 3 and 2 clips
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == list(range(n_total))
