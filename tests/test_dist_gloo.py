"""world_size-2 `gloo` test (CPU) of the multi-GPU plumbing: contiguous query
sharding covers every read exactly once, the timing reduction is the max over
ranks, and per-rank overlap lists re-assemble in read-id order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from hifiasm_b200 import dist as hdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_reads, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r0, r1 = hdist.shard_range(n_reads, rank, world)
    t, u = hdist.reduce_time_and_units(10.0 * (rank + 1), float(r1 - r0))
    # a fake per-rank overlap list: read i has (i % 3) records holding its id
    cnt = np.array([i % 3 for i in range(r0, r1)], dtype=np.uint64)
    off = np.zeros(r1 - r0 + 1, np.uint64); np.cumsum(cnt, out=off[1:])
    rec = np.repeat(np.arange(r0, r1, dtype=np.uint64), cnt.astype(np.int64))
    q.put((rank, r0, r1, t, u, rec, off))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_reduce_world2():
    world, n_reads = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_reads, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_reads  # contiguous cover
    for r in res:
        assert r[3] == 20.0 and r[4] == float(n_reads)  # max time, summed units
    rec, off = hdist.merge_shards([(r[5], r[6]) for r in res])
    assert off.size == n_reads + 1 and int(off[-1]) == rec.size
    for i in (0, 1, 500, 501, 1000):
        assert (rec[int(off[i]):int(off[i + 1])] == i).all() and int(off[i + 1] - off[i]) == i % 3


def test_shard_range_edges():
    assert hdist.shard_range(10, 0, 4) == (0, 3) and hdist.shard_range(10, 3, 4) == (9, 10)
    assert hdist.shard_range(2, 3, 4) == (2, 2)  # empty shard
    tot = sum(b - a for a, b in (hdist.shard_range(12345, r, 8) for r in range(8)))
    assert tot == 12345
