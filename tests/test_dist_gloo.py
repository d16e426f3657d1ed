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


# ---- the exchange step that ends an EC round on N GPUs (hifiasm_b200.dist.all_gather_ragged / cal_ec_r_sharded) ----
def _worker_ragged(rank, world, port, n_reads, q):
    from hifiasm_b200 import binio
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r0, r1 = hdist.shard_range(n_reads, rank, world)
    # overlap records (48-byte ma_hit_t mirror) and edit scripts (uint16 runs) of the shard: read i has i % 4 records and (7 * i) % 5 runs
    cnt = np.array([i % 4 for i in range(r0, r1)], dtype=np.uint64); off = np.zeros(r1 - r0 + 1, np.uint64); np.cumsum(cnt, out=off[1:])
    rec = np.zeros(int(off[-1]), binio.MA_MEM); rec["qns"] = np.repeat(np.arange(r0, r1, dtype=np.uint64), cnt.astype(np.int64)) << np.uint64(32); rec["tn"] = np.arange(rec.size) + 1000 * rank
    c2 = np.array([(7 * i) % 5 for i in range(r0, r1)], dtype=np.uint64); o2 = np.zeros(r1 - r0 + 1, np.uint64); np.cumsum(c2, out=o2[1:])
    sc = (np.repeat(np.arange(r0, r1, dtype=np.uint64), c2.astype(np.int64)) & np.uint64(0xffff)).astype(np.uint16)
    big = np.zeros(int(off[-1]) + 7, binio.MA_MEM); big[:rec.size] = rec   # buffers may be larger than what the offsets cover
    a_rec, a_off = hdist.all_gather_ragged(big, off)
    s_rec, s_off = hdist.all_gather_ragged(sc, o2)
    flags = np.concatenate(hdist.all_gather_bytes(np.full((r1 - r0, 3), rank, np.uint8))).reshape(-1, 3)
    q.put((rank, a_rec.tobytes(), a_off, s_rec.tobytes(), s_off, flags))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_ragged_world2():
    from hifiasm_b200 import binio
    world, n_reads = 2, 203
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker_ragged, args=(r, world, port, n_reads, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank ends with the same global arrays, in read-id order
    assert res[0][1] == res[1][1] and (res[0][2] == res[1][2]).all() and res[0][3] == res[1][3] and (res[0][4] == res[1][4]).all() and (res[0][5] == res[1][5]).all()
    rec = np.frombuffer(res[0][1], binio.MA_MEM); off = res[0][2]; sc = np.frombuffer(res[0][3], np.uint16); so = res[0][4]
    assert off.size == n_reads + 1 and int(off[-1]) == rec.size and so.size == n_reads + 1 and int(so[-1]) == sc.size
    for i in (0, 1, 50, 101, 102, 150, 202):
        assert int(off[i + 1] - off[i]) == i % 4 and ((rec["qns"][int(off[i]):int(off[i + 1])] >> np.uint64(32)) == i).all()
        assert int(so[i + 1] - so[i]) == (7 * i) % 5 and (sc[int(so[i]):int(so[i + 1])] == i).all()
    r0, r1 = hdist.shard_range(n_reads, 1, world)
    assert (res[0][5][:r0] == 0).all() and (res[0][5][r0:] == 1).all()


def test_single_process_is_the_identity():
    from hifiasm_b200 import binio
    rec = np.zeros(5, binio.MA_MEM); rec["tn"] = np.arange(5); off = np.array([0, 2, 2, 5], np.uint64)
    a, o = hdist.all_gather_ragged(rec, off)
    assert a.tobytes() == rec.tobytes() and (o == off).all()


# ---- the transport hb_stage_run calls for its exchange (hifiasm_b200.engine.torch_allgather: an hb_allgather_fn over torch.distributed) ----
def _cb_worker(rank, world, port, q):
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hifiasm_b200.engine import torch_allgather
    cb = torch_allgather(None)
    nbytes = 4096 + 16
    send = (C.c_uint8 * nbytes)(*([(rank * 37 + i) & 0xff for i in range(nbytes)])); recv = (C.c_uint8 * (nbytes * world))()
    rc = cb(None, C.addressof(send), C.addressof(recv), nbytes)
    q.put((rank, rc, bytes(recv)))
    dist.barrier(); dist.destroy_process_group()


def test_allgather_callback_world2():
    world = 2; ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_cb_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60); assert p.exitcode == 0
    nbytes = 4096 + 16
    want = b"".join(bytes([(r * 37 + i) & 0xff for i in range(nbytes)]) for r in range(world))
    for rank, rc, got in res:
        assert rc == 0 and got == want   # every rank holds every rank's bytes, in rank order


# ---- a failing rank must not leave the others waiting in the exchange: cal_ec_r_sharded agrees on the worst status first (round-1 advice) ----
class _FakeEngine:
    """Engine stand-in for the CPU: rank 1's first ec_round call overflows its buffers (HB_E_OVERFLOW = -5), every later call succeeds"""
    def __init__(self, rank, n):
        self.rank, self.n_reads, self.calls, self.caps = rank, n, 0, []

    def ec_stage_prev(self, *a): pass
    def read_lengths(self): return np.full(self.n_reads, 1000, np.uint32)
    def ec_stage_scc(self, *a): pass
    def ec_apply(self): pass
    def ec_update_paf(self, src, soff): return src, 0, 0
    def ec_post_rev(self, upd, soff, rev, roff): return upd, soff, rev, roff

    def ec_round(self, r0, r1, bw, e_rate, w_l, use_prev=1, caps=None):
        from hifiasm_b200.engine import HBError
        self.calls += 1; self.caps.append(caps[0])
        if self.rank == 1 and self.calls == 1:
            e = HBError("overlap output capacity"); e.code = -5; raise e
        m = r1 - r0; z = np.zeros(0, hdist_ma()); o = np.zeros(m + 1, np.uint64)
        return dict(src=z, src_off=o, rev=z.copy(), rev_off=o.copy(), scc=np.zeros(0, np.uint16), scc_off=o.copy(), is_fully_corrected=np.zeros(m, np.uint8),
                    is_abnormal=np.zeros(m, np.uint8), status=np.zeros(m, np.uint8), n_corrected=0)


def hdist_ma():
    from hifiasm_b200 import binio
    return binio.MA_MEM


def _worker_retry(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(rank, 10)
    r = hdist.cal_ec_r_sharded(eng, 0, 0, np.zeros(0, hdist_ma()), np.zeros(11, np.uint64))
    q.put((rank, eng.calls, eng.caps, int(r["src_off"].size)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_round_retries_on_every_rank_when_one_overflows():
    world = 2
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker_retry, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, caps, n_off in res:
        assert calls == 2 and caps[1] == 4 * caps[0], (rank, calls, caps)     # both ranks repeated the round, with four times the room
        assert n_off == 11                                                      # and left with all reads' (empty) lists
