"""Multi-GPU plumbing on the Python side: query reads shard across ranks, index + reads replicated.
The stage's own exchange (one all-gather per EC round + one of the final lists) lives in csrc/stage.cu behind the hb_allgather_fn callback
(engine.torch_allgather = NCCL); what is here: the shard arithmetic, timing / totals agreement, and cal_ec_r_sharded — the same round
composed from the step calls of the C-ABI, kept as the step-level check of the exchange (tests, tools/ec_sharded_check.py)."""
from __future__ import annotations


def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous shard of read ids for `rank` (the last ranks may be short)."""
    per = (n_reads + world - 1) // world
    return min(n_reads, rank * per), min(n_reads, (rank + 1) * per)


def reduce_time_and_units(dev_ms: float, units: float, device=None):
    """-> (max over ranks of dev_ms, sum over ranks of units)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dev_ms, units
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def agree_max(v: int, device=None) -> int:
    """MAX of one small integer over the ranks (v itself without a process group): how the ranks agree on an error before a collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(v)
    t = torch.tensor([int(v)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def merge_shards(parts):
    """Concatenate per-rank (records, offsets) results (each offsets array is
    relative to its shard) into one (records, offsets) pair in read-id order."""
    import numpy as np
    recs = [p[0] for p in parts]
    offs = [np.zeros(1, np.uint64)]
    base = 0
    for r, o in parts:
        offs.append(o[1:].astype(np.uint64) + np.uint64(base))
        base += int(o[-1])
    return np.concatenate(recs), np.concatenate(offs)


# ---- the EC rounds on N GPUs -------------------------------------------------------------------------------------------------
# Within a round every query read is independent (rows a8-a15 write slot i only), but the round ENDS with state every read of the
# next round needs: the corrected reads (every read is somebody's target) and both overlap lists (the next round's exact shortcut,
# the final pass).  On one node the reference gets this from shared memory; here it is the one real exchange step of the stage
# ("between passes: all-gather of corrected packed reads (or of edit scripts)"): each rank all-gathers its shard's
# edit scripts and list records — a few bytes per read, not the reads — and applies ALL scripts to its replica of the store.
def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_bytes(buf, device=None):
    """all-gather of one variable-length byte string per rank -> list of uint8 arrays in rank order (NCCL on `device`, gloo on the CPU)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world = _world()
    b = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    if world == 1:
        return [b]
    n = torch.tensor([b.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s[0]) for s in sizes]; m = max(max(sizes), 1)
    mine = torch.zeros(m, dtype=torch.uint8, device=device)
    if b.size:
        mine[:b.size] = torch.from_numpy(b.copy()).to(mine.device)
    parts = [torch.zeros(m, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(parts, mine)
    return [p[:s].cpu().numpy() for p, s in zip(parts, sizes)]


def all_gather_ragged(rec, off, device=None):
    """every rank holds (records, offsets) of its contiguous read shard (shard_range order; offsets relative to the shard)
    -> (records, offsets) of ALL reads, identical on every rank"""
    import numpy as np
    rec = np.ascontiguousarray(rec); off = np.ascontiguousarray(off, dtype=np.uint64)
    recs = [p.view(rec.dtype) for p in all_gather_bytes(rec[:int(off[-1])], device)]
    offs = [p.view(np.uint64) for p in all_gather_bytes(off, device)]
    return merge_shards(list(zip(recs, offs)))


def cal_ec_r_sharded(eng, round_, is_sv, prev_src, prev_src_off, e_rate=0.04, w_l=775, device=None):
    """cal_ec_r (ecovlp.cpp:6268) with the round's work sharded over the ranks: rows a8-a15 of this rank's reads (hb_ec_round on
    [r0, r1)), the all-gather of edit scripts / lists / flags, then the closing steps a16-a18 on the rank's replica (hb_ec_stage_scc +
    hb_ec_apply, hb_ec_update_paf, hb_ec_post_rev) — every rank leaves with the same reads and the same lists as the single-GPU
    hb_cal_ec_r.  Same result dict as Engine.cal_ec_r."""
    import numpy as np
    rank, world = _world()
    n = eng.n_reads
    r0, r1 = shard_range(n, rank, world)
    eng.ec_stage_prev(prev_src, prev_src_off)
    ln = eng.read_lengths()
    tot_b = int(ln.sum())                                                            # cnt[0]: bases of the reads as they enter the round (ecovlp.cpp:3276)
    # one pass with these capacities (no sizing pass).  A rank whose shard overflows them must not leave the others waiting in the all-gather: every rank
    # reports its status, the worst one is agreed on (MAX of -code), and either all ranks repeat the round with four times the room or all raise together.
    from .engine import HBError
    cap = 2 * int(np.asarray(prev_src).size) + 256 * (r1 - r0) + 1024
    scc_cap = 256 * (r1 - r0) + int(ln[r0:r1].sum()) // 8 + 1024
    for attempt in range(4):
        r, err = None, None
        try:
            r = eng.ec_round(r0, r1, 0.02, e_rate, w_l, use_prev=1, caps=(cap, cap, scc_cap))
        except HBError as e:
            err = e
        worst = agree_max(-(err.code or -1) if err is not None else 0, device)
        if worst == 0:
            break
        if worst == 5 and attempt < 3:                                                # HB_E_OVERFLOW on some rank: all ranks retry with more room
            cap, scc_cap = cap * 4, scc_cap * 4
            continue
        raise err if err is not None else HBError("another rank failed the EC round (code %d)" % -worst)
    src, soff = all_gather_ragged(r["src"], r["src_off"], device)
    rev, roff = all_gather_ragged(r["rev"], r["rev_off"], device)
    scc, scc_off = all_gather_ragged(r["scc"], r["scc_off"], device)
    flags = np.concatenate([p for p in all_gather_bytes(np.stack([r["is_fully_corrected"], r["is_abnormal"], r["status"]], axis=1) if r1 > r0 else np.zeros((0, 3), np.uint8), device)]).reshape(-1, 3)
    _, tot_e = reduce_time_and_units(0.0, float(r["n_corrected"]), device=device)
    eng.ec_stage_scc(scc, scc_off)
    eng.ec_apply()
    upd, n_ex, n_inex = eng.ec_update_paf(src, soff)
    if not is_sv or (round_ & 1):
        upd, soff, rev, roff = eng.ec_post_rev(upd, soff, rev, roff)
    return dict(src=upd, src_off=soff, rev=rev, rev_off=roff, is_fully_corrected=flags[:, 0].copy(), is_abnormal=flags[:, 1].copy(), status=flags[:, 2].copy(),
                tot_b=tot_b, tot_e=int(tot_e), n_exact=int(n_ex), n_inexact=int(n_inex))
