#!/usr/bin/env python
"""The whole stage on N GPUs of one node with the work of every pass sharded over the ranks:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/ec_sharded_check.py [g1|g2|g3|g4]
Every rank holds the reads and the index (replicated); an EC round runs rows a8-a15 on the rank's shard, all-gathers edit scripts / lists
(hifiasm_b200.dist.cal_ec_r_sharded, the one exchange step of the stage), closes the round on its replica; the final pass runs on the shard
and the lists are all-gathered.  Every rank must end with the reference's corrected reads and final lists (golden digests).  Prints OK / FAIL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hifiasm_b200  # noqa: E402
from hifiasm_b200 import binio, dist as hdist  # noqa: E402
from goldenlib import Golden  # noqa: E402
import roundlib  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "g1"
    local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = Golden(name); rd = roundlib.Rounds(name)
    eng = hifiasm_b200.Engine(local)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    n = g.raw.n
    src = np.zeros(0, binio.MA_MEM); soff = np.zeros(n + 1, np.uint64); ok = True
    for K in range(3):
        a, b = eng.pt_gen(); eng.set_opt(hom_cov=a, het_cov=b)
        r = hdist.cal_ec_r_sharded(eng, K, 1 if K == 2 else 0, src, soff, device=dev)
        src, soff, rev, roff = r["src"], r["src_off"], r["rev"], r["rev_off"]
        ok &= not r["status"].any() and r["tot_e"] == int(rd.params(K)["tot_e"])
        ok &= bool((roundlib.list_digests(src, soff, 0) == rd.digest(K, "post_src")).all() and (roundlib.list_digests(rev, roff, 1) == rd.digest(K, "post_rev")).all())
    ok &= bool((roundlib.reads_digests(eng.download_reads()) == roundlib.reads_digests(g.pre)).all())
    a, b = eng.pt_gen(); eng.set_opt(hom_cov=a, het_cov=b)
    r0, r1 = hdist.shard_range(n, dist.get_rank() if world > 1 else 0, world)
    o0, q0, o1, q1, _ = eng.cal_ov_r(src, soff, rev, roff, r0, r1)
    f0, fo0 = hdist.all_gather_ragged(o0, q0, dev); f1, fo1 = hdist.all_gather_ragged(o1, q1, dev)
    w0, wo0, _, _ = g.fin_src; w1, wo1, _, _ = g.fin_rev
    ok &= bool((fo0 == wo0).all() and (fo1 == wo1).all() and all((f0[f] == w0[f]).all() and (f1[f] == w1[f]).all() for f in binio.MA_DISK.names))
    print("rank %d/%d %s: %s" % (dist.get_rank() if world > 1 else 0, world, name, "OK" if ok else "FAIL"), flush=True)
    eng.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
