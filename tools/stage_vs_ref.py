#!/usr/bin/env python
"""Whole stage on N GPUs against the unmodified reference binary, files compared byte for byte (GPU box; single process or under torchrun):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/stage_vs_ref.py [GENOME_MB [SEED]]
Every rank runs stage.run_stage on the same FASTA (query reads sharded, one all-gather per EC round); rank 0 writes the files, runs the reference and compares."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import simgen
    from hifiasm_b200 import stage
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0; seed = int(sys.argv[2]) if len(sys.argv) > 2 else 77
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    td = os.environ.get("HB_STAGE_TMP") or tempfile.mkdtemp(prefix="hb_stage_")
    if world > 1:  # one directory for all ranks
        import torch.distributed as dist
        obj = [td]; dist.broadcast_object_list(obj, src=0); td = obj[0]
    fa = os.path.join(td, "reads.fa")
    if rank == 0:
        simgen.make(mb, 30, seed=seed, n_rate=0.0002, fasta=fa)
    if world > 1:
        dist.barrier()
    t = time.time()
    info = stage.run_stage(fa, os.path.join(td, "gpu"), device=local)
    dt = time.time() - t
    rc = 0
    if rank == 0:
        import test_gpu_scale as tg
        ref = os.path.join(ROOT, "oracle", "_ref", "hifiasm")
        p = subprocess.run([ref, "-o", os.path.join(td, "ref"), "-t%d" % len(os.sched_getaffinity(0)), "-f0", "--write-paf", "--write-ec", fa], capture_output=True, text=True)
        diff = tg._cmp_files(td, "ref", "gpu") if p.returncode == 0 else ["the reference binary failed"]
        print("stage on %d GPU(s), %g Mb genome (seed %d): %d reads, %d bases, run_stage %.1f s (stage %.1f s, exchange %.2f s), corrected per round %s, overlaps %d + %d: %s"
              % (world, mb, seed, info["reads"], info["bases"], dt, info["ms"]["total"] / 1e3, info["ms"]["exchange"] / 1e3, info["corrected_bases"], info["overlaps_src"], info["overlaps_rev"],
                 "IDENTICAL to the reference binary's files" if not diff else "DIFFERENT: %s" % diff), flush=True)
        rc = 1 if diff else 0
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
