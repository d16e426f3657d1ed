#!/usr/bin/env python
"""Whole overlap + EC stage on one GPU with a per-call wall clock and the per-kernel device times of every pass, next to the unmodified reference
binary on the host cores (same reads, from FASTA).  Development tool for the GPU box:
    python tools/profile_stage.py GENOME_MB [--ref] [--cov 30] [--out gpurun_out/stage_profile.json]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import simgen  # noqa: E402


def ref_stage(fa, threads, extra=()):
    """the unmodified reference binary: -> dict(stage_s = the wall stamp of 'found overlaps for the final round', ft0_s = stamp of the first ha_ft_gen line)"""
    ref = os.path.join(ROOT, "oracle", "_ref", "hifiasm")
    with tempfile.TemporaryDirectory() as td:
        t = time.time()
        p = subprocess.run([ref, "-o", os.path.join(td, "ref"), "-t%d" % threads, "-f0", "--bin-only", *extra, fa], capture_output=True, text=True)
        wall = time.time() - t
    out = {"wall_s": wall, "rc": p.returncode, "threads": threads}
    m = re.search(r"\[M::ha_assemble::([0-9.]+)\*[0-9.]+@[0-9.]+GB\] ==> found overlaps for the final round", p.stderr)
    if m:
        out["stage_s"] = float(m.group(1))
    m = re.search(r"\[M::ha_ft_gen::([0-9.]+)\*", p.stderr)
    if m:
        out["ft0_s"] = float(m.group(1))
    out["pec_s"] = [float(x) for x in re.findall(r"\[M::pec::([0-9.]+)\]", p.stderr)]
    m = re.search(r"# running time: ([0-9.]+)", p.stderr)
    if m:
        out["final_s"] = float(m.group(1))
    m = re.search(r"# bases: (\d+)", p.stderr)
    if m:
        out["bases"] = int(m.group(1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("genome_mb", type=float)
    ap.add_argument("--cov", type=float, default=30)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    res = {"genome_mb": a.genome_mb, "cov": a.cov}
    t = time.time()
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "reads.fa") if a.ref else None
        rs = simgen.make(a.genome_mb, a.cov, fasta=fa)
        res.update(reads=rs.n, bases=rs.bases, gen_s=round(time.time() - t, 1))
        if not a.no_gpu:
            import hifiasm_b200
            from hifiasm_b200 import binio
            eng = hifiasm_b200.Engine(0)
            calls = []; kms = {}

            def timed(name, f):
                eng.profile_reset(); t0 = time.time(); r = f(); dt = time.time() - t0
                pr = eng.profile()
                calls.append((name, round(dt * 1e3, 1), round(sum(v[1] for v in pr.values()), 1)))   # wall ms, kernel ms inside the call
                for k, v in pr.items():
                    e = kms.setdefault(k, [0, 0.0]); e[0] += v[0]; e[1] += v[1]
                return r
            n = rs.n
            tw = time.time()
            timed("upload", lambda: eng.upload_reads(rs.length, rs.flat, rs.byte_off, rs.n_pos, rs.n_off))
            hom = timed("ft_gen", eng.ft_gen); eng.update_cov(hom)
            src = np.zeros(0, binio.MA_MEM); soff = np.zeros(n + 1, np.uint64); rev = src.copy(); roff = soff.copy(); tot_e = []
            for k in range(3):
                h, t2 = timed("pt_gen%d" % k, eng.pt_gen); eng.set_opt(hom_cov=h, het_cov=t2)
                r = timed("cal_ec_r%d" % k, lambda: eng.cal_ec_r(k, 1 if k == 2 else 0, src, soff))
                src, soff, rev, roff = r["src"], r["src_off"], r["rev"], r["rev_off"]; tot_e.append(r["tot_e"])
                res.setdefault("unfinished", []).append(int((r["status"] != 0).sum()))
            h, t2 = timed("pt_gen3", eng.pt_gen); eng.set_opt(hom_cov=h, het_cov=t2)
            o0, q0, o1, q1, stat = timed("cal_ov_r", lambda: eng.cal_ov_r(src, soff, rev, roff))
            wall = time.time() - tw
            res.update(gpu_stage_s=wall, gpu_gbp_s=rs.bases / wall / 1e9, calls=calls, corrected=tot_e, overlaps_src=int(o0.size), overlaps_rev=int(o1.size), ft_size=eng.ft_size(),
                       kernels_ms={k: [v[0], round(v[1], 2)] for k, v in sorted(kms.items(), key=lambda kv: -kv[1][1])}, kernel_ms_total=round(sum(v[1] for v in kms.values()), 1))
            eng.close()
        if a.ref:
            res["reference"] = ref_stage(fa, a.threads)
            if "stage_s" in res["reference"]:
                res["reference"]["gbp_s"] = rs.bases / res["reference"]["stage_s"] / 1e9
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
