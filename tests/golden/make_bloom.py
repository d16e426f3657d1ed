#!/usr/bin/env python
"""Regenerates tests/golden/bloom.npz: the UNMODIFIED reference's filter table (ha_ft_gen, htab.cpp:1136) built with its Bloom filter
(-f 21 / 22 / 24, and -f 0 = exact) on a read set made so that the filter matters: a 400 kb genome with a repeat family, 8x reads, and
34 copies of one extra 5 kb read appended at the END — its k-mers occur 34 times, one below the cut-off of 35 (5 x hom_cov 7), and their first occurrence
comes late, when the small filters are saturated, so they enter the table (with 35) exactly where their first occurrence is a false
positive.  `refdump ftq` probes ha_ft_cnt (htab.cpp:1064) for every distinct k-mer; stored per -f: the non-zero (hash, count) pairs and hom_cov.
Only runs in the build container (needs oracle/_ref/refdump)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from hifiasm_b200 import sim  # noqa: E402
import ha_oracle as ho  # noqa: E402

GK = dict(glen=400000, seed=3, snp_rate=0.001, repeat_frac=0.08, repeat_len=2000, repeat_div=0.002)
RK = dict(cov=8, mean_len=9000, seed=3, sd_len=2000, min_len=2000, err=0.002)
SHIFTS = (0, 21, 22, 24)


def reads():
    h1, h2 = sim.sim_genome(**GK)
    rr = sim.sim_reads(h1, h2, **RK)
    extra = np.random.default_rng(5).integers(0, 4, 5000, dtype=np.uint8)
    return rr + [extra] * 34


def main():
    refdump = os.path.join(ROOT, "oracle", "_ref", "refdump")
    if not os.path.exists(refdump):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    rr = reads()
    flat, boff, ln, npos, noff = sim.pack_reads(rr)
    st = ho.Store(ln, boff, flat, noff, npos)
    km = np.unique(ho.all_kmers(st, ho.default_opt()))
    arrs = {"n_distinct": np.array([km.size], np.uint64)}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "r.fa"); sim.write_fasta(fa, rr); km.tofile(os.path.join(td, "h.bin"))
        for sh in SHIFTS:
            out = subprocess.run([refdump, "ftq", os.path.join(td, "h.bin"), os.path.join(td, "c.bin"), "-o", os.path.join(td, "asm"), "-t4", "-f%d" % sh, fa], capture_output=True, text=True, check=True)
            r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            c = np.fromfile(os.path.join(td, "c.bin"), dtype=np.int32)
            arrs["f%d_key" % sh] = km[c != 0]; arrs["f%d_cnt" % sh] = c[c != 0]; arrs["f%d_hom" % sh] = np.array([r["hom_cov"]], np.int32)
            print("-f%d: hom_cov %d, %d k-mers in the table, sum of counts %d" % (sh, r["hom_cov"], int((c != 0).sum()), int(c[c != 0].astype(np.int64).sum())))
    out = os.path.join(ROOT, "tests", "golden", "bloom.npz")
    np.savez_compressed(out, **arrs)
    print("->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
