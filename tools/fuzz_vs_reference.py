#!/usr/bin/env python
"""Fuzzing the device bodies against the UNMODIFIED reference on fresh read sets (CPU only; needs oracle/_ref, so build container only).

For every configuration: simulate reads, let the reference dump every state of the stage (tests/golden/make_golden.py + make_rounds.py into a
temporary directory), and run the host-emulation suites (tests/test_hostemu.py, tests/test_hostemu_round.py: sketch, anchors, chains, the EC
alignment stage incl. the rescue, phasing, consensus, lists, closing steps, final pass) on it.  A failure is a read on which the product's
code and the reference disagree.    usage: python tools/fuzz_vs_reference.py [first_seed [n_configs]]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def config(i, seed):
    k = i % 11
    g = dict(glen=90000, seed=seed, snp_rate=0.002, repeat_frac=0.1, repeat_len=1200, repeat_div=0.01)
    r = dict(cov=16, mean_len=7000, seed=seed, sd_len=2000, min_len=800, err=0.003, n_rate=1e-4)
    if k == 1:   # short reads, deep
        r.update(cov=40, mean_len=2500, sd_len=800, min_len=300); g.update(glen=40000)
    elif k == 2:  # long reads with blocks (rescue), few N
        g.update(glen=200000, repeat_frac=0.0); r.update(cov=14, mean_len=36000, sd_len=5000, min_len=20000, burst_rate=5e-5, sv_rate=3e-5, block_rate=4e-5)
    elif k == 3:  # noisy
        r.update(err=0.012, burst_rate=2e-4, sv_rate=8e-5, n_rate=5e-4)
    elif k == 4:  # clean, many repeats
        g.update(repeat_frac=0.35, repeat_len=2500, repeat_div=0.003); r.update(err=0.0005, cov=24)
    elif k == 5:  # high heterozygosity
        g.update(snp_rate=0.01); r.update(cov=20, sv_rate=4e-5)
    elif k == 6:  # exact duplicates of reads
        r.update(dup_rate=0.15, cov=14)
    elif k == 7:  # tiny reads mixed in (shorter than k)
        g.update(glen=50000); r.update(cov=25, mean_len=1500, sd_len=1400, min_len=30)
    elif k == 9:  # very long, very clean reads: match runs longer than one cigar entry holds (0x3fff)
        g.update(glen=260000, repeat_frac=0.0, snp_rate=0.0005); r.update(cov=12, mean_len=60000, sd_len=8000, min_len=30000, err=0.00004, n_rate=0.0)
    elif k == 10:  # N-rich reads
        r.update(n_rate=0.004, cov=18)
    elif k == 8:  # very deep on a small genome (chain capping)
        g.update(glen=20000, repeat_frac=0.0); r.update(cov=90, mean_len=4000, sd_len=1000, min_len=500)
    return g, r


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000; n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    bad = []
    for i in range(n):
        name = "fz%d" % i; g, r = config(i, seed0 + i)
        with tempfile.TemporaryDirectory() as td:
            env = dict(os.environ, HB_GOLDEN_DIR=td, HB_GOLDEN_EXTRA=json.dumps({name: [g, r]}), HB_GOLDEN_NAMES=name)
            for script in ("make_golden.py", "make_rounds.py", "make_outputs.py"):
                subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "golden", script), name], env=env, stdout=subprocess.DEVNULL)
            rc = subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_hostemu.py"), os.path.join(ROOT, "tests", "test_hostemu_round.py"), os.path.join(ROOT, "tests", "test_outputs.py"), "-k", "not ingest and not io_errors", "-x", "-q", "-s"], env=env, cwd=ROOT)
            print("== config %d (seed %d, kind %d): %s" % (i, seed0 + i, i % 11, "ok" if rc == 0 else "MISMATCH"), flush=True)
            if rc:
                bad.append((i, seed0 + i))
    print("mismatching configurations:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
