#!/usr/bin/env python
"""The filter table behind the reference's Bloom filter on fresh read sets: the unmodified reference (refdump ftq) vs the oracle restatement
and vs the device formulation in host emulation, for several -f.  CPU only, build container only.   usage: python tools/fuzz_bloom.py [seed [n]]"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
from hifiasm_b200 import sim  # noqa: E402
import ha_oracle as ho  # noqa: E402
import emu  # noqa: E402


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 100; n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    refdump = os.path.join(ROOT, "oracle", "_ref", "refdump"); bad = 0
    for i in range(n):
        rng = np.random.default_rng(seed0 + i)
        h1, h2 = sim.sim_genome(int(rng.integers(150000, 600000)), seed0 + i, snp_rate=0.001, repeat_frac=float(rng.uniform(0.02, 0.15)), repeat_len=int(rng.integers(800, 3000)), repeat_div=0.002)
        rr = sim.sim_reads(h1, h2, cov=float(rng.uniform(5, 12)), mean_len=8000, seed=seed0 + i, sd_len=2000, min_len=1500, err=0.002)
        for _ in range(int(rng.integers(1, 4))):  # late high-count k-mers around the cut-off: copies of extra reads appended at the end
            rr = rr + [rng.integers(0, 4, int(rng.integers(2000, 6000)), dtype=np.uint8)] * int(rng.integers(20, 70))
        flat, boff, ln, npos, noff = sim.pack_reads(rr)
        st = ho.Store(ln, boff, flat, noff, npos); opt = ho.default_opt()
        allk = ho.all_kmers(st, opt); km = np.unique(allk)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "r.fa"); sim.write_fasta(fa, rr); km.tofile(os.path.join(td, "h.bin"))
            for sh in (0, 21, 22, 23, 25):
                out = subprocess.run([refdump, "ftq", os.path.join(td, "h.bin"), os.path.join(td, "c.bin"), "-o", os.path.join(td, "asm"), "-t4", "-f%d" % sh, fa], capture_output=True, text=True, check=True)
                hom = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])["hom_cov"]
                ref = np.fromfile(os.path.join(td, "c.bin"), dtype=np.int32); rk, rc = km[ref != 0], ref[ref != 0].astype(np.int64)
                ft, ohom = ho.ft_gen(st, opt, sh)
                mine = np.array([ho.lib().hao_ft_cnt(C.c_void_p(ft), C.c_uint64(int(h))) for h in rk], dtype=np.int64)
                ok_o = ohom == hom and int(ho.lib().hao_ft_size(C.c_void_p(ft))) == rk.size and bool((mine == rc).all())
                key, cnt = emu.bf_counts(allk, sh); keep = cnt >= int(hom * 5.0); c = cnt[keep].astype(np.int64); c[c > 2000] = 2**31 - 1
                ok_e = int(keep.sum()) == rk.size and bool((key[keep] == rk).all()) and bool((c == rc).all())
                print("set %d (-f%d): hom %d, table %d, sum %d: oracle %s, device formulation %s" % (i, sh, hom, rk.size, int(rc[rc < 2**31 - 1].sum()), "ok" if ok_o else "MISMATCH", "ok" if ok_e else "MISMATCH"), flush=True)
                bad += (not ok_o) + (not ok_e)
    print("mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
