#!/usr/bin/env python
"""Regenerates tests/golden/g*_rounds.npz: the states of every error-correction round
(cal_ec_r, ecovlp.cpp:6268) of the UNMODIFIED reference on the golden data sets.

`refdump roundK` (oracle/refdump.cpp) runs EC rounds 0..K and, of round K, writes R_INF with
the reference's own writers at four points: after cal_ec_multiple (`hap`: paf / reverse_paf as
push_ne_ovlp left them, is_fully_corrected / is_abnormal, the edit scripts scc.a[i]), after
sl_ec_r (`sl`: corrected reads), after cal_update_ec_multiple (`upd`) and after the round
(`post`: reverse-complemented when the round does that).  Stored per round K:
    r<K>_scc / r<K>_scc_off          edit scripts (uint16 runs, push_trace_bp_f Levenshtein_distance.h:640)
    r<K>_hap_src / r<K>_hap_rev      the reference's .bin dumps of the two overlap lists (bytes)
    r<K>_dg_{sl,post}_reads          per-read digest of the read store (readdg below)
    r<K>_dg_{upd,post}_src, r<K>_dg_post_rev   per-read digests of the overlap lists (listdg below)
    r<K>_params                      tot_b / tot_e (bases, corrected bases), how often the graph consensus ran
Only runs in the build container (needs oracle/_ref/refdump)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from hifiasm_b200 import sim, binio  # noqa: E402
import make_golden as mg  # noqa: E402
import roundlib  # noqa: E402


def main():
    if not os.path.exists(mg.REFDUMP):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    only = [a for a in sys.argv[1:] if a in mg.DATASETS]
    for name, (gk, rk) in mg.DATASETS.items():
        if only and name not in only:
            continue
        h1, h2 = sim.sim_genome(**gk)
        reads = sim.sim_reads(h1, h2, **rk)
        arrs = {}
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "reads.fa")
            sim.write_fasta(fa, reads)
            for K in range(3):
                pfx = os.path.join(td, "r%d" % K)
                subprocess.check_call([mg.REFDUMP, "round%d" % K, pfx, "-o", os.path.join(td, "asm%d" % K), "-t4", "-f0", fa],
                                      stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
                p = mg.params(pfx)
                n = int(p["n_reads"])
                scc, off = [], np.zeros(n + 1, np.uint64)
                with open(pfx + ".hap.scc.bin", "rb") as f:
                    for i in range(n):
                        c = int(np.frombuffer(f.read(4), "<u4")[0])
                        scc.append(np.frombuffer(f.read(2 * c), "<u2"))
                        off[i + 1] = off[i] + c
                arrs["r%d_scc" % K] = np.concatenate(scc) if scc else np.zeros(0, np.uint16)
                arrs["r%d_scc_off" % K] = off
                for tag in ("src", "rev"):
                    arrs["r%d_hap_%s" % (K, tag)] = np.fromfile("%s.hap.ovlp.%s.bin" % (pfx, "source" if tag == "src" else "reverse"), dtype=np.uint8)
                for tag in ("sl", "post"):
                    rs = binio.load_ec_bin("%s.%s.ec.bin" % (pfx, tag))
                    arrs["r%d_dg_%s_reads" % (K, tag)] = roundlib.reads_digests(rs)
                for tag, which, is_rev in (("upd", "source", 0), ("post", "source", 0), ("post", "reverse", 1)):
                    rec, o, fc, ab = binio.load_ovlp_bin("%s.%s.ovlp.%s.bin" % (pfx, tag, which))
                    arrs["r%d_dg_%s_%s" % (K, tag, "rev" if is_rev else "src")] = roundlib.list_digests(rec, o, is_rev)
                arrs["r%d_params" % K] = np.array(["%s=%s" % kv for kv in sorted(p.items())])
                print(name, "round", K, {k: p[k] for k in ("tot_b", "tot_e", "full_calls", "full_bases")})
        out = os.path.join(mg.OUT_DIR, name + "_rounds.npz")
        np.savez_compressed(out, **arrs)
        print(name, "->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
