#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference.

Needs oracle/_ref/refdump (built by `make -C oracle ref` from /root/reference),
so it only runs in the build container; the .npz files it writes are committed
and are what the CPU and GPU test suites read.

For every data set it runs
    refdump raw   (filter table + round-0 index on raw reads, stage dumps)
    refdump final (3 EC rounds, pre-final bins, final index, stage dumps,
                   cal_ov_r, final bins)
and stores: the reference's own .bin dumps (bytes), the run parameters, and a
blake2b-64 digest per read of each intermediate array (sketch, index probes,
anchors, chains + fake cigars, chain hits), plus the full arrays of the first
few reads for debugging.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hifiasm_b200 import sim  # noqa: E402
import alnlib  # noqa: E402

REFDUMP = os.path.join(ROOT, "oracle", "_ref", "refdump")
CH = np.dtype([("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_id", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
               ("y_pos_strand", "<u4"), ("shared_seed", "<i4"), ("first_hit", "<u4"), ("n_fc", "<u4")])

DATASETS = {
    # name: (genome kwargs, read kwargs)
    "g1": (dict(glen=120000, seed=5, snp_rate=0.002, repeat_frac=0.2, repeat_len=1500, repeat_div=0.005),
           dict(cov=20, mean_len=6000, seed=5, sd_len=2000, min_len=400, n_rate=2e-4)),
    "g2": (dict(glen=150000, seed=9, snp_rate=0.001, repeat_frac=0.0),
           dict(cov=16, mean_len=9000, seed=9, sd_len=2500, min_len=1000)),
    # damaged reads (error bursts, block indels, 0.4 % errors): windows fail, so the gap-filling / extension /
    # re-chaining paths of the EC alignment stage run
    "g3": (dict(glen=100000, seed=13, snp_rate=0.002, repeat_frac=0.1, repeat_len=1200, repeat_div=0.01),
           dict(cov=18, mean_len=7000, seed=13, sd_len=2000, min_len=800, err=0.004, n_rate=1e-4, burst_rate=1.5e-4, sv_rate=6e-5)),
    # long reads with tandem-repeat blocks of 540-900 bp (plus noisy stretches): accepted overlaps keep unaligned windows of >= 512 bp on
    # both reads, which the reference re-seeds (rechain_aln_hc, Correct.cpp:17669) — all three window modes, the replaced-window case, the
    # quick check, the fixed-end DP in both directions and the end-point re-chaining (lchain_simple) occur
    "g4": (dict(glen=300000, seed=53, snp_rate=0.002, repeat_frac=0.0),
           dict(cov=16, mean_len=40000, seed=53, sd_len=6000, min_len=25000, err=0.003, n_rate=1e-4, burst_rate=5e-5, sv_rate=3e-5, block_rate=3e-5)),
}


# fuzzing against the reference (tools/fuzz_vs_reference.py): extra data sets and another output directory through the environment
OUT_DIR = os.environ.get("HB_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden"))
if os.environ.get("HB_GOLDEN_EXTRA"):
    import json
    DATASETS.update({k: (v[0], v[1]) for k, v in json.loads(os.environ["HB_GOLDEN_EXTRA"]).items()})


def dg(b: bytes) -> int:
    return int.from_bytes(hashlib.blake2b(b, digest_size=8).digest(), "little")


def read_stages(pfx, n_reads, keep=4):
    out = {k: np.zeros(n_reads, dtype=np.uint64) for k in ("mz", "idx", "anchors", "chains", "chain_hits", "windows")}
    cnt = {k: np.zeros(n_reads, dtype=np.uint64) for k in ("mz", "anchors", "chains", "chain_hits", "windows")}
    full = {}
    with open(pfx + ".mz.bin", "rb") as fm, open(pfx + ".idx.bin", "rb") as fi, \
            open(pfx + ".anchors.bin", "rb") as fa, open(pfx + ".chains.bin", "rb") as fc, open(pfx + ".windows.bin", "rb") as fw:
        for i in range(n_reads):
            n = int(np.frombuffer(fm.read(4), dtype="<u4")[0])
            mz = fm.read(16 * n)
            out["mz"][i] = dg(mz); cnt["mz"][i] = n
            h = hashlib.blake2b(digest_size=8)
            for _ in range(n):
                hd = fi.read(12)
                c = int(np.frombuffer(hd[8:], dtype="<u4")[0])
                h.update(hd); h.update(fi.read(8 * c))
            out["idx"][i] = int.from_bytes(h.digest(), "little")
            na = int(np.frombuffer(fa.read(8), dtype="<u8")[0])
            an = fa.read(16 * na)
            out["anchors"][i] = dg(an); cnt["anchors"][i] = na
            nc = int(np.frombuffer(fc.read(4), dtype="<u4")[0]); nh = int(np.frombuffer(fc.read(8), dtype="<u8")[0])
            h = hashlib.blake2b(digest_size=8)
            chl = []
            for _ in range(nc):
                rb = fc.read(CH.itemsize)
                r = np.frombuffer(rb, dtype=CH)[0]
                fcb = fc.read(8 * int(r["n_fc"]))
                h.update(rb); h.update(fcb)
                chl.append(rb)
            out["chains"][i] = int.from_bytes(h.digest(), "little"); cnt["chains"][i] = nc
            hb = fc.read(16 * nh)
            out["chain_hits"][i] = dg(hb); cnt["chain_hits"][i] = nh
            nwin = int(np.frombuffer(fw.read(4), dtype="<u4")[0])
            wb = fw.read(40 * nwin)  # {chain, q_s, q_e, t_s, t_pri_l, thre, aux_beg, aux_end, err, pe} per window
            out["windows"][i] = dg(wb); cnt["windows"][i] = nwin
            if i < keep:
                full["mz_%d" % i] = np.frombuffer(mz, dtype=np.uint8)
                full["anchors_%d" % i] = np.frombuffer(an, dtype=np.uint8)
                full["chains_%d" % i] = np.frombuffer(b"".join(chl), dtype=np.uint8)
                full["chain_hits_%d" % i] = np.frombuffer(hb, dtype=np.uint8)
    # alignment stage of an EC round (refdump step 6): per-read digests of the state after steps A, B, C
    aln = alnlib.read_aln(pfx + ".aln.bin")
    assert len(aln) == n_reads
    for k in ("alnA", "alnB", "alnC"):
        out[k] = np.zeros(n_reads, dtype=np.uint64)
    cnt["aln_ok"] = np.zeros(n_reads, dtype=np.uint64)
    for i, rl in enumerate(aln):
        out["alnA"][i] = alnlib.digest_A((d["st"], d["align_length"], d["rr"], d["re"], d["A"][0], d["A"][1]) for d in rl)
        acc = [d for d in rl if d["st"] == 2]
        out["alnB"][i] = alnlib.digest_B((d["reB"], d["B"][0], d["B"][1]) for d in acc)
        out["alnC"][i] = alnlib.digest_C((d["nheC"], d["xyC"], d["C"][0], d["C"][1]) for d in acc)
        cnt["aln_ok"][i] = len(acc)
    # phasing + de-duplication (refdump step 7): digest of the per-overlap records after rphase_hc and after dedup_chains
    ph = alnlib.read_phase(pfx + ".phase.bin")
    assert len(ph) == n_reads
    out["phase"] = np.zeros(n_reads, dtype=np.uint64); out["dedup"] = np.zeros(n_reads, dtype=np.uint64)
    cnt["phase_hap2"] = np.zeros(n_reads, dtype=np.uint64); cnt["dedup"] = np.zeros(n_reads, dtype=np.uint64)
    out["rpaf"] = np.zeros(n_reads, dtype=np.uint64); cnt["rpaf"] = np.zeros(n_reads, dtype=np.uint64)
    for i, (a, b, rp) in enumerate(ph):
        out["rpaf"][i] = dg(np.ascontiguousarray(rp).tobytes()); cnt["rpaf"][i] = rp.size
        out["phase"][i] = dg(np.ascontiguousarray(a).tobytes()); out["dedup"][i] = dg(np.ascontiguousarray(b).tobytes())
        cnt["phase_hap2"][i] = int((a["is_match"] == 2).sum()); cnt["dedup"][i] = b.size
    # the whole alignment stage through gen_hc_r_alin_ea with the previous round's overlaps (refdump step 8, row a12)
    ea = alnlib.read_ea(pfx + ".ea.bin")
    assert len(ea) == n_reads
    out["ea"] = np.zeros(n_reads, dtype=np.uint64); cnt["ea"] = np.zeros(n_reads, dtype=np.uint64)
    for i, rl in enumerate(ea):
        out["ea"][i] = alnlib.digest_ea((r8, w, c) for r8, (w, c) in rl); cnt["ea"][i] = len(rl)
    return out, cnt, full


def params(pfx):
    return {k: v for k, v in (l.split() for l in open(pfx + ".params.txt"))}


def make_ed_cases(n_cases=3000, seed=3):
    """Window-alignment cases shaped like align_hc_ed_post_extz's calls
    (Correct.cpp:12951-13011): pattern = target slice of q_l + 2*thre bases
    (clipped by aux_beg / aux_end at read ends), text = query window."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cases = []
    for c in range(n_cases):
        ql = int(rng.integers(4, 776)) if c % 7 else 775
        e_rate = 0.04 if c % 3 else 0.07
        thre = int(ql * e_rate)
        if thre == 0 and ql >= 4:
            thre = 1
        thre = min(thre, 31)
        text = rng.integers(0, 4, ql, dtype=np.uint8)
        core = text.copy()
        nerr = int(rng.integers(0, 2 * thre + 2)) if c % 5 else 0
        for _ in range(nerr):
            p = int(rng.integers(0, max(1, core.size)))
            t = int(rng.integers(0, 3))
            if t == 0:
                core[p] = (core[p] + 1 + rng.integers(0, 3)) & 3
            elif t == 1:
                core = np.insert(core, p, rng.integers(0, 4))
            elif core.size > 1:
                core = np.delete(core, p)
        shift = int(rng.integers(-thre, thre + 1)) if c % 4 == 0 else 0
        left = rng.integers(0, 4, max(0, thre + shift), dtype=np.uint8)
        right = rng.integers(0, 4, 4 * thre + 8, dtype=np.uint8)
        pat = np.concatenate([left, core, right]).astype(np.uint8)
        aux_beg = int(rng.integers(0, thre + 1)) if c % 6 == 0 else 0
        aux_end = int(rng.integers(0, thre + 1)) if c % 9 == 0 else 0
        pn = ql + 2 * thre - aux_beg - aux_end
        pat = pat[aux_beg:aux_beg + pn]
        if c % 11 == 0 and pat.size:
            pat[int(rng.integers(0, pat.size))] = 4
        tab = np.frombuffer(b"ACGTN", dtype=np.uint8)
        cases.append((pat.size, ql, thre, aux_beg, tab[pat].tobytes(), tab[text].tobytes()))
    return cases


def make_ed_golden():
    cases = make_ed_cases()
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "c.bin"), os.path.join(td, "o.bin")
        with open(fi, "wb") as f:
            f.write(np.array([len(cases)], dtype="<i4").tobytes())
            for pn, tn, thre, ab, p, t in cases:
                f.write(np.array([pn, tn, thre, ab], dtype="<i4").tobytes()); f.write(p); f.write(t)
        subprocess.check_call([REFDUMP, "edsemi", fi, fo])
        res = np.fromfile(fo, dtype="<i4").reshape(-1, 2)
    hdr = np.array([[c[0], c[1], c[2], c[3]] for c in cases], dtype=np.int32)
    pat = np.frombuffer(b"".join(c[4] for c in cases), dtype=np.uint8)
    txt = np.frombuffer(b"".join(c[5] for c in cases), dtype=np.uint8)
    out = os.path.join(ROOT, "tests", "golden", "ed_semi.npz")
    np.savez_compressed(out, hdr=hdr, pat=pat, txt=txt, res=res)
    print("ed_semi", len(cases), "cases; aligned", int((res[:, 0] != 2**31 - 1).sum()), "->", out, os.path.getsize(out))


def main():
    if not os.path.exists(REFDUMP):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    only = [a for a in sys.argv[1:] if a in DATASETS]   # `make_golden.py g4` regenerates only that set
    if not only:
        make_ed_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "ed":
        return
    for name, (gk, rk) in DATASETS.items():
        if only and name not in only:
            continue
        h1, h2 = sim.sim_genome(**gk)
        reads = sim.sim_reads(h1, h2, **rk)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "reads.fa")
            sim.write_fasta(fa, reads)
            arrs = {}
            for mode in ("raw", "final"):
                pfx = os.path.join(td, mode)
                subprocess.check_call([REFDUMP, mode, pfx, "-o", os.path.join(td, "asm_" + mode), "-t4", "-f0", fa],
                                      stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
                p = params(pfx)
                n = int(p["n_reads"])
                d, c, full = read_stages(pfx, n)
                for k, v in d.items():
                    arrs["%s_dg_%s" % (mode, k)] = v
                for k, v in c.items():
                    arrs["%s_n_%s" % (mode, k)] = v
                for k, v in full.items():
                    arrs["%s_full_%s" % (mode, k)] = v
                arrs["%s_params" % mode] = np.array(["%s=%s" % kv for kv in sorted(p.items())])
                tags = ["raw"] if mode == "raw" else ["pre", "fin"]
                for tag in tags:
                    for suf in ("ec", "ovlp.source", "ovlp.reverse"):
                        if tag == "raw" and suf != "ec":
                            continue
                        if tag == "fin" and suf == "ec":
                            continue
                        arrs["%s_%s" % (tag, suf.replace(".", "_"))] = np.fromfile("%s.%s.%s.bin" % (pfx, tag, suf), dtype=np.uint8)
            out = os.path.join(OUT_DIR, name + ".npz")
            np.savez_compressed(out, **arrs)
            print(name, len(reads), "reads", sum(r.size for r in reads), "bases ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
