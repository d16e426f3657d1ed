"""Seeded synthetic HiFi read sets of the shapes SURVEY.md §8(d) names.

Test/bench tooling only (numpy on the host).  Genome: uniform random ACGT
plus an optional repeat family (so the high-occurrence filter table,
`ha_ft_cnt`, high-occ thinning and `max_n_chain` capping are exercised);
haplotype 2 = haplotype 1 + SNPs; reads sampled uniformly from both
haplotypes and both strands, Gaussian length clipped at `min_len`; HiFi
errors split sub/ins/del 35/30/35 %.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def sim_genome(glen: int, seed: int, snp_rate: float = 0.001,
               repeat_frac: float = 0.0, repeat_len: int = 5000,
               repeat_div: float = 0.01):
    """Return (hap1, hap2) as uint8 code arrays (A0 C1 G2 T3)."""
    rng = np.random.default_rng(seed)
    hap1 = rng.integers(0, 4, glen, dtype=np.uint8)
    if repeat_frac > 0 and glen > 4 * repeat_len:
        unit = rng.integers(0, 4, repeat_len, dtype=np.uint8)
        n_copies = max(2, int(glen * repeat_frac / repeat_len))
        starts = rng.integers(0, glen - repeat_len, n_copies)
        for s in starts:
            cp = unit.copy()
            m = rng.random(repeat_len) < repeat_div
            cp[m] = (cp[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
            hap1[s:s + repeat_len] = cp
    hap2 = hap1.copy()
    m = rng.random(glen) < snp_rate
    hap2[m] = (hap2[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    return hap1, hap2


def _mutate(seq: np.ndarray, rng, err: float) -> np.ndarray:
    n = seq.size
    m = rng.random(n) < err
    k = int(m.sum())
    if k == 0:
        return seq
    pos = np.nonzero(m)[0]
    kind = rng.random(k)
    out = seq.copy()
    cnt = np.ones(n, dtype=np.int64)
    sub = pos[kind < 0.35]
    out[sub] = (out[sub] + rng.integers(1, 4, sub.size, dtype=np.uint8)) & 3
    ins = pos[(kind >= 0.35) & (kind < 0.65)]
    cnt[ins] = 2
    dele = pos[kind >= 0.65]
    cnt[dele] = 0
    res = np.repeat(out, cnt)
    if ins.size:
        # the second copy of an inserted position becomes a random base
        cs = np.cumsum(cnt)
        ipos = cs[ins] - 1
        res[ipos] = rng.integers(0, 4, ins.size, dtype=np.uint8)
    return res


def _damage(seq: np.ndarray, rng, burst_rate: float, sv_rate: float, block_rate: float = 0.0) -> np.ndarray:
    """Local damage that makes alignment windows fail (exercises the gap-filling /
    extension / re-chaining paths of the EC rounds): error bursts (60-300 bp at
    8-25 % error), block insertions / deletions of 20-400 bp, and (block_rate)
    replacements of 540-900 bp by a short tandem repeat of about the same length, often with a noisy
    stretch (6-15 % errors) next to it — the
    unaligned stretches >= 512 bp on both reads that rechain_aln_hc re-seeds."""
    n = seq.size
    nb = rng.poisson(burst_rate * n)
    for _ in range(nb):
        if seq.size < 400:
            break
        ln = int(rng.integers(60, 300)); s = int(rng.integers(0, seq.size - ln))
        seq = np.concatenate([seq[:s], _mutate(seq[s:s + ln], rng, float(rng.uniform(0.08, 0.25))), seq[s + ln:]])
    ns = rng.poisson(sv_rate * n)
    for _ in range(ns):
        if seq.size < 1000:
            break
        ln = int(rng.integers(20, 400)); s = int(rng.integers(0, seq.size - ln))
        if rng.random() < 0.5:
            seq = np.concatenate([seq[:s], rng.integers(0, 4, ln, dtype=np.uint8), seq[s:]])
        else:
            seq = np.concatenate([seq[:s], seq[s + ln:]])
    if block_rate > 0:
        for _ in range(rng.poisson(block_rate * n)):
            if seq.size < 6000:
                break
            ln = int(rng.integers(540, 900)); ln2 = ln + int(rng.integers(-20, 21)); s = int(rng.integers(50, seq.size - ln - 50))
            unit = rng.integers(0, 4, int(rng.integers(2, 7)), dtype=np.uint8)
            nz = int(rng.integers(0, int(0.4 * ln))) if rng.random() < 0.5 else 0   # a noisy stretch next to it: too many errors for 51-mers, exact runs of >= 10 bases remain
            e = min(seq.size, s + ln + nz)
            noisy = _mutate(seq[s + ln:e], rng, float(rng.uniform(0.06, 0.15))) if nz else seq[s + ln:e]
            parts = [seq[:s], np.resize(unit, ln2), noisy, seq[e:]]
            if rng.random() < 0.5:
                parts = [seq[:s], noisy, np.resize(unit, ln2), seq[e:]]
            seq = np.concatenate(parts)
    return seq


def sim_reads(hap1: np.ndarray, hap2: np.ndarray, cov: float, mean_len: int,
              seed: int, sd_len: int = 2000, min_len: int = 2000,
              err: float = 0.002, n_rate: float = 0.0,
              burst_rate: float = 0.0, sv_rate: float = 0.0, block_rate: float = 0.0, dup_rate: float = 0.0):
    """Return list of uint8 code arrays (values 0..3, 4 = N)."""
    rng = np.random.default_rng(seed + 1000003)
    glen = hap1.size
    n_reads = int(glen * cov / mean_len)
    lens = np.clip(rng.normal(mean_len, sd_len, n_reads).astype(np.int64), min_len, glen)
    starts = (rng.random(n_reads) * (glen - lens + 1)).astype(np.int64)
    hap = rng.integers(0, 2, n_reads)
    strand = rng.integers(0, 2, n_reads)
    reads = []
    for i in range(n_reads):
        src = hap2 if hap[i] else hap1
        s = src[starts[i]:starts[i] + lens[i]]
        if strand[i]:
            s = (3 - s[::-1]).astype(np.uint8)
        if err > 0:
            s = _mutate(s, rng, err)
        if burst_rate > 0 or sv_rate > 0 or block_rate > 0:
            s = _damage(s, rng, burst_rate, sv_rate, block_rate)
        if n_rate > 0:
            s = s.copy()
            s[rng.random(s.size) < n_rate] = 4
        reads.append(np.ascontiguousarray(s))
        if dup_rate > 0 and rng.random() < dup_rate:   # an exact duplicate of the read (PCR-like): identical overlaps, ties everywhere
            reads.append(reads[-1].copy())
    return reads


def write_fasta(path: str, reads) -> None:
    tab = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(tab[r].tobytes())
            f.write(b"\n")


def pack_reads(reads):
    """2-bit pack like `ha_compress_base` (Process_Read.cpp:792): byte =
    b0<<6|b1<<4|b2<<2|b3, len/4+1 bytes per read, N stored as A + side list.
    Returns (packed_flat u8, byte_off u64[n+1], length u64[n], n_pos u64 flat,
    n_off u64[n+1])."""
    n = len(reads)
    length = np.array([r.size for r in reads], dtype=np.uint64)
    nbytes = (length // 4 + 1).astype(np.uint64)
    byte_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nbytes, out=byte_off[1:])
    flat = np.zeros(int(byte_off[-1]), dtype=np.uint8)
    npos = []
    n_off = np.zeros(n + 1, dtype=np.uint64)
    for i, r in enumerate(reads):
        isn = r > 3
        if isn.any():
            p = np.nonzero(isn)[0].astype(np.uint64)
            npos.append(p)
            n_off[i + 1] = p.size
            r = np.where(isn, 0, r).astype(np.uint8)
        L = r.size
        pad = (-L) % 4
        q = np.concatenate([r, np.zeros(pad, dtype=np.uint8)]) if pad else r
        q = q.reshape(-1, 4)
        b = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        o = int(byte_off[i])
        flat[o:o + b.size] = b
    np.cumsum(n_off[1:], out=n_off[1:])
    n_pos = np.concatenate(npos) if npos else np.zeros(0, dtype=np.uint64)
    return flat, byte_off, length, n_pos, n_off
