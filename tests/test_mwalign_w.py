"""The band-over-lanes aligner (hb_mwalign_w.cuh) against the per-thread aligner (hb_mw_align, pinned by the golden step-B digests) on random
pairs: all four modes, bands of 1-64 words (both lane layouts), N bases, both strands.  Host emulation (virtual lanes)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(HERE, "hostemu")); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
from hifiasm_b200 import sim, binio  # noqa: E402


def cases(seed, n):
    """(reads, list of (qid, tid, rev, mode, ps0, pn, qs0, tn, thre, abs_diag))"""
    rng = np.random.default_rng(seed)
    reads, out = [], []
    for _ in range(n):
        L = int(rng.choice([40, 150, 700, 2500, 6000]))
        a = rng.integers(0, 4, L, dtype=np.uint8)
        b = sim._mutate(a.copy(), rng, float(rng.choice([0.0, 0.01, 0.05, 0.15, 0.3])))
        if rng.random() < 0.3 and b.size > 50:   # a block indel
            s = int(rng.integers(0, b.size - 20)); ln = int(rng.integers(1, min(400, b.size - s)))
            b = np.concatenate([b[:s], b[s + ln:]]) if rng.random() < 0.5 else np.concatenate([b[:s], rng.integers(0, 4, ln, dtype=np.uint8), b[s:]])
        if rng.random() < 0.2:
            b = b.copy(); b[rng.random(b.size) < 0.01] = 4
        rev = int(rng.integers(0, 2))
        t = (3 - b[::-1]).astype(np.uint8) if rev else b
        if rev:
            t = np.where(b[::-1] == 4, 4, t).astype(np.uint8)
        qid = len(reads); reads.append(a); tid = len(reads); reads.append(np.ascontiguousarray(t))
        mode = int(rng.integers(0, 4)); tn = int(a.size); pn = int(b.size)
        thre = int(rng.choice([1, 5, 20, 31, 32, 40, 100, 300, 700, 1023, 1024, 1500, 2047]))
        thre = max(1, min(thre, max(tn, pn)))
        abs_diag = 0; ps0 = 0; qs0 = 0
        if mode == 3:   # semi-global: pattern = query length + 2 * thre around the diagonal, abs_diag leading positions absent
            abs_diag = int(rng.integers(0, thre + 1)); pn = min(pn, tn + 2 * thre - abs_diag)
        out.append((qid, tid, rev, mode, ps0, pn, qs0, tn, thre, abs_diag))
    return reads, out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lanes_equal_thread(seed):
    import emu
    reads, cs = cases(seed, 120)
    flat, boff, ln, npos, noff = sim.pack_reads(reads)
    R = emu.Reads(binio.ReadStore(length=ln, byte_off=boff, packed=flat, n_off=noff, n_pos=npos))
    n_aligned = 0; words = set()
    for c in cs:
        o0, r0, c0 = emu.mw_align(R, *c, 0)
        o1, r1, c1 = emu.mw_align(R, *c, 1)
        assert o0 == 0 and o1 == 0
        assert r0[0] == r1[0], (c, r0, r1)
        if r0[0] <= c[8]:
            assert r0 == r1 and c0.tobytes() == c1.tobytes(), (c, r0, r1)
            n_aligned += 1; words.add((2 * c[8] + 1 + 63) // 64)
    assert n_aligned > 20 and max(words) > 32 and min(words) == 1, (n_aligned, sorted(words))
