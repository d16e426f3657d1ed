"""Digests of the states an error-correction round passes through (tests/golden/make_rounds.py writes them from the
unmodified reference; the CPU and GPU suites compute the same digests from the product's results)."""
import hashlib

import numpy as np

from hifiasm_b200 import binio


def dg(b: bytes) -> int:
    return int.from_bytes(hashlib.blake2b(b, digest_size=8).digest(), "little")


def read_digest(length, packed, n_pos) -> int:
    """one read: length, the len/4+1 packed bytes with the undefined pad bits masked, the N positions"""
    L = int(length)
    b = np.array(packed[:L // 4 + 1], dtype=np.uint8, copy=True)
    if L % 4 == 0:
        b[L // 4] = 0
    else:
        b[L // 4] &= (0xFF << (2 * (4 - L % 4))) & 0xFF
    return dg(np.array([L], "<u8").tobytes() + b.tobytes() + np.asarray(n_pos, dtype="<u8").tobytes())


def reads_digests(rs) -> np.ndarray:
    out = np.zeros(rs.n, np.uint64)
    for i in range(rs.n):
        o = int(rs.byte_off[i])
        out[i] = read_digest(rs.length[i], rs.packed[o:o + int(rs.length[i]) // 4 + 1], rs.n_pos[int(rs.n_off[i]):int(rs.n_off[i + 1])])
    return out


def canon_list(rec, is_rev) -> np.ndarray:
    """overlap records as MA_DISK with the fields the EC rounds leave undefined zeroed: `del` everywhere (push_ne_ovlp,
    ecovlp.cpp:2585, never writes it) and `el` of the reverse lists (only written when an edit script is passed)"""
    d = np.zeros(rec.size, dtype=binio.MA_DISK)
    for f in binio.MA_DISK.names:
        d[f] = rec[f]
    d["del"] = 0
    if is_rev:
        d["el"] = 0
    return d


def list_digests(rec, off, is_rev) -> np.ndarray:
    n = off.size - 1
    d = canon_list(rec, is_rev)
    out = np.zeros(n, np.uint64)
    for i in range(n):
        out[i] = dg(d[int(off[i]):int(off[i + 1])].tobytes())
    return out


class Rounds:
    """tests/golden/<name>_rounds.npz"""
    def __init__(self, name):
        import os
        import goldenlib
        self.z = np.load(os.path.join(goldenlib.GOLDEN, name + "_rounds.npz"))
        self._cache = {}

    def params(self, K):
        return dict(s.split("=") for s in self.z["r%d_params" % K])

    def scc(self, K):
        return self.z["r%d_scc" % K], self.z["r%d_scc_off" % K]

    def hap(self, K, which):
        """-> (records MA_DISK, off, is_fully_corrected, is_abnormal) of the list after cal_ec_multiple"""
        import goldenlib
        key = (K, which)
        if key not in self._cache:
            self._cache[key] = goldenlib._load_bin(self.z["r%d_hap_%s" % (K, which)], binio.load_ovlp_bin)
        return self._cache[key]

    def digest(self, K, what):
        return self.z["r%d_dg_%s" % (K, what)]
