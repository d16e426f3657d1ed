"""ctypes wrapper of tools/simgen.c: benchmark-scale synthetic HiFi read sets (bench / test tooling)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsimgen.so")


class _Par(C.Structure):
    _fields_ = [("glen", C.c_uint64), ("seed", C.c_uint64), ("n_reads", C.c_uint64), ("mean_len", C.c_double), ("sd_len", C.c_double),
                ("min_len", C.c_uint32), ("err", C.c_double), ("n_rate", C.c_double)]


def build():
    src = os.path.join(_HERE, "simgen.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"])


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(_SO)
    return _LIB


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class ReadSet:
    """flat, byte_off, length, n_pos, n_off: what Engine.upload_reads takes"""
    def __init__(self, flat, byte_off, length, n_pos, n_off):
        self.flat, self.byte_off, self.length, self.n_pos, self.n_off = flat, byte_off, length, n_pos, n_off
        self.n = int(length.size); self.bases = int(length.sum())


def make(genome_mb: float, cov: float, seed: int = 20260923, mean_len: int = 15000, sd_len: int = 2000, min_len: int = 2000, err: float = 0.002, snp: float = 0.001,
         repeat_frac: float = 0.05, repeat_len: int = 5000, copies: int = 50, repeat_div: float = 0.01, n_rate: float = 0.0, fasta: str | None = None) -> ReadSet:
    """SURVEY.md §8(d)'s synthetic HiFi set: genome_mb Mb diploid, cov x, 15 kb reads, 0.2 % errors, 5 % of the genome in 50-copy 5 kb repeats."""
    L = _lib()
    G = int(genome_mb * 1000000)
    h1 = np.empty(G, np.uint8); h2 = np.empty(G, np.uint8)
    L.simgen_genome(C.c_uint64(G), C.c_uint64(seed), C.c_double(snp), C.c_double(repeat_frac), C.c_uint32(repeat_len), C.c_uint32(copies), C.c_double(repeat_div), _p(h1), _p(h2))
    n = int(G * cov / mean_len)
    par = _Par(G, seed, n, float(mean_len), float(sd_len), min_len, float(err), float(n_rate))
    length = np.zeros(n, np.uint64)
    L.simgen_lengths(C.byref(par), _p(h1), _p(h2), _p(length))
    byte_off = np.zeros(n + 1, np.uint64); np.cumsum(length // 4 + 1, out=byte_off[1:])
    n_off = np.zeros(n + 1, np.uint64); n_pos = np.zeros(1, np.uint64)
    if n_rate > 0:
        cnt = np.zeros(n, np.uint64)
        L.simgen_ncount(C.byref(par), _p(h1), _p(h2), _p(length), _p(cnt))
        np.cumsum(cnt, out=n_off[1:]); n_pos = np.zeros(max(1, int(n_off[-1])), np.uint64)
    flat = np.zeros(int(byte_off[-1]), np.uint8)
    fa_off = fa = None
    if fasta:
        nl = np.array([len(">r%d\n" % i) for i in range(n)], np.uint64)
        fa_off = np.zeros(n + 1, np.uint64); np.cumsum(nl + length + 1, out=fa_off[1:])
        fa = np.empty(int(fa_off[-1]) + 16, np.uint8)
    L.simgen_fill(C.byref(par), _p(h1), _p(h2), _p(length), _p(byte_off), _p(flat), _p(n_off) if n_rate > 0 else C.c_void_p(0), _p(n_pos) if n_rate > 0 else C.c_void_p(0), _p(fa_off), _p(fa))
    if fasta:
        with open(fasta, "wb") as f:
            f.write(memoryview(fa)[:int(fa_off[-1])])
    return ReadSet(flat, byte_off, length, n_pos, n_off)
