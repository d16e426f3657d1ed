"""ctypes binding of oracle/libha_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MZ = np.dtype([("x", "<u8"), ("info", "<u8")])
HIT = np.dtype([("id_strand", "<u4"), ("offset", "<u4"), ("self_offset", "<u4"), ("cnt", "<u4")])
MA = np.dtype([("qns", "<u8"), ("qe", "<u4"), ("tn", "<u4"), ("ts", "<u4"), ("te", "<u4"),
               ("ml", "<u4"), ("rev", "<u4"), ("bl", "<u4"), ("del", "<u4"),
               ("el", "u1"), ("no_l_indel", "u1"), ("pad", "u1", (6,))])
OVLP = np.dtype([("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_id", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                 ("y_pos_strand", "<u4"), ("shared_seed", "<i4"), ("align_length", "<u4"),
                 ("non_homopolymer_errors", "<u4"), ("overlapLen", "<u4"), ("is_match", "u1"),
                 ("without_large_indel", "u1"), ("strong", "i1"), ("pad", "u1"), ("fc_off", "<u4"), ("fc_n", "<u4")])


class Reads(C.Structure):
    _fields_ = [("n", C.c_uint64), ("len", C.c_void_p), ("off", C.c_void_p), ("packed", C.c_void_p),
                ("n_off", C.c_void_p), ("n_pos", C.c_void_p)]


class Opt(C.Structure):
    _fields_ = [("k", C.c_int), ("w", C.c_int), ("is_hpc", C.c_int), ("mz_sample_dist", C.c_int),
                ("mz_rewin", C.c_int), ("min_hist_kmer_cnt", C.c_int), ("high_factor", C.c_double),
                ("max_kmer_cnt", C.c_int), ("max_n_chain", C.c_int), ("hom_cov", C.c_int), ("het_cov", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libha_oracle.so")
        srcs = [os.path.join(_HERE, f) for f in ("ha_oracle.c", "ha_ec.c")]
        if (not os.path.exists(so)) or any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
            subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.hao_ft_gen.restype = C.c_void_p
        L.hao_ft_gen_bf.restype = C.c_void_p
        L.hao_pt_gen.restype = C.c_void_p
        L.hao_pt_get.restype = C.c_void_p
        L.hao_ft_cnt.restype = C.c_int32
        L.hao_ft_size.restype = C.c_uint64
        L.hao_pt_tot_pos.restype = C.c_uint64
        L.hao_pt_n_keys.restype = C.c_uint64
        _LIB = L
    return _LIB


class Store:
    """Keeps numpy arrays alive behind a hao_reads_t."""

    def __init__(self, length, byte_off, packed, n_off=None, n_pos=None):
        self.length = np.ascontiguousarray(length, dtype=np.uint64)
        self.byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        self.packed = np.ascontiguousarray(packed, dtype=np.uint8)
        n = self.length.size
        self.n_off = np.ascontiguousarray(n_off if n_off is not None else np.zeros(n + 1), dtype=np.uint64)
        self.n_pos = np.ascontiguousarray(n_pos if n_pos is not None else np.zeros(1), dtype=np.uint64)
        self.c = Reads(n, self.length.ctypes.data, self.byte_off.ctypes.data, self.packed.ctypes.data,
                       self.n_off.ctypes.data, self.n_pos.ctypes.data)

    @property
    def n(self):
        return int(self.length.size)

    def decode(self, i) -> bytes:
        buf = C.create_string_buffer(int(self.length[i]) + 1)
        lib().hao_decode(C.byref(self.c), C.c_uint64(i), buf)
        return buf.raw[:int(self.length[i])]

    def decode_sub(self, i, start, ln, strand) -> bytes:
        buf = C.create_string_buffer(int(ln) + 1)
        lib().hao_decode_sub(C.byref(self.c), C.c_uint64(i), C.c_int64(start), C.c_int64(ln), C.c_int(strand), buf)
        return buf.raw[:int(ln)]


def default_opt() -> Opt:
    o = Opt()
    lib().hao_opt_default(C.byref(o))
    return o


def _take(ptr, n, dtype):
    if n == 0:
        if ptr.value:
            lib().hao_free(ptr)
        return np.zeros(0, dtype=dtype)
    a = np.frombuffer((C.c_uint8 * (n * dtype.itemsize)).from_address(ptr.value), dtype=dtype).copy()
    lib().hao_free(ptr)
    return a


def sketch(seq: bytes, w, k, rid, is_hpc, ft, sample_dist, rewin):
    out = C.c_void_p(); n = C.c_uint32()
    rc = lib().hao_sketch(seq, C.c_int(len(seq)), C.c_int(w), C.c_int(k), C.c_uint32(rid), C.c_int(is_hpc),
                          C.c_void_p(ft), C.c_int(sample_dist), C.c_int(rewin), C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n.value, MZ)


def ft_gen(store: Store, opt: Opt, bf_shift: int = 0):
    """bf_shift = hifiasm's -f: 0 = exact counting, >= 21 = through the blocked Bloom filter (37 is hifiasm's default)"""
    hom = C.c_int()
    ft = lib().hao_ft_gen_bf(C.byref(store.c), C.byref(opt), C.c_int(bf_shift), C.byref(hom))
    return ft, hom.value


def all_kmers(store: Store, opt: Opt):
    """every HPC k-mer hash of the store, in read / position order"""
    L = lib(); L.hao_all_kmers.restype = C.c_uint64
    n = L.hao_all_kmers(C.byref(store.c), C.byref(opt), C.c_void_p(0), C.c_uint64(0))
    out = np.zeros(max(int(n), 1), np.uint64)
    L.hao_all_kmers(C.byref(store.c), C.byref(opt), C.c_void_p(out.ctypes.data), C.c_uint64(int(n)))
    return out[:int(n)]


def pt_gen(store: Store, ft, opt: Opt):
    hom = C.c_int(); het = C.c_int()
    pt = lib().hao_pt_gen(C.byref(store.c), C.c_void_p(ft), C.byref(opt), C.byref(hom), C.byref(het))
    return pt, hom.value, het.value


def pt_get(pt, h):
    n = C.c_int()
    p = lib().hao_pt_get(C.c_void_p(pt), C.c_uint64(int(h)), C.byref(n))
    if n.value == 0:
        return np.zeros(0, dtype=np.uint64)
    return np.frombuffer((C.c_uint8 * (n.value * 8)).from_address(p), dtype=np.uint64).copy()


def anchors(store: Store, pt, mz: np.ndarray, high_occ, low_occ):
    out = C.c_void_p(); n = C.c_uint64()
    mz = np.ascontiguousarray(mz)
    lib().hao_anchors(C.byref(store.c), C.c_void_p(pt), C.c_void_p(mz.ctypes.data), C.c_uint32(mz.size),
                      C.c_uint32(high_occ), C.c_uint32(low_occ), C.byref(out), C.byref(n))
    return _take(out, n.value, HIT)


def lchain(store: Store, rid, hits: np.ndarray, bw, k, max_n_chain):
    """-> (chains OVLP[], compacted hits HIT[], fake cigar pool u64[])"""
    hits = np.ascontiguousarray(hits).copy()
    nh = C.c_uint64(hits.size); out = C.c_void_p(); n = C.c_uint32(); fc = C.c_void_p(); nfc = C.c_uint64()
    lib().hao_lchain(C.byref(store.c), C.c_uint32(rid), C.c_void_p(hits.ctypes.data), C.byref(nh), C.c_double(bw),
                     C.c_int(k), C.c_int(max_n_chain), C.byref(out), C.byref(n), C.byref(fc), C.byref(nfc))
    return _take(out, n.value, OVLP), hits[:nh.value], _take(fc, nfc.value, np.dtype("<u8"))


def final_read(store: Store, pt, ft, opt: Opt, rid, in0: np.ndarray, in1: np.ndarray):
    in0 = np.ascontiguousarray(in0).copy(); in1 = np.ascontiguousarray(in1)
    o0 = C.c_void_p(); o1 = C.c_void_p(); m0 = C.c_uint32(); m1 = C.c_uint32()
    lib().hao_final_read(C.byref(store.c), C.c_void_p(pt), C.c_void_p(ft), C.byref(opt), C.c_uint32(rid),
                         C.c_void_p(in0.ctypes.data), C.c_uint32(in0.size), C.c_void_p(in1.ctypes.data), C.c_uint32(in1.size),
                         C.byref(o0), C.byref(m0), C.byref(o1), C.byref(m1))
    return _take(o0, m0.value, MA), _take(o1, m1.value, MA)


WIN = np.dtype([(f, "<i4") for f in ("chain", "q_s", "q_e", "t_s", "t_pri_l", "thre", "aux_beg", "aux_end", "err", "pe")])


def windows(store: Store, rid, chains: np.ndarray, fc: np.ndarray, e_rate=0.04, w_l=775):
    chains = np.ascontiguousarray(chains); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1), dtype=np.uint64)
    out = C.c_void_p(); n = C.c_uint32()
    rc = lib().hao_windows(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size), C.c_void_p(fc.ctypes.data),
                           C.c_double(e_rate), C.c_int64(w_l), C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n.value, WIN)


WL = np.dtype([("x_start", "<i4"), ("x_end", "<i4"), ("y_start", "<i4"), ("y_end", "<i4"),
               ("extra_begin", "<i2"), ("extra_end", "<i2"), ("error", "<i2"), ("error_threshold", "<i2"),
               ("cidx", "<u4"), ("clen", "<u4")])
ALN_A = np.dtype([("st", "<i4"), ("align_length", "<u4"), ("rr", "<f8"), ("re", "<i8"),
                  ("w_off", "<u8"), ("w_n", "<u8"), ("c_off", "<u8"), ("c_n", "<u8")])


def ec_align_A(store: Store, rid, chains: np.ndarray, fc: np.ndarray, e_rate=0.04, w_l=775):
    """step A of gen_hc_r_alin for every chain of the read -> (ALN_A[n_ch], WL[], cigar u16[])"""
    chains = np.ascontiguousarray(chains); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1), dtype=np.uint64)
    out = C.c_void_p(); wl = C.c_void_p(); cg = C.c_void_p(); nw = C.c_uint64(); nc = C.c_uint64()
    rc = lib().hao_ec_align_A(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size), C.c_void_p(fc.ctypes.data),
                              C.c_double(e_rate), C.c_int64(w_l), C.byref(out), C.byref(wl), C.byref(nw), C.byref(cg), C.byref(nc))
    assert rc == 0
    return _take(out, chains.size, ALN_A), _take(wl, nw.value, WL), _take(cg, nc.value, np.dtype("<u2"))


ALN_B = np.dtype([("st", "<i4"), ("need_rechain", "<i4"), ("re", "<i8"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                  ("w_off", "<u8"), ("w_n", "<u8"), ("c_off", "<u8"), ("c_n", "<u8")])


def ec_align_B(store: Store, rid, chains, fc, hits, A, WA, e_rate=0.04, w_l=775):
    """step B (base-level CIGAR) for the overlaps step A accepted -> (ALN_B[n_ch], WL[], cigar u16[])"""
    chains = np.ascontiguousarray(chains); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1), dtype=np.uint64)
    hits = np.ascontiguousarray(hits).copy(); A = np.ascontiguousarray(A); WA = np.ascontiguousarray(WA if WA.size else np.zeros(1, WL))
    out = C.c_void_p(); wl = C.c_void_p(); cg = C.c_void_p(); nw = C.c_uint64(); nc = C.c_uint64()
    rc = lib().hao_ec_align_B(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size), C.c_void_p(fc.ctypes.data),
                              C.c_void_p(hits.ctypes.data), C.c_uint64(hits.size), C.c_void_p(A.ctypes.data), C.c_void_p(WA.ctypes.data),
                              C.c_double(e_rate), C.c_int64(w_l), C.byref(out), C.byref(wl), C.byref(nw), C.byref(cg), C.byref(nc))
    assert rc == 0
    return _take(out, chains.size, ALN_B), _take(wl, nw.value, WL), _take(cg, nc.value, np.dtype("<u2"))


ALN_C = np.dtype([("nh_err", "<i8"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                  ("w_off", "<u8"), ("w_n", "<u8"), ("c_off", "<u8"), ("c_n", "<u8")])


def ec_align_C(store: Store, rid, chains, A, B, WB, CB):
    """step C (reassign_gaps) -> (ALN_C[n_ch], WL[], cigar u16[]); cidx relative to the overlap's c_off"""
    chains = np.ascontiguousarray(chains); A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    WB = np.ascontiguousarray(WB if WB.size else np.zeros(1, WL)); CB = np.ascontiguousarray(CB if CB.size else np.zeros(1, np.uint16))
    out = C.c_void_p(); wl = C.c_void_p(); cg = C.c_void_p(); nw = C.c_uint64(); nc = C.c_uint64()
    rc = lib().hao_ec_align_C(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size), C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data),
                              C.c_void_p(WB.ctypes.data), C.c_void_p(CB.ctypes.data), C.byref(out), C.byref(wl), C.byref(nw), C.byref(cg), C.byref(nc))
    assert rc == 0
    return _take(out, chains.size, ALN_C), _take(wl, nw.value, WL), _take(cg, nc.value, np.dtype("<u2"))


PHASE = np.dtype([("y_id", "<u4"), ("rev", "<u4"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"), ("nh_err", "<u4"),
                  ("is_match", "<u4"), ("strong", "<i4")])


def ec_phase(store: Store, rid, chains, A, Cc, WC, CC):
    """rphase_hc + dedup_chains over the accepted overlaps -> (PHASE after phasing, PHASE after dedup)"""
    chains = np.ascontiguousarray(chains); A = np.ascontiguousarray(A); Cc = np.ascontiguousarray(Cc)
    WC = np.ascontiguousarray(WC if WC.size else np.zeros(1, WL)); CC = np.ascontiguousarray(CC if CC.size else np.zeros(1, np.uint16))
    o = C.c_void_p(); d = C.c_void_p(); no = C.c_uint32(); nd = C.c_uint32()
    rc = lib().hao_ec_phase(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size), C.c_void_p(A.ctypes.data), C.c_void_p(Cc.ctypes.data),
                            C.c_void_p(WC.ctypes.data), C.c_void_p(CC.ctypes.data), C.byref(o), C.byref(no), C.byref(d), C.byref(nd))
    assert rc == 0
    a = _take(o, max(no.value, 1), PHASE)[:no.value]; b = _take(d, max(nd.value, 1), PHASE)[:nd.value]
    return a, b


def ec_ea_flags(store: Store, rid, chains, prev: np.ndarray):
    """row a12: which chains gen_hc_r_alin_ea accepts without alignment, given the read's overlap list of the previous round (MA records)"""
    chains = np.ascontiguousarray(chains); prev = np.ascontiguousarray(prev, dtype=MA)
    flag = np.zeros(max(chains.size, 1), np.uint8)
    rc = lib().hao_ec_ea_flags(C.byref(store.c), C.c_uint32(rid), C.c_void_p(chains.ctypes.data), C.c_uint32(chains.size),
                               C.c_void_p((prev if prev.size else np.zeros(1, MA)).ctypes.data), C.c_uint32(prev.size), C.c_void_p(flag.ctypes.data))
    assert rc == 0
    return flag[:chains.size]
