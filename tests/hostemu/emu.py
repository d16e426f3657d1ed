"""ctypes binding of tests/hostemu/libhostemu.so (TEST INFRASTRUCTURE ONLY):
host-compiled bodies of the product's per-thread kernels."""
import ctypes as C
import os
import subprocess

import numpy as np

import ha_oracle as ho

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None
MZ = ho.MZ
HIT = ho.HIT


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhostemu.so")
        srcs = [os.path.join(_HERE, "hostemu.cpp")] + [os.path.join(_ROOT, "hifiasm_b200", "csrc", f)
                                                      for f in os.listdir(os.path.join(_ROOT, "hifiasm_b200", "csrc")) if f.endswith(".cuh")]
        if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unused-function",
                                   "-o", so, srcs[0]])
        L = C.CDLL(so)
        for f in ("emu_reads_create", "emu_ft_create", "emu_pt_create"):
            getattr(L, f).restype = C.c_void_p
        _LIB = L
    return _LIB


def _p(a):
    return C.c_void_p(a.ctypes.data)


class Reads:
    def __init__(self, rs):
        self.keep = [np.ascontiguousarray(x) for x in (rs.length.astype(np.uint64), rs.packed, rs.byte_off.astype(np.uint64),
                                                      rs.n_pos.astype(np.uint64) if rs.n_pos.size else np.zeros(1, np.uint64),
                                                      rs.n_off.astype(np.uint64))]
        ln, pk, bo, npos, noff = self.keep
        self.n = ln.size
        self.length = ln
        self.h = C.c_void_p(lib().emu_reads_create(C.c_uint64(self.n), _p(ln), _p(pk), _p(bo), _p(npos), _p(noff)))


def ft_from_oracle(ft):
    n = int(ho.lib().hao_ft_size(C.c_void_p(ft)))
    key = np.zeros(max(n, 1), np.uint64); val = np.zeros(max(n, 1), np.int32)
    if n:
        ho.lib().hao_ft_dump(C.c_void_p(ft), _p(key), _p(val))
    return C.c_void_p(lib().emu_ft_create(C.c_uint64(n), _p(key), _p(val))), key[:n], val[:n]


def pt_from_oracle(pt):
    nk = int(ho.lib().hao_pt_n_keys(C.c_void_p(pt))); npos = int(ho.lib().hao_pt_tot_pos(C.c_void_p(pt)))
    key = np.zeros(nk + 1, np.uint64); off = np.zeros(nk + 1, np.uint64); cnt = np.zeros(nk + 1, np.uint32); pos = np.zeros(npos + 1, np.uint64)
    ho.lib().hao_pt_dump(C.c_void_p(pt), _p(key), _p(off), _p(cnt), _p(pos))
    return C.c_void_p(lib().emu_pt_create(C.c_uint64(nk), _p(key), _p(off), _p(cnt), C.c_uint64(npos), _p(pos)))


def sketch(reads, ft, p, rid, rid_out=0, two_stage=False):
    cap = int(reads.length[rid]) // 4 + 64
    out = np.zeros(cap, MZ); n = C.c_uint32()
    fn = lib().emu_sketch2 if two_stage else lib().emu_sketch
    ovf = fn(reads.h, ft, C.c_int(int(p["w"])), C.c_int(int(p["k"])), C.c_int(int(p["is_hpc"])),
                           C.c_int(int(p["mz_sample_dist"])), C.c_int(int(p["mz_rewin"])), C.c_uint64(rid), C.c_uint32(rid_out),
                           _p(out), C.c_uint32(cap), C.byref(n))
    assert ovf == 0
    return out[:n.value]


CHAIN = np.dtype([("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_id", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                  ("y_pos_strand", "<u4"), ("shared_seed", "<i4"), ("first_hit", "<u4"), ("n_hits", "<u4"),
                  ("fc_off", "<u4"), ("fc_n", "<u4"), ("pad", "<u4")])
MA = ho.MA


def anchors(reads, pt, mz, high_occ, low_occ):
    L = lib(); L.emu_anchors.restype = C.c_uint64
    mz = np.ascontiguousarray(mz)
    na = L.emu_anchors(reads.h, pt, _p(mz), C.c_uint32(mz.size), C.c_uint32(high_occ), C.c_uint32(low_occ), C.c_void_p(0), C.c_uint64(0))
    out = np.zeros(na + 1, HIT)
    L.emu_anchors(reads.h, pt, _p(mz), C.c_uint32(mz.size), C.c_uint32(high_occ), C.c_uint32(low_occ), _p(out), C.c_uint64(na))
    return out[:na]


def chains(reads, rid, hits, bw, k, max_n_chain):
    hits = np.ascontiguousarray(hits).copy()
    n = hits.size
    out = np.zeros(n + 4, CHAIN); no = C.c_uint32(); ch = np.zeros(n + 4, HIT); nch = C.c_uint64(); fc = np.zeros(3 * n + 16, np.uint64); nfc = C.c_uint64()
    lib().emu_chains(reads.h, C.c_uint32(rid), _p(hits), C.c_uint64(n), C.c_double(bw), C.c_int(k), C.c_int(max_n_chain),
                     _p(out), C.byref(no), _p(ch), C.byref(nch), _p(fc), C.byref(nfc))
    return out[:no.value], ch[:nch.value], fc[:nfc.value], hits


def final_read(reads, ft, pt, p, hom_cov, max_n_chain, rid, in0, in1):
    in0 = np.ascontiguousarray(in0).copy(); in1 = np.ascontiguousarray(in1)
    cap = 4096 + in0.size
    o0 = np.zeros(cap, MA); o1 = np.zeros(cap, MA); m0 = C.c_uint32(); m1 = C.c_uint32()
    rc = lib().emu_final_read(reads.h, ft, pt, C.c_int(int(p["w"])), C.c_int(int(p["k"])), C.c_int(int(p["is_hpc"])),
                              C.c_int(int(p["mz_sample_dist"])), C.c_int(int(p["mz_rewin"])), C.c_int(hom_cov), C.c_int(max_n_chain),
                              C.c_uint32(rid), _p(in0), C.c_uint32(in0.size), _p(in1), C.c_uint32(in1.size),
                              _p(o0), C.byref(m0), _p(o1), C.byref(m1))
    assert rc == 0
    return o0[:m0.value], o1[:m1.value]


ALN = np.dtype([("st", "<i4"), ("align_length", "<u4"), ("rr", "<f8"), ("re", "<i8"), ("w_off", "<u8"), ("w_n", "<u4"), ("pad", "<u4")])
WL = ho.WL


def to_chain(ovlp):
    """oracle chain records (ho.OVLP) -> hb_chain_t"""
    c = np.zeros(ovlp.size, CHAIN)
    for f in ("x_pos_s", "x_pos_e", "y_id", "y_pos_s", "y_pos_e", "y_pos_strand", "shared_seed", "fc_off", "fc_n"):
        c[f] = ovlp[f]
    c["first_hit"] = ovlp["non_homopolymer_errors"]
    return c


def ec_align_A(reads, rid, ch, fc, win, e_rate=0.04, w_l=775):
    """-> (ALN[n_ch], WL[n_win] (an overlap's list at w_off, w_n entries), cigar pool u16[])"""
    ch = np.ascontiguousarray(ch); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1, np.uint64)); win = np.ascontiguousarray(win)
    out = np.zeros(ch.size + 1, ALN); wl = np.zeros(win.size + 1, WL); cap = 64 * win.size + 64; pool = np.zeros(cap, np.uint16); used = C.c_uint64()
    rc = lib().emu_ec_align_A(reads.h, C.c_uint32(rid), _p(ch), C.c_uint32(ch.size), _p(fc), _p(win if win.size else np.zeros(1, ho.WIN)), C.c_uint32(win.size),
                              C.c_double(e_rate), C.c_int32(w_l), _p(out), _p(wl), _p(pool), C.c_uint64(cap), C.byref(used))
    assert rc == 0, rc
    return out[:ch.size], wl[:win.size], pool[:used.value]


ALNB = np.dtype([("st", "<i4"), ("need_rechain", "<i4"), ("re", "<i8"), ("nh_err", "<i8"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                 ("w_off", "<u8"), ("w_n", "<u4"), ("pad", "<u4")])


def ec_align_B(reads, rid, ch, fc, hits, A, WA, e_rate=0.04, w_l=775, path_words=1 << 22, cig_words=1 << 15, gaps=0, poolA=None, rechain=0):
    """-> (rc, ALNB[n_ch], WL[], cigar pool); rc & 1 = some overlap deferred for lack of scratch"""
    ch = np.ascontiguousarray(ch); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1, np.uint64)); hits = np.ascontiguousarray(hits).copy()
    A = np.ascontiguousarray(A); WA = np.ascontiguousarray(WA if WA.size else np.zeros(1, WL))
    out = np.zeros(ch.size + 1, ALNB); cap_w = hits.size + (2 + 8) * ch.size + 16; wl = np.zeros(cap_w, WL)
    cap = 1 << 22; pool = np.zeros(cap, np.uint16); used = C.c_uint64(); nw = C.c_uint64()
    rc = lib().emu_ec_align_B(reads.h, C.c_uint32(rid), _p(ch), C.c_uint32(ch.size), _p(fc), _p(hits), C.c_uint64(hits.size), _p(A), _p(WA),
                              C.c_double(e_rate), C.c_int32(w_l), C.c_int32(gaps), C.c_uint64(path_words), C.c_int32(cig_words),
                              _p(out), _p(wl), C.c_uint64(cap_w), _p(pool), C.c_uint64(cap), C.byref(used), C.byref(nw),
                              _p(np.ascontiguousarray(poolA if poolA is not None and poolA.size else np.zeros(1, np.uint16))), C.c_int32(rechain))
    return rc, out[:ch.size], wl[:nw.value], pool[:used.value]


def ec_align_B_par(reads, rid, ch, fc, hits, A, WA, e_rate=0.04, w_l=775, gaps=0, cig_words=1 << 16, poolA=None, rechain=0):
    """step B (+C) the way the GPU pipeline runs it: prep, independent segments with tiered scratch, merge.
    -> (rc, ALNB[n_ch], WL[], cigar pool, (segments computed in tier 0, recomputed in the large tier))"""
    ch = np.ascontiguousarray(ch); fc = np.ascontiguousarray(fc if fc.size else np.zeros(1, np.uint64)); hits = np.ascontiguousarray(hits).copy()
    A = np.ascontiguousarray(A); WA = np.ascontiguousarray(WA if WA.size else np.zeros(1, WL))
    out = np.zeros(ch.size + 1, ALNB); cap_w = hits.size + (2 + 8) * ch.size + 16; wl = np.zeros(cap_w, WL)
    cap = 1 << 22; pool = np.zeros(cap, np.uint16); used = C.c_uint64(); nw = C.c_uint64(); nt = (C.c_uint64 * 2)()
    rc = lib().emu_ec_align_B_par(reads.h, C.c_uint32(rid), _p(ch), C.c_uint32(ch.size), _p(fc), _p(hits), C.c_uint64(hits.size), _p(A), _p(WA),
                                  C.c_double(e_rate), C.c_int32(w_l), C.c_int32(gaps), C.c_int32(cig_words),
                                  _p(out), _p(wl), C.c_uint64(cap_w), _p(pool), C.c_uint64(cap), C.byref(used), C.byref(nw), nt,
                                  _p(np.ascontiguousarray(poolA if poolA is not None and poolA.size else np.zeros(1, np.uint16))), C.c_int32(rechain))
    return rc, out[:ch.size], wl[:nw.value], pool[:used.value], (nt[0], nt[1])


def ec_phase(reads, rid, ch, A, B, WB, CB):
    """rphase_hc over the accepted overlaps (B = steps B + C output) -> (is_match u8[], strong i8[]) per accepted overlap"""
    ch = np.ascontiguousarray(ch); A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    WB = np.ascontiguousarray(WB if WB.size else np.zeros(1, WL)); CB = np.ascontiguousarray(CB if CB.size else np.zeros(1, np.uint16))
    im = np.zeros(ch.size + 1, np.uint8); st = np.zeros(ch.size + 1, np.int8); n = C.c_uint32()
    rc = lib().emu_ec_phase(reads.h, C.c_uint32(rid), _p(ch), C.c_uint32(ch.size), _p(A), _p(B), _p(WB), _p(CB), _p(im), _p(st), C.byref(n))
    assert rc == 0
    return im[:n.value], st[:n.value]


PHASE = np.dtype([("st", "<i4"), ("y_id", "<u4"), ("rev", "<u4"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"), ("nh_err", "<u4"),
                  ("is_match", "<u4"), ("strong", "<i4"), ("need_rechain", "<u4"), ("pad", "<u4")])


def ec_reverse(reads, rid, ph):
    """dedup_chains + push_ne_ovlp(flag 2) over the phased overlaps (PHASE records) -> MA records"""
    ph = np.ascontiguousarray(ph, dtype=PHASE); out = np.zeros(ph.size + 1, MA); n = C.c_uint32()
    rc = lib().emu_ec_reverse(reads.h, C.c_uint32(rid), _p(ph if ph.size else np.zeros(1, PHASE)), C.c_uint32(ph.size), _p(out), C.byref(n))
    assert rc == 0
    return out[:n.value]


# ---- closing steps of an EC round (hb_ecround.cuh: rows a15-a18) ----
class DerivedReads:
    """a read store made by emu_ec_apply / emu_ec_rc (owned by the library)"""
    def __init__(self, h, n):
        self.h = C.c_void_p(h); self.n = n
        ln = np.zeros(n, np.uint64)
        lib().emu_reads_export(self.h, _p(ln), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0))
        self.length = ln

    def export(self):
        """-> binio.ReadStore in the All_reads layout"""
        from hifiasm_b200 import binio
        L = lib(); L.emu_reads_n_npos.restype = C.c_uint64
        nb = self.length // 4 + 1
        boff = np.zeros(self.n + 1, np.uint64); np.cumsum(nb, out=boff[1:])
        packed = np.zeros(int(boff[-1]) + 1, np.uint8); noff = np.zeros(self.n + 1, np.uint64); npos = np.zeros(int(L.emu_reads_n_npos(self.h)) + 1, np.uint64)
        L.emu_reads_export(self.h, C.c_void_p(0), _p(packed), _p(noff), _p(npos))
        return binio.ReadStore(length=self.length.copy(), byte_off=boff, packed=packed[:int(boff[-1])], n_off=noff, n_pos=npos[:int(noff[-1])])


def ec_apply(reads, scc, scc_off):
    L = lib(); L.emu_ec_apply.restype = C.c_void_p
    scc = np.ascontiguousarray(scc if scc.size else np.zeros(1, np.uint16)); scc_off = np.ascontiguousarray(scc_off, dtype=np.uint64)
    return DerivedReads(L.emu_ec_apply(reads.h, _p(scc), _p(scc_off)), reads.n)


def ec_rc(reads):
    L = lib(); L.emu_ec_rc.restype = C.c_void_p
    return DerivedReads(L.emu_ec_rc(reads.h), reads.n)


def ec_update(reads, paf, off, scc, scc_off):
    """worker_update_dc_ec over every list -> (updated MA records, number of exact records)"""
    L = lib(); L.emu_ec_update.restype = C.c_uint64
    paf = np.ascontiguousarray(paf, dtype=MA).copy(); off = np.ascontiguousarray(off, dtype=np.uint64)
    scc = np.ascontiguousarray(scc if scc.size else np.zeros(1, np.uint16)); scc_off = np.ascontiguousarray(scc_off, dtype=np.uint64)
    ne = L.emu_ec_update(reads.h, _p(paf if paf.size else np.zeros(1, MA)), _p(off), _p(scc), _p(scc_off))
    return paf, int(ne)


def ec_flip(reads, paf, off):
    """flip_paf_rc over every list -> (records, off) compacted"""
    paf = np.ascontiguousarray(paf, dtype=MA).copy(); off = np.ascontiguousarray(off, dtype=np.uint64)
    n = off.size - 1; no = np.zeros(n + 1, np.uint32)
    lib().emu_ec_flip(reads.h, _p(paf if paf.size else np.zeros(1, MA)), _p(off), _p(no))
    keep = np.zeros(paf.size, bool); noff = np.zeros(n + 1, np.uint64)
    for i in range(n):
        keep[int(off[i]):int(off[i]) + int(no[i])] = True
        noff[i + 1] = noff[i] + int(no[i])
    return paf[keep], noff


def ec_source(reads, rid, ph, alnb, wl, pool, ec):
    """push_ne_ovlp(flag 1, ec) + check_well_cal -> (MA records, is_fully_corrected, is_abnormal); ph (PHASE) / alnb (ALNB) parallel, one per overlap"""
    ph = np.ascontiguousarray(ph, dtype=PHASE); alnb = np.ascontiguousarray(alnb, dtype=ALNB); assert ph.size == alnb.size
    wl = np.ascontiguousarray(wl if wl.size else np.zeros(1, WL)); pool = np.ascontiguousarray(pool if pool.size else np.zeros(1, np.uint16))
    ec = np.ascontiguousarray(ec, dtype=np.uint16); out = np.zeros(ph.size + 1, MA); n = C.c_uint32(); fl = np.zeros(2, np.uint8)
    rc = lib().emu_ec_source(reads.h, C.c_uint32(rid), _p(ph if ph.size else np.zeros(1, PHASE)), _p(alnb if alnb.size else np.zeros(1, ALNB)), C.c_uint32(ph.size), _p(wl), _p(pool),
                             _p(ec if ec.size else np.zeros(1, np.uint16)), C.c_uint64(ec.size), _p(out), C.byref(n), _p(fl))
    assert rc == 0
    return out[:n.value], int(fl[0]), int(fl[1])


def ec_cns(reads, rid, ph, alnb, wl, pool, cap=1 << 16, g_nodes=0, g_arcs=0):
    """wcns_gen -> (status, edit script u16[], corrected bases); status 0 = done, 1 = the read needs the graph consensus and no arena was given
    (g_nodes = 0: the first launch of the GPU path), 3 = the arena (g_nodes nodes, g_arcs arcs) was too small"""
    ph = np.ascontiguousarray(ph, dtype=PHASE); alnb = np.ascontiguousarray(alnb, dtype=ALNB); assert ph.size == alnb.size
    wl = np.ascontiguousarray(wl if wl.size else np.zeros(1, WL)); pool = np.ascontiguousarray(pool if pool.size else np.zeros(1, np.uint16))
    out = np.zeros(cap, np.uint16); n = C.c_uint32(); nec = C.c_uint64()
    rc = lib().emu_ec_cns(reads.h, C.c_uint32(rid), _p(ph if ph.size else np.zeros(1, PHASE)), _p(alnb if alnb.size else np.zeros(1, ALNB)), C.c_uint32(ph.size), _p(wl), _p(pool),
                          _p(out), C.c_uint32(cap), C.byref(n), C.byref(nec), C.c_uint32(g_nodes), C.c_uint32(g_arcs))
    assert rc in (0, 1, 3), rc
    return rc, out[:n.value], int(nec.value)


def bf_counts(hashes, bf_shift):
    """counting behind the reference's Bloom filter as index.cu evaluates it -> (keys, counts) of the k-mers in the table, by key"""
    h = np.ascontiguousarray(hashes, dtype=np.uint64); L = lib(); L.emu_bf_counts.restype = C.c_uint64
    key = np.zeros(h.size + 1, np.uint64); cnt = np.zeros(h.size + 1, np.uint32)
    m = L.emu_bf_counts(_p(h), C.c_uint64(h.size), C.c_int(bf_shift), _p(key), _p(cnt), C.c_uint64(h.size))
    return key[:m], cnt[:m]


def mw_align(reads, qid, tid, rev, mode, ps0, pn, qs0, tn, thre, abs_diag, which):
    """one call of the step-B aligner (which = 0: hb_mw_align, 1: hb_mw_align_w over virtual lanes) -> (ovf, (err, ps, pe, ts, te), cigar u16[])"""
    res = np.zeros(6, np.int32); cig = np.zeros(1 << 16, np.uint16)
    ovf = lib().emu_mw_align(reads.h, C.c_uint32(qid), C.c_uint32(tid), C.c_uint32(rev), C.c_int(mode), C.c_int64(ps0), C.c_int32(pn), C.c_int64(qs0), C.c_int32(tn),
                             C.c_int32(thre), C.c_int32(abs_diag), C.c_int(which), _p(res), _p(cig), C.c_int32(cig.size))
    return ovf, tuple(int(x) for x in res[:5]), cig[:int(res[5])].copy()
