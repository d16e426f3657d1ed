"""ctypes mirror of include/hifiasm_b200.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import binio

_HERE = os.path.dirname(os.path.abspath(__file__))
MZ = np.dtype([("x", "<u8"), ("info", "<u8")])
HIT = np.dtype([("id_strand", "<u4"), ("offset", "<u4"), ("self_offset", "<u4"), ("cnt", "<u4")])
CHAIN = np.dtype([("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_id", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                  ("y_pos_strand", "<u4"), ("shared_seed", "<i4"), ("first_hit", "<u4"), ("n_hits", "<u4"),
                  ("fc_off", "<u4"), ("fc_n", "<u4"), ("pad", "<u4")])
MA = binio.MA_MEM
WIN = np.dtype([(f, "<i4") for f in ("chain", "q_s", "q_e", "t_s", "t_pri_l", "thre", "aux_beg", "aux_end", "err", "pe")])


WL = np.dtype([("x_start", "<i4"), ("x_end", "<i4"), ("y_start", "<i4"), ("y_end", "<i4"),
               ("extra_begin", "<i2"), ("extra_end", "<i2"), ("error", "<i2"), ("error_threshold", "<i2"),
               ("cidx", "<u4"), ("clen", "<u4")])  # window_list, Hash_Table.h:54-62
ALN = np.dtype([("st", "<i4"), ("align_length", "<u4"), ("rr", "<f8"), ("re", "<i8"), ("w_off", "<u8"), ("w_n", "<u4"), ("pad", "<u4")])


ALNB = np.dtype([("st", "<i4"), ("need_rechain", "<i4"), ("re", "<i8"), ("nh_err", "<i8"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"),
                 ("w_off", "<u8"), ("w_n", "<u4"), ("pad", "<u4")])


PHASE = np.dtype([("st", "<i4"), ("y_id", "<u4"), ("rev", "<u4"), ("x_pos_s", "<u4"), ("x_pos_e", "<u4"), ("y_pos_s", "<u4"), ("y_pos_e", "<u4"), ("nh_err", "<u4"),
                  ("is_match", "<u4"), ("strong", "<i4"), ("need_rechain", "<u4"), ("pad", "<u4")])



class StageResult(C.Structure):
    """hb_stage_result_t"""
    _fields_ = [("n_reads", C.c_uint64), ("n_src", C.c_uint64), ("n_rev", C.c_uint64), ("src", C.c_void_p), ("rev", C.c_void_p), ("src_off", C.c_void_p), ("rev_off", C.c_void_p),
                ("is_fully_corrected", C.c_void_p), ("is_abnormal", C.c_void_p), ("hom_cov", C.c_int32), ("het_cov", C.c_int32), ("corrected_bases", C.c_uint64 * 8), ("n_unfinished", C.c_uint64),
                ("ms_ft", C.c_double), ("ms_pt", C.c_double * 9), ("ms_ec", C.c_double * 8), ("ms_final", C.c_double), ("ms_exchange", C.c_double), ("ms_total", C.c_double), ("device_ms", C.c_double)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)


def torch_allgather(device):
    """hb_allgather_fn over torch.distributed (NCCL when `device` is a CUDA device, gloo on the CPU): fixed-size byte strings in host memory"""
    import torch
    import torch.distributed as dist

    def cb(user, send, recv, nbytes):
        try:
            world = dist.get_world_size()
            a = torch.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=torch.uint8)
            r = torch.frombuffer((C.c_uint8 * (world * nbytes)).from_address(recv), dtype=torch.uint8)
            if device is not None and str(device) != "cpu":   # NCCL: the library's buffers are page-locked, so both copies are plain DMA
                out = torch.empty(world * nbytes, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(out, a.to(device, non_blocking=True))
                r.copy_(out)
            else:
                parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
                dist.all_gather(parts, a.clone())
                r.copy_(torch.cat(parts))
            return 0
        except Exception as ex:  # the C side turns a non-zero return into an error of the call
            import sys
            sys.stderr.write("[hifiasm_b200] all-gather callback failed: %r\n" % (ex,))
            return 1
    return ALLGATHER_FN(cb)


class HBError(RuntimeError):
    code = 0                                                                          # the HB_E_* code of include/hifiasm_b200.h, when the error came from the library


class Opt(C.Structure):
    _fields_ = [("k_mer_length", C.c_int32), ("mz_win", C.c_int32), ("is_hpc", C.c_int32), ("mz_sample_dist", C.c_int32),
                ("mz_rewin", C.c_int32), ("min_hist_kmer_cnt", C.c_int32), ("max_kmer_cnt", C.c_int32), ("max_n_chain", C.c_int32),
                ("high_factor", C.c_double), ("hom_cov", C.c_int32), ("het_cov", C.c_int32), ("is_ont", C.c_int32), ("bf_shift", C.c_int32)]


def lib_path() -> str:
    return os.path.join(_HERE, "libhifiasm_b200.so")


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise HBError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % p)
        L = C.CDLL(p)
        L.hb_last_error.restype = C.c_char_p
        L.hb_device_count.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Engine:
    """One context per GPU (include/hifiasm_b200.h: hb_create)."""

    def __init__(self, device: int = 0, **opt):
        L = _lib()
        self.opt = Opt()
        L.hb_opt_init(C.byref(self.opt))
        for k, v in opt.items():
            setattr(self.opt, k, v)
        self.h = C.c_void_p()
        rc = L.hb_create(C.byref(self.h), C.c_int(device), C.byref(self.opt))
        if rc != 0:
            raise HBError("hb_create failed (%d): no CUDA device / bad options; this engine has no CPU path" % rc)
        self.n_reads = 0

    def close(self):
        if self.h:
            _lib().hb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            e = HBError("hifiasm_b200 error %d: %s" % (rc, _lib().hb_last_error(self.h).decode()))
            e.code = rc
            raise e

    def set_opt(self, **kw):
        for k, v in kw.items():
            setattr(self.opt, k, v)
        self._ck(_lib().hb_set_opt(self.h, C.byref(self.opt)))

    def get_opt(self):
        self._ck(_lib().hb_get_opt(self.h, C.byref(self.opt)))
        return self.opt

    def update_cov(self, hom):
        _lib().hb_opt_update_cov(C.byref(self.opt), C.c_int(hom))
        self._ck(_lib().hb_set_opt(self.h, C.byref(self.opt)))

    # ---- read store
    def upload_reads(self, length, packed, byte_off, n_pos=None, n_off=None):
        length = np.ascontiguousarray(length, dtype=np.uint64); packed = np.ascontiguousarray(packed, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        if n_off is None:
            n_off = np.zeros(length.size + 1, np.uint64); n_pos = np.zeros(1, np.uint64)
        n_pos = np.ascontiguousarray(n_pos if n_pos.size else np.zeros(1), dtype=np.uint64); n_off = np.ascontiguousarray(n_off, dtype=np.uint64)
        self._ck(_lib().hb_reads_upload(self.h, C.c_uint64(length.size), _p(length), _p(packed), _p(byte_off), _p(n_pos), _p(n_off)))
        self.n_reads = int(length.size)

    def upload_store(self, rs: binio.ReadStore):
        self.upload_reads(rs.length, rs.packed, rs.byte_off, rs.n_pos, rs.n_off)

    # ---- index
    def ft_gen(self) -> int:
        hom = C.c_int()
        self._ck(_lib().hb_ft_gen(self.h, C.byref(hom)))
        return hom.value

    def ft_size(self) -> int:
        n = C.c_uint64(); _lib().hb_ft_size(self.h, C.byref(n)); return n.value

    def ft_cnt(self, hashes):
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64); out = np.zeros(hashes.size, np.int32)
        self._ck(_lib().hb_ft_cnt(self.h, _p(hashes), C.c_uint64(hashes.size), _p(out)))
        return out

    def pt_gen(self):
        hom = C.c_int(); het = C.c_int()
        self._ck(_lib().hb_pt_gen(self.h, C.byref(hom), C.byref(het)))
        return hom.value, het.value

    def pt_stat(self):
        a = C.c_uint64(); b = C.c_uint64(); _lib().hb_pt_stat(self.h, C.byref(a), C.byref(b)); return a.value, b.value

    def pt_get(self, hashes):
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64); cnt = np.zeros(hashes.size, np.uint32)
        self._ck(_lib().hb_pt_get(self.h, _p(hashes), C.c_uint64(hashes.size), _p(cnt), C.c_void_p(0), C.c_uint64(0)))
        tot = int(cnt.sum()); pos = np.zeros(tot + 1, np.uint64)
        self._ck(_lib().hb_pt_get(self.h, _p(hashes), C.c_uint64(hashes.size), _p(cnt), _p(pos), C.c_uint64(tot)))
        return cnt, pos[:tot]

    # ---- stages
    def sketch(self, r0, r1):
        off = np.zeros(r1 - r0 + 1, np.uint64)
        self._ck(_lib().hb_sketch(self.h, C.c_uint64(r0), C.c_uint64(r1), _p(off), C.c_void_p(0), C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, MZ)
        self._ck(_lib().hb_sketch(self.h, C.c_uint64(r0), C.c_uint64(r1), _p(off), _p(rec), C.c_uint64(int(off[-1]))))
        return off, rec[:int(off[-1])]

    def anchors(self, r0, r1):
        off = np.zeros(r1 - r0 + 1, np.uint64)
        self._ck(_lib().hb_anchors(self.h, C.c_uint64(r0), C.c_uint64(r1), _p(off), C.c_void_p(0), C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, HIT)
        self._ck(_lib().hb_anchors(self.h, C.c_uint64(r0), C.c_uint64(r1), _p(off), _p(rec), C.c_uint64(int(off[-1]))))
        return off, rec[:int(off[-1])]

    def chains(self, r0, r1, bw):
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64); hoff = np.zeros(n + 1, np.uint64); foff = np.zeros(n + 1, np.uint64)
        z = C.c_void_p(0)
        self._ck(_lib().hb_chains(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), _p(off), z, C.c_uint64(0), _p(hoff), z, C.c_uint64(0), _p(foff), z, C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, CHAIN); hits = np.zeros(int(hoff[-1]) + 1, HIT); fc = np.zeros(int(foff[-1]) + 1, np.uint64)
        self._ck(_lib().hb_chains(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), _p(off), _p(rec), C.c_uint64(rec.size),
                                  _p(hoff), _p(hits), C.c_uint64(hits.size), _p(foff), _p(fc), C.c_uint64(fc.size)))
        return off, rec[:int(off[-1])], hoff, hits[:int(hoff[-1])], foff, fc[:int(foff[-1])]

    def windows(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775):
        """window pass of an EC round (row a8) -> (off, records WIN)"""
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64)
        self._ck(_lib().hb_windows(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), C.c_void_p(0), C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, WIN)
        self._ck(_lib().hb_windows(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), _p(rec), C.c_uint64(rec.size)))
        return off, rec[:int(off[-1])]

    def ec_align(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775):
        """step A of the alignment stage of an EC round (rows a8 + a9): -> (off, ALN records, WL window lists, cigar pool)"""
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64); nw = C.c_uint64(); nc = C.c_uint64(); z = C.c_void_p(0)
        a = (self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off))
        self._ck(_lib().hb_ec_align(*a, z, C.c_uint64(0), z, C.c_uint64(0), z, C.c_uint64(0), C.byref(nw), C.byref(nc)))
        rec = np.zeros(int(off[-1]) + 1, ALN); wl = np.zeros(nw.value + 1, WL); cig = np.zeros(2 * nc.value + 4096, np.uint16)
        self._ck(_lib().hb_ec_align(*a, _p(rec), C.c_uint64(rec.size), _p(wl), C.c_uint64(wl.size), _p(cig), C.c_uint64(cig.size), C.byref(nw), C.byref(nc)))
        return off, rec[:int(off[-1])], wl[:nw.value], cig[:nc.value]

    def ec_stage_prev(self, prev_src, prev_src_off):
        """stage R_INF.paf[] of the previous EC round (MA records + offsets) for the exact shortcut of row a12 (ec_cigar gaps & 2)"""
        prev_src = np.ascontiguousarray(prev_src, dtype=MA); o = np.ascontiguousarray(prev_src_off, dtype=np.uint64)
        self._ck(_lib().hb_ec_stage_prev(self.h, _p(prev_src if prev_src.size else np.zeros(1, MA)), _p(o)))

    def ec_cigar(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775, gaps=0):
        """steps A + B (+ C when gaps) of the alignment stage of an EC round (rows a8-a11): base-level CIGARs of the accepted overlaps
        -> (off, ALNB records, WL window lists, cigar pool)"""
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64); nw = C.c_uint64(); nc = C.c_uint64(); z = C.c_void_p(0)
        a = (self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), C.c_int32(gaps), _p(off))
        self._ck(_lib().hb_ec_cigar(*a, z, C.c_uint64(0), z, C.c_uint64(0), z, C.c_uint64(0), C.byref(nw), C.byref(nc)))
        rec = np.zeros(int(off[-1]) + 1, ALNB); wl = np.zeros(nw.value + 1, WL); cig = np.zeros(2 * nc.value + 4096, np.uint16)
        self._ck(_lib().hb_ec_cigar(*a, _p(rec), C.c_uint64(rec.size), _p(wl), C.c_uint64(wl.size), _p(cig), C.c_uint64(cig.size), C.byref(nw), C.byref(nc)))
        return off, rec[:int(off[-1])], wl[:nw.value], cig[:nc.value]

    def ec_phase(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775):
        """alignment stage (rows a8-a11) + phasing (row a13) of an EC round -> (off, PHASE records, one per chain)"""
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64)
        self._ck(_lib().hb_ec_phase(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), C.c_void_p(0), C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, PHASE)
        self._ck(_lib().hb_ec_phase(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), _p(rec), C.c_uint64(rec.size)))
        return off, rec[:int(off[-1])]

    def ec_reverse_paf(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775):
        """R_INF.reverse_paf[i] of an EC round (other-haplotype overlaps after dedup_chains) -> (off, MA records)"""
        n = r1 - r0
        off = np.zeros(n + 1, np.uint64)
        self._ck(_lib().hb_ec_reverse_paf(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), C.c_void_p(0), C.c_uint64(0)))
        rec = np.zeros(int(off[-1]) + 1, MA)
        self._ck(_lib().hb_ec_reverse_paf(self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), _p(off), _p(rec), C.c_uint64(rec.size)))
        return off, rec[:int(off[-1])]

    # ---- the rest of an EC round (rows a15-a18)
    def ec_stage_scc(self, scc, scc_off):
        """stage the round's edit scripts (scc.a[i]: uint16 runs, offsets n_reads + 1)"""
        scc = np.ascontiguousarray(scc, dtype=np.uint16); o = np.ascontiguousarray(scc_off, dtype=np.uint64)
        assert o.size == self.n_reads + 1
        self._ck(_lib().hb_ec_stage_scc(self.h, _p(scc if scc.size else np.zeros(1, np.uint16)), _p(o)))

    def ec_round_lists(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775, use_prev=0):
        """what worker_hap_ec leaves in paf[i] / reverse_paf[i] -> (src_off, src MA, rev_off, rev MA, is_fully_corrected u8[], is_abnormal u8[])"""
        n = r1 - r0
        so = np.zeros(n + 1, np.uint64); ro = np.zeros(n + 1, np.uint64); fl = np.zeros(2 * n + 2, np.uint8); z = C.c_void_p(0)
        a = (self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), C.c_int32(use_prev))
        self._ck(_lib().hb_ec_round_lists(*a, _p(so), z, C.c_uint64(0), _p(ro), z, C.c_uint64(0), _p(fl)))
        src = np.zeros(int(so[-1]) + 1, MA); rev = np.zeros(int(ro[-1]) + 1, MA)
        self._ck(_lib().hb_ec_round_lists(*a, _p(so), _p(src), C.c_uint64(src.size), _p(ro), _p(rev), C.c_uint64(rev.size), _p(fl)))
        return so, src[:int(so[-1])], ro, rev[:int(ro[-1])], fl[0:2 * n:2].copy(), fl[1:2 * n:2].copy()

    def ec_round(self, r0, r1, bw=0.02, e_rate=0.04, w_l=775, use_prev=0, caps=None):
        """rows a14 + a15: consensus on the device -> dict(src_off, src, rev_off, rev, is_fully_corrected, is_abnormal, scc_off, scc, status, n_corrected).
        caps = (paf records, reverse_paf records, edit-script words): one pass with these capacities; None: a sizing pass first"""
        n = r1 - r0
        so = np.zeros(n + 1, np.uint64); ro = np.zeros(n + 1, np.uint64); co = np.zeros(n + 1, np.uint64); fl = np.zeros(2 * n + 2, np.uint8); st = np.zeros(n + 1, np.uint8); z = C.c_void_p(0); nc = C.c_uint64()
        a = (self.h, C.c_uint64(r0), C.c_uint64(r1), C.c_double(bw), C.c_double(e_rate), C.c_int32(w_l), C.c_int32(use_prev))
        if caps is None:
            self._ck(_lib().hb_ec_round(*a, _p(so), z, C.c_uint64(0), _p(ro), z, C.c_uint64(0), _p(fl), _p(co), z, C.c_uint64(0), _p(st), C.byref(nc)))
            caps = (int(so[-1]), int(ro[-1]), int(co[-1]))
        src = np.zeros(caps[0] + 1, MA); rev = np.zeros(caps[1] + 1, MA); scc = np.zeros(caps[2] + 1, np.uint16)
        self._ck(_lib().hb_ec_round(*a, _p(so), _p(src), C.c_uint64(src.size), _p(ro), _p(rev), C.c_uint64(rev.size), _p(fl), _p(co), _p(scc), C.c_uint64(scc.size), _p(st), C.byref(nc)))
        return dict(src_off=so, src=src[:int(so[-1])], rev_off=ro, rev=rev[:int(ro[-1])], is_fully_corrected=fl[0:2 * n:2].copy(), is_abnormal=fl[1:2 * n:2].copy(),
                    scc_off=co, scc=scc[:int(co[-1])], status=st[:n].copy(), n_corrected=int(nc.value))

    def cal_ec_r(self, round_, is_sv, prev_src, prev_src_off, e_rate=0.04, w_l=775, cap=None):
        """cal_ec_r as one call -> dict(src, src_off, rev, rev_off, is_fully_corrected, is_abnormal, status, tot_b, tot_e, n_exact, n_inexact)"""
        n = self.n_reads
        prev_src = np.ascontiguousarray(prev_src, dtype=MA); po = np.ascontiguousarray(prev_src_off, dtype=np.uint64)
        cap = cap or (2 * prev_src.size + 256 * n + 1024)
        src = np.zeros(cap, MA); rev = np.zeros(cap, MA); so = np.zeros(n + 1, np.uint64); ro = np.zeros(n + 1, np.uint64); fl = np.zeros(2 * n + 2, np.uint8); st = np.zeros(n + 1, np.uint8)
        tb = C.c_uint64(); te = C.c_uint64(); ne = C.c_uint64(); ni = C.c_uint64()
        self._ck(_lib().hb_cal_ec_r(self.h, C.c_uint64(round_), C.c_uint64(0), C.c_uint64(is_sv), C.c_double(e_rate), C.c_int32(w_l), _p(prev_src if prev_src.size else np.zeros(1, MA)), _p(po),
                                    _p(src), _p(so), C.c_uint64(cap), _p(rev), _p(ro), C.c_uint64(cap), _p(fl), _p(st), C.byref(tb), C.byref(te), C.byref(ne), C.byref(ni)))
        return dict(src=src[:int(so[-1])], src_off=so, rev=rev[:int(ro[-1])], rev_off=ro, is_fully_corrected=fl[0:2 * n:2].copy(), is_abnormal=fl[1:2 * n:2].copy(), status=st[:n].copy(),
                    tot_b=int(tb.value), tot_e=int(te.value), n_exact=int(ne.value), n_inexact=int(ni.value))

    def ec_apply(self):
        """sl_ec_r: apply the staged edit scripts to the resident reads -> (reads changed, total bases)"""
        a = C.c_uint64(); b = C.c_uint64()
        self._ck(_lib().hb_ec_apply(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ec_update_paf(self, paf, off):
        """cal_update_ec_multiple on the corrected store -> (updated MA records, exact, inexact)"""
        paf = np.ascontiguousarray(paf, dtype=MA).copy(); off = np.ascontiguousarray(off, dtype=np.uint64)
        a = C.c_uint64(); b = C.c_uint64()
        self._ck(_lib().hb_ec_update_paf(self.h, _p(paf if paf.size else np.zeros(1, MA)), _p(off), C.byref(a), C.byref(b)))
        return paf, a.value, b.value

    def ec_post_rev(self, paf, off, rpaf, roff):
        """worker_hap_post_rev: reverse-complement the resident reads, flip both lists -> (paf, off, rpaf, roff)"""
        paf = np.ascontiguousarray(paf, dtype=MA).copy(); off = np.ascontiguousarray(off, dtype=np.uint64).copy()
        rpaf = np.ascontiguousarray(rpaf, dtype=MA).copy(); roff = np.ascontiguousarray(roff, dtype=np.uint64).copy()
        self._ck(_lib().hb_ec_post_rev(self.h, _p(paf if paf.size else np.zeros(1, MA)), _p(off), _p(rpaf if rpaf.size else np.zeros(1, MA)), _p(roff)))
        return paf[:int(off[-1])], off, rpaf[:int(roff[-1])], roff

    def read_lengths(self):
        """lengths of the resident reads (u64[n])"""
        n = self.n_reads
        ln = np.zeros(n, np.uint64); noff = np.zeros(n + 1, np.uint64); z = C.c_void_p(0)
        self._ck(_lib().hb_reads_download(self.h, _p(ln), z, C.c_uint64(0), _p(noff), z, C.c_uint64(0)))
        return ln

    def download_reads(self):
        """the resident read store -> binio.ReadStore (All_reads layout)"""
        from . import binio
        n = self.n_reads
        ln = np.zeros(n, np.uint64); noff = np.zeros(n + 1, np.uint64); z = C.c_void_p(0)
        self._ck(_lib().hb_reads_download(self.h, _p(ln), z, C.c_uint64(0), _p(noff), z, C.c_uint64(0)))
        boff = np.zeros(n + 1, np.uint64); np.cumsum(ln // 4 + 1, out=boff[1:])
        packed = np.zeros(int(boff[-1]) + 1, np.uint8); npos = np.zeros(int(noff[-1]) + 1, np.uint64)
        self._ck(_lib().hb_reads_download(self.h, z, _p(packed), C.c_uint64(int(boff[-1])), z, _p(npos), C.c_uint64(int(noff[-1]))))
        return binio.ReadStore(length=ln, byte_off=boff, packed=packed[:int(boff[-1])], n_off=noff, n_pos=npos[:int(noff[-1])])

    # ---- final pass
    def cal_ov_r(self, prev_src, prev_src_off, prev_rev, prev_rev_off, r0=0, r1=None, cap=None, out=None):
        """out = (out0, out1) preallocated MA arrays to receive the records (reused across calls)"""
        r1 = self.n_reads if r1 is None else r1
        prev_src = np.ascontiguousarray(prev_src, dtype=MA); prev_rev = np.ascontiguousarray(prev_rev, dtype=MA)
        o0 = np.ascontiguousarray(prev_src_off, dtype=np.uint64); o1 = np.ascontiguousarray(prev_rev_off, dtype=np.uint64)
        cap = cap or (4 * (prev_src.size + prev_rev.size) + 256 * (r1 - r0) + 1024)
        if out is not None:
            out0, out1 = out
        else:
            out0 = np.empty(cap, MA); out1 = np.empty(cap, MA)
        oo0 = np.zeros(r1 - r0 + 1, np.uint64); oo1 = np.zeros(r1 - r0 + 1, np.uint64); stat = np.zeros(8, np.uint64)
        self._ck(_lib().hb_cal_ov_r(self.h, C.c_uint64(r0), C.c_uint64(r1), _p(prev_src), _p(o0), _p(prev_rev), _p(o1),
                                    _p(out0), _p(oo0), C.c_uint64(out0.size), _p(out1), _p(oo1), C.c_uint64(out1.size), _p(stat)))
        return out0[:int(oo0[-1])], oo0, out1[:int(oo1[-1])], oo1, stat

    def cal_ov_r_resident(self, r0=0, r1=None):
        r1 = self.n_reads if r1 is None else r1
        a = C.c_uint64(); b = C.c_uint64(); stat = np.zeros(8, np.uint64)
        self._ck(_lib().hb_cal_ov_r_resident(self.h, C.c_uint64(r0), C.c_uint64(r1), C.byref(a), C.byref(b), _p(stat)))
        return a.value, b.value, stat


    # ---- the whole stage as one call
    def stage_run(self, n_round=3, rank=0, world=1, allgather=None, copy=True):
        """hb_stage_run -> dict(src, src_off, rev, rev_off, is_fully_corrected, is_abnormal, hom_cov, het_cov, corrected_bases, n_unfinished, ms = per-step host wall clock, device_ms)"""
        res = StageResult(); n = self.n_reads
        cb = allgather if allgather is not None else C.cast(None, ALLGATHER_FN)
        self._ck(_lib().hb_stage_run(self.h, C.c_int(n_round), C.c_int(rank), C.c_int(world), cb, C.c_void_p(0), C.byref(res)))
        def arr(ptr, cnt, dt):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(1, cnt) * np.dtype(dt).itemsize,)).view(dt)[:cnt]
            return a.copy() if copy else a
        return dict(src=arr(res.src, res.n_src, MA), src_off=arr(res.src_off, n + 1, np.uint64), rev=arr(res.rev, res.n_rev, MA), rev_off=arr(res.rev_off, n + 1, np.uint64),
                    is_fully_corrected=arr(res.is_fully_corrected, n, np.uint8), is_abnormal=arr(res.is_abnormal, n, np.uint8), hom_cov=res.hom_cov, het_cov=res.het_cov,
                    corrected_bases=[int(x) for x in res.corrected_bases[:n_round]], n_unfinished=int(res.n_unfinished), device_ms=res.device_ms,
                    ms=dict(ft=res.ms_ft, pt=list(res.ms_pt[:n_round + 1]), ec=list(res.ms_ec[:n_round]), final=res.ms_final, exchange=res.ms_exchange, total=res.ms_total))

    # ---- window alignment
    def ed_semi_64(self, pat, pat_off, txt, txt_off, thre, abs_diag):
        n = len(thre)
        pat = np.ascontiguousarray(pat, dtype=np.uint8); txt = np.ascontiguousarray(txt, dtype=np.uint8)
        po = np.ascontiguousarray(pat_off, dtype=np.uint64); to = np.ascontiguousarray(txt_off, dtype=np.uint64)
        th = np.ascontiguousarray(thre, dtype=np.int32); ab = np.ascontiguousarray(abs_diag, dtype=np.int32)
        err = np.zeros(n, np.int32); pe = np.zeros(n, np.int32)
        self._ck(_lib().hb_ed_semi_64(self.h, C.c_uint64(n), _p(pat), _p(po), _p(txt), _p(to), _p(th), _p(ab), _p(err), _p(pe)))
        return err, pe

    # ---- instrumentation
    def profile(self):
        names = (C.c_char_p * 64)(); launches = (C.c_uint64 * 64)(); ms = (C.c_double * 64)()
        n = _lib().hb_profile(self.h, names, launches, ms, 64)
        return {names[i].decode(): (int(launches[i]), float(ms[i])) for i in range(n)}

    def stage_profile(self):
        """(kernels {name: (launches, ms)}, counters) summed over every pass of the last stage_run"""
        names = (C.c_char_p * 96)(); launches = (C.c_uint64 * 96)(); ms = (C.c_double * 96)(); c = (C.c_uint64 * 12)()
        n = _lib().hb_stage_profile(self.h, names, launches, ms, 96, c, 12)
        keys = ("reads", "bases", "minimizers", "anchors", "groups", "chain_slots", "groups_unordered", "groups_sequential", "windows", "ec_overlaps", "ec_deferred", "ec_segments")
        return {names[i].decode(): (int(launches[i]), float(ms[i])) for i in range(n)}, dict(zip(keys, [int(x) for x in c]))

    def profile_reset(self):
        _lib().hb_profile_reset(self.h)

    def last_pass_ms(self) -> float:
        ms = C.c_double(); _lib().hb_last_pass_ms(self.h, C.byref(ms)); return ms.value

    def counters(self):
        c = (C.c_uint64 * 12)(); _lib().hb_counters(self.h, c, 12)
        return dict(zip(("reads", "bases", "minimizers", "anchors", "groups", "chain_slots", "groups_unordered", "groups_sequential", "windows", "ec_overlaps", "ec_deferred", "ec_segments"),
                        [int(x) for x in c[:12]]))
