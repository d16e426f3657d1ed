"""Reader of refdump's alignment-stage dump (<pfx>.aln.bin, oracle/refdump.cpp step 6): per read and per
overlap the state after each step of gen_hc_r_alin (Correct.cpp:25617): A = align_hc_ed_post_extz +
gen_extend_err_exz, B = gen_hc_fast_cigar, C = reassign_gaps."""
import struct

import numpy as np

WL = np.dtype([("x_start", "<i4"), ("x_end", "<i4"), ("y_start", "<i4"), ("y_end", "<i4"),
               ("extra_begin", "<i2"), ("extra_end", "<i2"), ("error", "<i2"), ("error_threshold", "<i2"),
               ("cidx", "<u4"), ("clen", "<u4")])  # window_list, Hash_Table.h:54-62


def _wl(buf, o):
    n, cn = struct.unpack_from("<II", buf, o); o += 8
    w = np.frombuffer(buf, dtype=WL, count=n, offset=o); o += 32 * n
    c = np.frombuffer(buf, dtype="<u2", count=cn, offset=o); o += 2 * cn
    return (w, c), o


def read_aln(path):
    """-> list over reads of list over overlaps of dict(st, align_length, rr, re, A=(w,c)[, reB, B, C])"""
    buf = open(path, "rb").read(); o = 0; out = []
    while o < len(buf):
        nc, = struct.unpack_from("<I", buf, o); o += 4
        rl = []
        for _ in range(nc):
            st, al, rr, re = struct.unpack_from("<iIdq", buf, o); o += 24
            d = dict(st=st, align_length=al, rr=rr, re=re)
            d["A"], o = _wl(buf, o)
            if st == 2:
                d["reB"], = struct.unpack_from("<q", buf, o); o += 8
                d["B"], o = _wl(buf, o)
                d["C"], o = _wl(buf, o)
                d["nheC"], = struct.unpack_from("<q", buf, o); o += 8
                d["xyC"] = struct.unpack_from("<4I", buf, o); o += 16
            rl.append(d)
        out.append(rl)
    return out


def wl_canon(w, c, mask_extra=False):
    """window list with every cigar resolved (bytes): what must be identical, independent of pool order.
    mask_extra: push_alnw (Correct.cpp:15988) leaves extra_begin / extra_end of an aligned window undefined (stale heap
    content in the reference), so steps B / C compare them only for windows without a cigar."""
    parts = []
    for r in w:
        eb, ee = (0, 0) if (mask_extra and r["clen"] > 0) else (r["extra_begin"], r["extra_end"])
        h = np.array([r["x_start"], r["x_end"], r["y_start"], r["y_end"], eb, ee, r["error"],
                      r["error_threshold"], r["clen"]], dtype="<i4")
        parts.append(h.tobytes()); parts.append(np.ascontiguousarray(c[int(r["cidx"]):int(r["cidx"]) + int(r["clen"])]).tobytes())
    return b"".join(parts)


def _dg(parts):
    import hashlib
    h = hashlib.blake2b(digest_size=8)
    for b in parts:
        h.update(b)
    return int.from_bytes(h.digest(), "little")


def digest_A(ovl):
    """per-read digest of step A: ovl = iterable of (st, align_length, rr, re, w, c)"""
    return _dg(struct.pack("<iIdq", int(st), int(al), float(rr), int(re)) + wl_canon(w, c) for st, al, rr, re, w, c in ovl)


def digest_B(ovl):
    """per-read digest of step B (accepted overlaps only): ovl = iterable of (re, w, c)"""
    return _dg(struct.pack("<q", int(re)) + wl_canon(w, c, True) for re, w, c in ovl)


def digest_C(ovl):
    """per-read digest of step C: ovl = iterable of (non_homopolymer_errors, (x_pos_s, x_pos_e, y_pos_s, y_pos_e), w, c)"""
    return _dg(struct.pack("<q4I", int(nhe), *[int(v) for v in xy]) + wl_canon(w, c, True) for nhe, xy, w, c in ovl)


PH = np.dtype([(f, "<u4") for f in ("y_id", "rev", "x_pos_s", "x_pos_e", "y_pos_s", "y_pos_e", "nh_err", "is_match", "strong")])


RPAF = np.dtype([("qns", "<u8"), ("qe", "<u4"), ("tn", "<u4"), ("ts", "<u4"), ("te", "<u4"), ("rev", "<u4"), ("bl", "<u4"), ("ml", "<u4"), ("no_l_indel", "<u4")])


def read_phase(path):
    """refdump's <pfx>.phase.bin: per read (accepted overlaps after rphase_hc, the list after dedup_chains, the round's reverse_paf records)"""
    buf = open(path, "rb").read(); o = 0; out = []
    while o < len(buf):
        pair = []
        for _ in range(2):
            n, = struct.unpack_from("<I", buf, o); o += 4
            pair.append(np.frombuffer(buf, dtype=PH, count=n, offset=o)); o += 36 * n
        n, = struct.unpack_from("<I", buf, o); o += 4
        pair.append(np.frombuffer(buf, dtype=RPAF, count=n, offset=o)); o += 40 * n
        out.append(tuple(pair))
    return out


def read_ea(path):
    """refdump's <pfx>.ea.bin: per read the accepted overlaps after gen_hc_r_alin_ea: (record of 8 u32, (w, c))"""
    buf = open(path, "rb").read(); o = 0; out = []
    while o < len(buf):
        n, = struct.unpack_from("<I", buf, o); o += 4
        rl = []
        for _ in range(n):
            r8 = struct.unpack_from("<8I", buf, o); o += 32
            wc, o = _wl(buf, o)
            rl.append((r8, wc))
        out.append(rl)
    return out


def digest_ea(ovl):
    """ovl = iterable of ((y_id, rev, x_pos_s, x_pos_e, y_pos_s, y_pos_e, nh_err, is_match), w, c)"""
    return _dg(struct.pack("<8I", *[int(v) & 0xffffffff for v in r8]) + wl_canon(w, c, True) for r8, w, c in ovl)
