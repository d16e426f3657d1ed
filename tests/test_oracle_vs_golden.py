"""Pins the CPU oracle (oracle/ha_oracle.c) to the reference: every stage of
the hot path is compared with digests / dumps produced by the unmodified
reference (tests/golden/*.npz <- oracle/_ref/refdump)."""
import ctypes as C
import os

import numpy as np
import pytest

import ha_oracle as ho
from goldenlib import Golden, dg, chain_digest
import alnlib
from hifiasm_b200 import binio


def _store(rs):
    return ho.Store(rs.length, rs.byte_off, rs.packed, rs.n_off, rs.n_pos)


@pytest.fixture(scope="module", params=os.environ["HB_GOLDEN_NAMES"].split(",") if os.environ.get("HB_GOLDEN_NAMES") else ["g1", "g2", "g3", "g4"])
def ctx(request):
    g = Golden(request.param)
    raw = _store(g.raw)
    opt = ho.default_opt()
    ft, hom = ho.ft_gen(raw, opt)
    ho.lib().hao_opt_update_cov(C.byref(opt), hom)
    return g, raw, opt, ft


n_rechain = [0]


def _ea_records(ch, fl, A, Cc, WC, CC):
    """accepted overlaps after gen_hc_r_alin_ea: the exact shortcut's single window, or steps A-C's result"""
    out = []
    for j in range(ch.size):
        c = ch[j]
        if fl[j]:
            L = int(c["x_pos_e"]) + 1 - int(c["x_pos_s"]); cg = []
            while L >= 0x3fff:
                cg.append(0x3fff); L -= 0x3fff
            if L:
                cg.append(L)
            w = np.zeros(1, alnlib.WL)
            w[0]["x_start"] = c["x_pos_s"]; w[0]["x_end"] = c["x_pos_e"]; w[0]["y_start"] = c["y_pos_s"]; w[0]["y_end"] = c["y_pos_e"]; w[0]["clen"] = len(cg)
            out.append(((c["y_id"], c["y_pos_strand"], c["x_pos_s"], c["x_pos_e"], c["y_pos_s"], c["y_pos_e"], 0, 1), w, np.array(cg, np.uint16)))
        elif A[j]["st"] == 2:
            cc = Cc[j]
            out.append(((c["y_id"], c["y_pos_strand"], cc["x_pos_s"], cc["x_pos_e"], cc["y_pos_s"], cc["y_pos_e"], cc["nh_err"], 1),
                        WC[int(cc["w_off"]):int(cc["w_off"] + cc["w_n"])], CC[int(cc["c_off"]):int(cc["c_off"] + cc["c_n"])]))
    return out


def _stages(g, mode, st, ft, opt, bw, prev=None):
    p = g.params(mode)
    pt, hom, het = ho.pt_gen(st, ft, opt)
    assert (hom, het) == (int(p["hom_cov"]), int(p["het_cov"]))
    assert opt.max_n_chain == int(p["max_n_chain"])
    for i in range(st.n):
        mz = ho.sketch(st.decode(i), int(p["w"]), int(p["k"]), 0, 1, ft, int(p["mz_sample_dist"]), int(p["mz_rewin"]))
        assert dg(mz.tobytes()) == int(g.digest(mode, "mz")[i]), "sketch read %d" % i
        import hashlib
        h = hashlib.blake2b(digest_size=8)
        for x in mz["x"]:
            pos = ho.pt_get(pt, x)
            h.update(np.array([x], dtype="<u8").tobytes()); h.update(np.array([pos.size], dtype="<u4").tobytes()); h.update(pos.tobytes())
        assert int.from_bytes(h.digest(), "little") == int(g.digest(mode, "idx")[i]), "index read %d" % i
        an = ho.anchors(st, pt, mz, int(p["high_occ"]), int(p["low_occ"]))
        assert dg(an.tobytes()) == int(g.digest(mode, "anchors")[i]), "anchors read %d" % i
        ch, hits, fc = ho.lchain(st, i, an, bw, int(p["k"]), int(p["max_n_chain"]))
        assert chain_digest(ch, fc) == int(g.digest(mode, "chains")[i]), "chains read %d" % i
        assert dg(hits.tobytes()) == int(g.digest(mode, "chain_hits")[i]), "chain hits read %d" % i
        win = ho.windows(st, i, ch, fc)
        assert win.size == int(g.count(mode, "windows")[i]) and dg(win.tobytes()) == int(g.digest(mode, "windows")[i]), "window pass read %d" % i
        # step A of the alignment stage (rows a8 + a9): gap filling, traced windows, extension error estimate
        A, W, Cg = ho.ec_align_A(st, i, ch, fc)
        da = alnlib.digest_A((a["st"], a["align_length"], a["rr"], a["re"], W[int(a["w_off"]):int(a["w_off"] + a["w_n"])],
                              Cg[int(a["c_off"]):int(a["c_off"] + a["c_n"])]) for a in A)
        assert da == int(g.digest(mode, "alnA")[i]), "EC alignment step A, read %d" % i
        assert int((A["st"] == 2).sum()) == int(g.count(mode, "aln_ok")[i])
        # step B (row a10): base-level CIGAR of the accepted overlaps
        B, WB, CB = ho.ec_align_B(st, i, ch, fc, hits, A, W)
        acc = B[B["st"] == 2]
        if not acc["need_rechain"].any():  # an overlap that needs the re-chaining rescue (not restated yet) is not comparable
            db = alnlib.digest_B((b["re"], WB[int(b["w_off"]):int(b["w_off"] + b["w_n"])], CB[int(b["c_off"]):int(b["c_off"] + b["c_n"])]) for b in acc)
            assert db == int(g.digest(mode, "alnB")[i]), "EC alignment step B, read %d" % i
            # step C (row a11): reassign_gaps
            Cc, WC, CC = ho.ec_align_C(st, i, ch, A, B, WB, CB)
            dc = alnlib.digest_C((c["nh_err"], (c["x_pos_s"], c["x_pos_e"], c["y_pos_s"], c["y_pos_e"]), WC[int(c["w_off"]):int(c["w_off"] + c["w_n"])],
                                  CC[int(c["c_off"]):int(c["c_off"] + c["c_n"])]) for c, a in zip(Cc, A) if a["st"] == 2)
            assert dc == int(g.digest(mode, "alnC")[i]), "EC alignment step C, read %d" % i
            # phasing (row a13) + dedup_chains: per-overlap is_match / strong, then the best chain per target
            P, D = ho.ec_phase(st, i, ch, A, Cc, WC, CC)
            pa = np.zeros(P.size, alnlib.PH); da = np.zeros(D.size, alnlib.PH)
            for f in alnlib.PH.names:
                pa[f] = P[f].astype(np.int64).astype(np.uint32); da[f] = D[f].astype(np.int64).astype(np.uint32)
            assert dg(pa.tobytes()) == int(g.digest(mode, "phase")[i]), "rphase_hc, read %d" % i
            assert dg(da.tobytes()) == int(g.digest(mode, "dedup")[i]), "dedup_chains, read %d" % i
            # row a12: the exact shortcut from the previous round's overlap list (empty for the raw reads: round 0)
            pm, po = prev if prev is not None else (np.zeros(0, ho.MA), np.zeros(st.n + 1, np.uint64))
            fl = ho.ec_ea_flags(st, i, ch, pm[int(po[i]):int(po[i + 1])])
            assert alnlib.digest_ea(_ea_records(ch, fl, A, Cc, WC, CC)) == int(g.digest(mode, "ea")[i]), "gen_hc_r_alin_ea, read %d" % i
        else:
            n_rechain[0] += 1
    return pt, hom, het


def test_stages_raw(ctx):
    g, raw, opt, ft = ctx
    _stages(g, "raw", raw, ft, opt, 0.02)


def test_stages_and_final_pass(ctx):
    g, raw, opt, ft = ctx
    st = _store(g.pre)
    p0, o0, _, _ = g.pre_src
    pt, hom, het = _stages(g, "final", st, ft, opt, 0.001, prev=(binio.disk_to_mem(p0), o0))
    opt.hom_cov, opt.het_cov = hom, het
    p0, o0, _, _ = g.pre_src
    p1, o1, _, _ = g.pre_rev
    f0, fo0, _, _ = g.fin_src
    f1, fo1, _, _ = g.fin_rev
    m0, m1 = binio.disk_to_mem(p0), binio.disk_to_mem(p1)
    for i in range(st.n):
        a, b = ho.final_read(st, pt, ft, opt, i, m0[int(o0[i]):int(o0[i + 1])], m1[int(o1[i]):int(o1[i + 1])])
        ra, rb = f0[int(fo0[i]):int(fo0[i + 1])], f1[int(fo1[i]):int(fo1[i + 1])]
        assert a.size == ra.size and b.size == rb.size, "read %d" % i
        for f in binio.MA_DISK.names:
            assert (a[f] == ra[f]).all() and (b[f] == rb[f]).all(), "read %d field %s" % (i, f)


def test_myers_window_vs_reference():
    z = np.load(__import__("os").path.join(__import__("goldenlib").GOLDEN, "ed_semi.npz"))
    hdr, pat, txt, res = z["hdr"], z["pat"], z["txt"], z["res"]
    po = to = 0
    L = ho.lib()
    for i, (pn, tn, thre, ab) in enumerate(hdr):
        p = pat[po:po + pn].tobytes(); t = txt[to:to + tn].tobytes(); po += pn; to += tn
        pe = C.c_int32()
        err = L.hao_ed_semi_64_absent_diag(p, int(pn), t, int(tn), int(thre), int(ab), C.byref(pe))
        assert (err, pe.value) == (int(res[i, 0]), int(res[i, 1])), "case %d" % i


def test_bin_roundtrip(tmp_path):
    g = Golden("g1")
    rec, off, fc, ab = g.fin_src
    p = str(tmp_path / "x.bin")
    binio.write_ovlp_bin(p, rec, off, fc, ab)
    assert open(p, "rb").read() == g.z["fin_ovlp_source"].tobytes()
