"""Host-compiled bodies of the product's per-thread kernels (tests/hostemu)
against the golden vectors of the reference — the GPU-less half of the parity
check of hifiasm_b200/csrc/*.cuh."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu"))
import emu  # noqa: E402
import ha_oracle as ho  # noqa: E402
from goldenlib import Golden, dg  # noqa: E402
import roundlib  # noqa: E402


SETS = os.environ["HB_GOLDEN_NAMES"].split(",") if os.environ.get("HB_GOLDEN_NAMES") else ["g1", "g2", "g3", "g4"]   # the override: tools/fuzz_vs_reference.py


@pytest.fixture(scope="module", params=SETS)
def ctx(request):
    g = Golden(request.param)
    raw = ho.Store(g.raw.length, g.raw.byte_off, g.raw.packed, g.raw.n_off, g.raw.n_pos)
    opt = ho.default_opt()
    ft, hom = ho.ft_gen(raw, opt)
    ho.lib().hao_opt_update_cov(C.byref(opt), hom)
    eft, _, _ = emu.ft_from_oracle(ft)
    return g, opt, ft, eft


@pytest.mark.parametrize("mode", ["raw", "final"])
def test_sketch(ctx, mode):
    g, opt, ft, eft = ctx
    rs = g.raw if mode == "raw" else g.pre
    er = emu.Reads(rs)
    p = g.params(mode)
    for i in range(er.n):
        mz = emu.sketch(er, eft, p, i)
        assert dg(mz.tobytes()) == int(g.digest(mode, "mz")[i]), "read %d" % i
        mz2 = emu.sketch(er, eft, p, i, two_stage=True)  # the production formulation: event stream + per-event window minimum
        assert dg(mz2.tobytes()) == int(g.digest(mode, "mz")[i]), "two-stage sketch, read %d" % i


def _chain_digest(ch, fc):
    import hashlib
    from goldenlib import CH
    h = hashlib.blake2b(digest_size=8)
    for c in ch:
        r = np.zeros(1, dtype=CH)
        for f in ("x_pos_s", "x_pos_e", "y_id", "y_pos_s", "y_pos_e", "y_pos_strand", "shared_seed", "first_hit"):
            r[f] = c[f]
        r["n_fc"] = c["fc_n"]
        h.update(r.tobytes())
        h.update(np.ascontiguousarray(fc[int(c["fc_off"]):int(c["fc_off"]) + int(c["fc_n"])]).tobytes())
    return int.from_bytes(h.digest(), "little")


@pytest.mark.parametrize("mode", ["raw", "final"])
def test_anchors_chains(ctx, mode):
    g, opt, ft, eft = ctx
    rs = g.raw if mode == "raw" else g.pre
    st = ho.Store(rs.length, rs.byte_off, rs.packed, rs.n_off, rs.n_pos)
    pt, hom, het = ho.pt_gen(st, ft, opt)
    ept = emu.pt_from_oracle(pt)
    er = emu.Reads(rs)
    p = g.params(mode)
    bw = float(p["bw_thres"])
    for i in range(er.n):
        mz = emu.sketch(er, eft, p, i)
        an = emu.anchors(er, ept, mz, int(p["high_occ"]), int(p["low_occ"]))
        ch, chits, fc, srt_hits = emu.chains(er, i, an, bw, int(p["k"]), int(p["max_n_chain"]))
        # the chain kernel orders every group: the whole list is now minimizers_qgen0's cl->list
        assert dg(srt_hits.tobytes()) == int(g.digest(mode, "anchors")[i]), "anchors read %d" % i
        assert _chain_digest(ch, fc) == int(g.digest(mode, "chains")[i]), "chains read %d" % i
        assert dg(chits.tobytes()) == int(g.digest(mode, "chain_hits")[i]), "chain hits read %d" % i


# "warp": the forms the product runs (warp kernels as a warp of one lane, the band over virtual lanes); "thread": the one-thread forms kept in
# tests/hostemu/seq_ref.h and the thread aligner for every tier — both must reproduce the reference's fixtures
@pytest.mark.parametrize("forms", ["warp", "thread"])
def test_ec_align_step_A(ctx, forms, monkeypatch):
    if forms == "thread":
        monkeypatch.setenv("HB_EMU_CNS_THREAD", "1"); monkeypatch.setenv("HB_EMU_ALN_THREAD", "1")
    """body of k_ec_overlap (hb_ecaln.cuh) on the raw reads: gap filling, traced windows, extension estimate"""
    import alnlib
    g, opt, ft, eft = ctx
    rs = g.raw
    st = ho.Store(rs.length, rs.byte_off, rs.packed, rs.n_off, rs.n_pos)
    pt, hom, het = ho.pt_gen(st, ft, opt)
    er = emu.Reads(rs)
    p = g.params("raw")
    rd = roundlib.Rounds(g.name); rd_scc, rd_scc_off = rd.scc(0); h_src, h_soff, h_fc, h_ab = rd.hap(0, "src")
    n_full = n_cns = tot_nec = n_rechain = n_reported = 0
    for i in range(er.n):
        mz = ho.sketch(st.decode(i), int(p["w"]), int(p["k"]), 0, 1, ft, int(p["mz_sample_dist"]), int(p["mz_rewin"]))
        an = ho.anchors(st, pt, mz, int(p["high_occ"]), int(p["low_occ"]))
        ch, hits, fc = ho.lchain(st, i, an, float(p["bw_thres"]), int(p["k"]), int(p["max_n_chain"]))
        win = ho.windows(st, i, ch, fc)
        A, W, Cg = emu.ec_align_A(er, i, emu.to_chain(ch), fc, win)
        da = alnlib.digest_A((a["st"], a["align_length"], a["rr"], a["re"], W[int(a["w_off"]):int(a["w_off"]) + int(a["w_n"])], Cg) for a in A)
        assert da == int(g.digest("raw", "alnA")[i]), "read %d" % i
        # step B: body of k_ec_cigar, with a small trace scratch every third read so that the deferral path runs too
        small = (i % 3 == 0)
        # rechain = 1: overlaps that keep an unaligned window of >= 512 bp on both reads go through the re-seeding rescue (body of k_ecb_rechain)
        rc, B, WB, CB = emu.ec_align_B(er, i, emu.to_chain(ch), fc, hits, A, W, path_words=(8192 if small else 1 << 22), poolA=Cg, rechain=1)
        assert rc in ((0, 1) if small else (0,)), rc
        if rc & 1:  # what the second launch does: the deferred overlaps again with the large scratch
            rc, B, WB, CB = emu.ec_align_B(er, i, emu.to_chain(ch), fc, hits, A, W, poolA=Cg, rechain=1)
            assert rc == 0
        acc = B[B["st"] == 2]
        assert acc.size == int(g.count("raw", "aln_ok")[i])
        if g.name == "g4" or g.name.startswith("fz"):  # how many overlaps asked for the rescue; with a hit buffer that is too small the overlap is reported, never silently wrong
            _, B0, _, _ = emu.ec_align_B(er, i, emu.to_chain(ch), fc, hits, A, W)
            n_rechain += int(B0[B0["st"] == 2]["need_rechain"].sum())
            if B0[B0["st"] == 2]["need_rechain"].any():
                _, B1, _, _ = emu.ec_align_B(er, i, emu.to_chain(ch), fc, hits, A, W, poolA=Cg, rechain=(1 << 8) | 1)
                n_reported += int(B1[B1["st"] == 2]["need_rechain"].sum())
        assert not acc["need_rechain"].any()
        if True:
            db = alnlib.digest_B((b["re"], WB[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CB) for b in acc)
            assert db == int(g.digest("raw", "alnB")[i]), "step B, read %d" % i
            # steps B + C in one go (reassign_gaps applied as each window closes)
            rc, Bc, WCc, CCc = emu.ec_align_B(er, i, emu.to_chain(ch), fc, hits, A, W, gaps=1, poolA=Cg, rechain=1)
            assert rc == 0 and (Bc["re"] == B["re"]).all()
            dc = alnlib.digest_C((b["nh_err"], (b["x_pos_s"], b["x_pos_e"], b["y_pos_s"], b["y_pos_e"]), WCc[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CCc) for b in Bc[Bc["st"] == 2])
            assert dc == int(g.digest("raw", "alnC")[i]), "step C, read %d" % i
            # the same through the pieces the GPU runs as three kernels (prep / independent segments with tiered scratch / merge);
            # every other read with merge buffers too small for its longest cigars, so the deferral is reported
            cw = 64 if i % 2 else 1 << 16
            rc, Bp, WP, CP, tiers = emu.ec_align_B_par(er, i, emu.to_chain(ch), fc, hits, A, W, gaps=1, cig_words=cw, poolA=Cg, rechain=1)
            if rc & 1:
                assert cw == 64 and (Bp["st"] == -1).any()
                rc, Bp, WP, CP, tiers = emu.ec_align_B_par(er, i, emu.to_chain(ch), fc, hits, A, W, gaps=1, poolA=Cg, rechain=1)
            assert rc == 0 and (Bp["re"] == B["re"]).all()
            dp = alnlib.digest_C((b["nh_err"], (b["x_pos_s"], b["x_pos_e"], b["y_pos_s"], b["y_pos_e"]), WP[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CP) for b in Bp[Bp["st"] == 2])
            assert dp == int(g.digest("raw", "alnC")[i]), "steps B + C, segment-parallel pipeline, read %d" % i
            # phasing (row a13): bodies of k_ph_count / k_ph_decide
            im, sg = emu.ec_phase(er, i, emu.to_chain(ch), A, Bc, WCc, CCc)
            acc_b = Bc[Bc["st"] == 2]; acc_c = ch[Bc["st"] == 2]
            pa = np.zeros(acc_b.size, alnlib.PH)
            pa["y_id"] = acc_c["y_id"]; pa["rev"] = acc_c["y_pos_strand"]; pa["nh_err"] = acc_b["nh_err"].astype(np.uint32); pa["is_match"] = im; pa["strong"] = sg.astype(np.int32).astype(np.uint32)
            for f in ("x_pos_s", "x_pos_e", "y_pos_s", "y_pos_e"):
                pa[f] = acc_b[f]
            assert dg(pa.tobytes()) == int(g.digest("raw", "phase")[i]), "rphase_hc, read %d" % i
            # the round's reverse_paf[i]: dedup_chains + push_ne_ovlp(flag 2) (body of k_ec_rpaf)
            ph = np.zeros(pa.size, emu.PHASE); ph["st"] = 2
            for f in alnlib.PH.names:
                ph[f] = pa[f].astype(np.int64).astype(ph[f].dtype)
            rp = emu.ec_reverse(er, i, ph); ra = np.zeros(rp.size, alnlib.RPAF)
            for f in alnlib.RPAF.names:
                ra[f] = rp[f]
            assert ra.size == int(g.count("raw", "rpaf")[i]) and dg(ra.tobytes()) == int(g.digest("raw", "rpaf")[i]), "reverse_paf, read %d" % i
            # the round's paf[i] (row a15): push_ne_ovlp(flag 1) with the longest exact interval mapped through the read's edit script (here the
            # reference's own script of round 0), the large-indel flag and check_well_cal's two read flags — against the state after cal_ec_multiple
            sc = rd_scc[int(rd_scc_off[i]):int(rd_scc_off[i + 1])]
            # the read's edit script (row a14): window consensus, voted path; reads that need the graph consensus are reported, not corrected
            st_c, my_sc, nec = emu.ec_cns(er, i, ph, acc_b, WCc, CCc)
            if st_c == 1:  # the second launch: the same read with the arena of the graph consensus (cns_gen_full)
                n_full += 1
                st_c, my_sc, nec = emu.ec_cns(er, i, ph, acc_b, WCc, CCc, g_nodes=8192, g_arcs=1 << 16)
                assert st_c == 0, "graph consensus arena, read %d" % i
                st_small, _, _ = emu.ec_cns(er, i, ph, acc_b, WCc, CCc, g_nodes=1, g_arcs=1)
                assert st_small == 3  # an arena that is too small is reported, never silently wrong
            assert st_c == 0
            assert my_sc.tobytes() == sc.tobytes(), "edit script, read %d" % i
            n_cns += 1; tot_nec += nec
            sp, f_ec, f_ab = emu.ec_source(er, i, ph, acc_b, WCc, CCc, sc)
            want = roundlib.canon_list(h_src[int(h_soff[i]):int(h_soff[i + 1])], 0)
            assert sp.size == want.size, "paf, read %d" % i
            assert roundlib.canon_list(sp, 0).tobytes() == want.tobytes(), "paf, read %d" % i
            assert (f_ec, f_ab) == (int(h_fc[i]), int(h_ab[i])), "is_fully_corrected / is_abnormal, read %d" % i
    # almost every read is corrected by the voted path; the graph consensus ran cns_gen_full "full_calls" times in the reference's round 0
    assert n_cns > 0 and (0 < n_full or g.name.startswith("fz")) and n_full <= int(rd.params(0)["full_calls"]), (n_cns, n_full)
    if g.name == "g4":
        assert n_rechain >= 20 and 0 < n_reported <= n_rechain, (n_rechain, n_reported)
        print("re-seeding rescue (rechain_aln_hc): %d overlaps; %d of them reported when the hit buffer holds one hit" % (n_rechain, n_reported))
    print("window consensus: %d reads (%d corrected bases), %d of them through the graph consensus" % (n_cns, tot_nec, n_full))


def test_final_pass(ctx):
    from hifiasm_b200 import binio
    g, opt, ft, eft = ctx
    rs = g.pre
    st = ho.Store(rs.length, rs.byte_off, rs.packed, rs.n_off, rs.n_pos)
    pt, hom, het = ho.pt_gen(st, ft, opt)
    ept = emu.pt_from_oracle(pt)
    er = emu.Reads(rs)
    p = g.params("final")
    p0, o0, _, _ = g.pre_src
    p1, o1, _, _ = g.pre_rev
    f0, fo0, _, _ = g.fin_src
    f1, fo1, _, _ = g.fin_rev
    m0, m1 = binio.disk_to_mem(p0), binio.disk_to_mem(p1)
    for i in range(er.n):
        a, b = emu.final_read(er, eft, ept, p, hom, int(p["max_n_chain"]), i, m0[int(o0[i]):int(o0[i + 1])], m1[int(o1[i]):int(o1[i + 1])])
        ra, rb = f0[int(fo0[i]):int(fo0[i + 1])], f1[int(fo1[i]):int(fo1[i + 1])]
        assert a.size == ra.size and b.size == rb.size, "read %d" % i
        for f in binio.MA_DISK.names:
            assert (a[f] == ra[f]).all() and (b[f] == rb[f]).all(), "read %d field %s" % (i, f)
