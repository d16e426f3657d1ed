"""GPU parity suite: every stage of the CUDA path (through the C-ABI of
include/hifiasm_b200.h) against the golden vectors of the unmodified reference
and against the CPU oracle on the same inputs.  Bit-exact (integer/index work)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ha_oracle as ho  # noqa: E402
from goldenlib import Golden, dg, CH, GOLDEN  # noqa: E402
import alnlib  # noqa: E402
from hifiasm_b200 import binio  # noqa: E402


@pytest.fixture(scope="module")
def hb():
    import hifiasm_b200
    return hifiasm_b200


@pytest.fixture(scope="module", params=["g1", "g2", "g3"])
def ctx(request, hb):
    g = Golden(request.param)
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen()
    eng.update_cov(hom)
    return g, eng, hom


def _chain_digest(ch, fc):
    h = hashlib.blake2b(digest_size=8)
    for c in ch:
        r = np.zeros(1, dtype=CH)
        for f in ("x_pos_s", "x_pos_e", "y_id", "y_pos_s", "y_pos_e", "y_pos_strand", "shared_seed", "first_hit"):
            r[f] = c[f]
        r["n_fc"] = c["fc_n"]
        h.update(r.tobytes())
        h.update(np.ascontiguousarray(fc[int(c["fc_off"]):int(c["fc_off"]) + int(c["fc_n"])]).tobytes())
    return int.from_bytes(h.digest(), "little")


def _check_stages(g, eng, mode, rs):
    p = g.params(mode)
    eng.upload_store(rs)
    hom, het = eng.pt_gen()
    assert (hom, het) == (int(p["hom_cov"]), int(p["het_cov"]))
    assert eng.get_opt().max_n_chain == int(p["max_n_chain"])
    eng.set_opt(hom_cov=hom, het_cov=het)
    n = rs.n
    off, mz = eng.sketch(0, n)
    mzq = mz.copy(); mzq["info"] &= ~np.uint64(0xfffffff)  # the query path sketches with rid = 0 (anchor.cpp:1003)
    for i in range(n):
        assert dg(mzq[int(off[i]):int(off[i + 1])].tobytes()) == int(g.digest(mode, "mz")[i]), "sketch read %d" % i
    # index: every query minimizer of every read, like refdump's .idx.bin
    cnt, pos = eng.pt_get(mz["x"])
    po = np.zeros(cnt.size + 1, np.uint64); np.cumsum(cnt, out=po[1:])
    for i in range(n):
        h = hashlib.blake2b(digest_size=8)
        for j in range(int(off[i]), int(off[i + 1])):
            h.update(np.array([mz["x"][j]], dtype="<u8").tobytes()); h.update(np.array([cnt[j]], dtype="<u4").tobytes())
            h.update(pos[int(po[j]):int(po[j + 1])].tobytes())
        assert int.from_bytes(h.digest(), "little") == int(g.digest(mode, "idx")[i]), "index read %d" % i
    aoff, an = eng.anchors(0, n)
    for i in range(n):
        assert dg(an[int(aoff[i]):int(aoff[i + 1])].tobytes()) == int(g.digest(mode, "anchors")[i]), "anchors read %d" % i
    coff, ch, hoff, hits, foff, fc = eng.chains(0, n, float(p["bw_thres"]))
    for i in range(n):
        assert _chain_digest(ch[int(coff[i]):int(coff[i + 1])], fc[int(foff[i]):int(foff[i + 1])]) == int(g.digest(mode, "chains")[i]), "chains read %d" % i
        assert dg(hits[int(hoff[i]):int(hoff[i + 1])].tobytes()) == int(g.digest(mode, "chain_hits")[i]), "chain anchors read %d" % i
    # window pass of the EC rounds (row a8): every window of every chain
    woff, win = eng.windows(0, n, float(p["bw_thres"]), 0.04, 775)
    for i in range(n):
        w = win[int(woff[i]):int(woff[i + 1])]
        assert w.size == int(g.count(mode, "windows")[i]) and dg(w.tobytes()) == int(g.digest(mode, "windows")[i]), "window pass read %d" % i
    # step A of the alignment stage (rows a8 + a9): k_windows + k_ec_overlap
    ooff, A, W, Cg = eng.ec_align(0, n, float(p["bw_thres"]), 0.04, 775)
    assert (ooff == coff).all()
    for i in range(n):
        a_i = A[int(ooff[i]):int(ooff[i + 1])]
        da = alnlib.digest_A((a["st"], a["align_length"], a["rr"], a["re"], W[int(a["w_off"]):int(a["w_off"]) + int(a["w_n"])], Cg) for a in a_i)
        assert da == int(g.digest(mode, "alnA")[i]), "EC alignment step A, read %d" % i
        assert int((a_i["st"] == 2).sum()) == int(g.count(mode, "aln_ok")[i])
    # steps A + B (rows a8-a10): base-level CIGARs of the accepted overlaps, k_ec_cigar
    boff, B, WB, CB = eng.ec_cigar(0, n, float(p["bw_thres"]), 0.04, 775)
    assert (boff == coff).all() and (B["st"] == A["st"]).all()
    for i in range(n):
        acc = B[int(boff[i]):int(boff[i + 1])]; acc = acc[acc["st"] == 2]
        assert not acc["need_rechain"].any(), "re-seeding rescue out of scratch, read %d" % i
        db = alnlib.digest_B((b["re"], WB[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CB) for b in acc)
        assert db == int(g.digest(mode, "alnB")[i]), "EC alignment step B, read %d" % i
    # + step C (row a11): reassign_gaps fused into the same kernel
    goff, Gc, WG, CG = eng.ec_cigar(0, n, float(p["bw_thres"]), 0.04, 775, gaps=1)
    assert (Gc["st"] == B["st"]).all() and (Gc["re"] == B["re"]).all()
    for i in range(n):
        acc = Gc[int(goff[i]):int(goff[i + 1])]; acc = acc[acc["st"] == 2]
        assert not acc["need_rechain"].any(), "re-seeding rescue out of scratch, read %d" % i
        dc = alnlib.digest_C((b["nh_err"], (b["x_pos_s"], b["x_pos_e"], b["y_pos_s"], b["y_pos_e"]), WG[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CG) for b in acc)
        assert dc == int(g.digest(mode, "alnC")[i]), "EC alignment step C, read %d" % i
    # phasing (row a13) on top of the step-C state: k_ph_count / k_ph_decide
    poff, P = eng.ec_phase(0, n, float(p["bw_thres"]), 0.04, 775)
    assert (poff == coff).all() and (P["st"] == B["st"]).all()
    for i in range(n):
        acc = P[int(poff[i]):int(poff[i + 1])]; acc = acc[acc["st"] == 2]
        assert not acc["need_rechain"].any(), "re-seeding rescue out of scratch, read %d" % i
        pa = np.zeros(acc.size, alnlib.PH)
        for f in alnlib.PH.names:
            pa[f] = acc[f].astype(np.int64).astype(np.uint32)
        assert dg(pa.tobytes()) == int(g.digest(mode, "phase")[i]), "rphase_hc, read %d" % i
        assert int((acc["is_match"] == 2).sum()) == int(g.count(mode, "phase_hap2")[i])
    # the round's reverse_paf[i] (part of row a15): dedup_chains + push_ne_ovlp(flag 2)
    roff, RP = eng.ec_reverse_paf(0, n, float(p["bw_thres"]), 0.04, 775)
    for i in range(n):
        rp = RP[int(roff[i]):int(roff[i + 1])]; ra = np.zeros(rp.size, alnlib.RPAF)
        for f in alnlib.RPAF.names:
            ra[f] = rp[f]
        assert ra.size == int(g.count(mode, "rpaf")[i]) and dg(ra.tobytes()) == int(g.digest(mode, "rpaf")[i]), "reverse_paf of the round, read %d" % i
    # row a12: the whole alignment stage with the previous round's exact overlaps as a shortcut (gen_hc_r_alin_ea)
    if mode == "final":
        p0, o0, _, _ = g.pre_src
        eng.ec_stage_prev(binio.disk_to_mem(p0), o0)
    else:
        eng.ec_stage_prev(np.zeros(0, binio.MA_MEM), np.zeros(n + 1, np.uint64))
    eoff, E, WE, CE = eng.ec_cigar(0, n, float(p["bw_thres"]), 0.04, 775, gaps=3)
    n_short = 0
    for i in range(n):
        acc = E[int(eoff[i]):int(eoff[i + 1])]; cc = ch[int(coff[i]):int(coff[i + 1])][acc["st"] == 2]; acc = acc[acc["st"] == 2]
        assert not acc["need_rechain"].any(), "re-seeding rescue out of scratch, read %d" % i
        n_short += int(acc["pad"].sum())
        de = alnlib.digest_ea(((c["y_id"], c["y_pos_strand"], b["x_pos_s"], b["x_pos_e"], b["y_pos_s"], b["y_pos_e"], b["nh_err"], 1), WE[int(b["w_off"]):int(b["w_off"]) + int(b["w_n"])], CE) for b, c in zip(acc, cc))
        assert de == int(g.digest(mode, "ea")[i]), "gen_hc_r_alin_ea, read %d" % i
    assert (n_short > 0) == (mode == "final")
    return hom, het


def test_filter_table_vs_oracle(ctx):
    g, eng, hom = ctx
    raw = ho.Store(g.raw.length, g.raw.byte_off, g.raw.packed, g.raw.n_off, g.raw.n_pos)
    opt = ho.default_opt()
    ft, ohom = ho.ft_gen(raw, opt)
    assert hom == ohom
    n = int(ho.lib().hao_ft_size(C.c_void_p(ft)))
    assert eng.ft_size() == n
    key = np.zeros(max(n, 1), np.uint64); val = np.zeros(max(n, 1), np.int32)
    if n:
        ho.lib().hao_ft_dump(C.c_void_p(ft), C.c_void_p(key.ctypes.data), C.c_void_p(val.ctypes.data))
        assert (eng.ft_cnt(key[:n]) == val[:n]).all()
    rng = np.random.default_rng(1)
    assert (eng.ft_cnt(rng.integers(0, 2**63, 1000, dtype=np.uint64)) == 0).all()


def test_filter_table_in_hash_range_chunks(ctx, monkeypatch):
    """exact k-mer counting in 4 hash ranges (what a read set too large for one pass gets): same peak, same table contents"""
    g, eng0, hom = ctx
    import hifiasm_b200
    monkeypatch.setenv("HB_FT_CHUNK_BITS", "2")
    eng = hifiasm_b200.Engine(0)
    try:
        eng.upload_store(g.raw)
        assert eng.ft_gen() == hom and eng.ft_size() == eng0.ft_size()
        raw = ho.Store(g.raw.length, g.raw.byte_off, g.raw.packed, g.raw.n_off, g.raw.n_pos)
        ft, _ = ho.ft_gen(raw, ho.default_opt())
        n = int(ho.lib().hao_ft_size(C.c_void_p(ft)))
        assert eng.ft_size() == n
        if n:
            key = np.zeros(n, np.uint64); val = np.zeros(n, np.int32)
            ho.lib().hao_ft_dump(C.c_void_p(ft), C.c_void_p(key.ctypes.data), C.c_void_p(val.ctypes.data))
            assert (eng.ft_cnt(key) == val).all()
        assert (eng.ft_cnt(np.random.default_rng(2).integers(0, 2**63, 1000, dtype=np.uint64)) == 0).all()
    finally:
        eng.close()


def test_stages_raw(ctx):
    g, eng, hom = ctx
    _check_stages(g, eng, "raw", g.raw)


def test_stages_final_and_cal_ov_r(ctx, tmp_path):
    g, eng, hom = ctx
    _check_stages(g, eng, "final", g.pre)
    p0, o0, fc0, ab0 = g.pre_src
    p1, o1, _, _ = g.pre_rev
    out0, oo0, out1, oo1, stat = eng.cal_ov_r(binio.disk_to_mem(p0), o0, binio.disk_to_mem(p1), o1)
    f0, fo0, ffc, fab = g.fin_src
    f1, fo1, _, _ = g.fin_rev
    assert (oo0 == fo0).all() and (oo1 == fo1).all()
    for f in binio.MA_DISK.names:
        assert (out0[f] == f0[f]).all(), "source field %s" % f
        assert (out1[f] == f1[f]).all(), "reverse field %s" % f
    # byte-identical .bin files through the reference's own format
    a = str(tmp_path / "src.bin"); b = str(tmp_path / "rev.bin")
    binio.write_ovlp_bin(a, out0, oo0, ffc, fab); binio.write_ovlp_bin(b, out1, oo1, g.fin_rev[2], g.fin_rev[3])
    assert open(a, "rb").read() == g.z["fin_ovlp_source"].tobytes()
    assert open(b, "rb").read() == g.z["fin_ovlp_reverse"].tobytes()
    assert int(stat[0]) == f0.size and int(stat[1]) == f1.size
    # a second run must give the same answer (the staged previous overlaps are not consumed)
    n0, n1, stat2 = eng.cal_ov_r_resident()
    assert (n0, n1) == (f0.size, f1.size) and (stat2 == stat).all()
    prof = eng.profile()
    for k in ("k_probe_count", "k_expand", "sort_anchors", "k_chain", "k_post", "k_exact", "k_merge"):   # (no sketch kernels: the pass reads the sketch hb_pt_gen made of these reads)
        assert k in prof and prof[k][0] >= 1, k


def test_small_batches_same_result(ctx, monkeypatch):
    """anchor-budget batching must not change results"""
    g, eng0, hom = ctx
    import hifiasm_b200
    monkeypatch.setenv("HB_ANCHOR_BUDGET", "200000")
    eng = hifiasm_b200.Engine(0)
    eng.upload_store(g.raw); eng.update_cov(eng.ft_gen())
    eng.upload_store(g.pre)
    hom, het = eng.pt_gen(); eng.set_opt(hom_cov=hom, het_cov=het)
    p0, o0, _, _ = g.pre_src; p1, o1, _, _ = g.pre_rev
    out0, oo0, out1, oo1, stat = eng.cal_ov_r(binio.disk_to_mem(p0), o0, binio.disk_to_mem(p1), o1)
    f0, fo0, _, _ = g.fin_src; f1, fo1, _, _ = g.fin_rev
    assert (oo0 == fo0).all() and (oo1 == fo1).all()
    for f in binio.MA_DISK.names:
        assert (out0[f] == f0[f]).all() and (out1[f] == f1[f]).all()
    eng.close()


def test_sharded_query_ranges_equal_full_pass(ctx):
    """the multi-GPU decomposition: disjoint query ranges against the replicated
    index give exactly the per-read results of the full pass"""
    from hifiasm_b200 import dist as hdist
    g, eng, hom = ctx
    eng.upload_store(g.pre)
    h, t = eng.pt_gen(); eng.set_opt(hom_cov=h, het_cov=t)
    p0, o0, _, _ = g.pre_src; p1, o1, _, _ = g.pre_rev
    m0, m1 = binio.disk_to_mem(p0), binio.disk_to_mem(p1)
    full0, fo0, full1, fo1, _ = eng.cal_ov_r(m0, o0, m1, o1)
    n = g.pre.n; parts0, parts1 = [], []
    for rank in range(3):
        a, b = hdist.shard_range(n, rank, 3)
        x0, y0, x1, y1, _ = eng.cal_ov_r(m0, o0, m1, o1, a, b)
        parts0.append((x0, y0)); parts1.append((x1, y1))
    r0, q0 = hdist.merge_shards(parts0); r1, q1 = hdist.merge_shards(parts1)
    assert (q0 == fo0).all() and (q1 == fo1).all()
    assert r0.tobytes() == full0.tobytes() and r1.tobytes() == full1.tobytes()


def test_ec_cigar_small_scratch_defers_and_matches(ctx, monkeypatch):
    """scratch tiers too small for some segments / overlaps: they go through the next tier / the deferred merge, same result"""
    g, eng, hom = ctx
    p = g.params("raw")
    eng.upload_store(g.raw)
    h, t = eng.pt_gen(); eng.set_opt(hom_cov=h, het_cov=t)
    n = g.raw.n
    ref = eng.ec_cigar(0, n, float(p["bw_thres"]), 0.04, 775)
    import subprocess, sys, json
    # HB_ECB_PATH_WORDS is read once per context (hb_create): run the small-scratch pass in a child
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import hifiasm_b200; from goldenlib import Golden; import hashlib;"
            "g = Golden(%r); e = hifiasm_b200.Engine(0); e.upload_store(g.raw); hom = e.ft_gen(); e.update_cov(hom); h, t = e.pt_gen(); e.set_opt(hom_cov=h, het_cov=t);"
            "o, B, W, Cg = e.ec_cigar(0, g.raw.n, %f, 0.04, 775);"
            "d = hashlib.blake2b(digest_size=8); [d.update(B[f].tobytes()) for f in ('st', 're', 'x_pos_s', 'x_pos_e', 'y_pos_s', 'y_pos_e', 'w_n')];"
            "[d.update(np.ascontiguousarray(W[f]).tobytes()) for f in ('x_start', 'x_end', 'y_start', 'y_end', 'error', 'clen')];"
            "print(json.dumps({'dg': d.hexdigest(), 'deferred': e.counters()['ec_deferred']}))") % (
                os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), [k for k in ("g1", "g2", "g3") if Golden(k).raw.n == n][0], float(p["bw_thres"]))
    env = dict(os.environ, HB_ECB_PATH_WORDS="64", HB_ECB_CIG_WORDS="64")  # tier 1: 64 trace words per thread, tier 2: 4096 per warp
    res = json.loads(subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip().splitlines()[-1])
    o, B, W, Cg = ref
    d = hashlib.blake2b(digest_size=8)
    for f in ("st", "re", "x_pos_s", "x_pos_e", "y_pos_s", "y_pos_e", "w_n"):
        d.update(B[f].tobytes())
    for f in ("x_start", "x_end", "y_start", "y_end", "error", "clen"):
        d.update(np.ascontiguousarray(W[f]).tobytes())
    assert res["dg"] == d.hexdigest()
    if n == Golden("g3").raw.n:
        assert res["deferred"] > 0  # the damaged set has segments that need more than 4096 trace words: they reach the last tier


def test_myers_window_vs_reference(hb):
    z = np.load(os.path.join(GOLDEN, "ed_semi.npz"))
    hdr, pat, txt, res = z["hdr"], z["pat"], z["txt"], z["res"]
    po = np.zeros(hdr.shape[0] + 1, np.uint64); to = np.zeros(hdr.shape[0] + 1, np.uint64)
    np.cumsum(hdr[:, 0], out=po[1:]); np.cumsum(hdr[:, 1], out=to[1:])
    eng = hb.Engine(0)
    err, pe = eng.ed_semi_64(pat, po, txt, to, hdr[:, 2], hdr[:, 3])
    assert (err == res[:, 0]).all() and (pe == res[:, 1]).all()
    eng.close()


def test_empty_and_edge_inputs(hb):
    eng = hb.Engine(0)
    # a store of reads too short to carry any minimizer, plus one with only Ns
    reads = [np.zeros(10, np.uint8), np.full(60, 4, np.uint8), np.arange(200, dtype=np.uint8) % 4]
    from hifiasm_b200 import sim
    flat, boff, ln, npos, noff = sim.pack_reads(reads)
    eng.upload_reads(ln, flat, boff, npos, noff)
    hom = eng.ft_gen()
    hom2, het2 = eng.pt_gen()
    off, mz = eng.sketch(0, 3)
    assert off[1] == 0 and off[2] == 0  # no k-mer fits in the first two reads
    e0 = np.zeros(0, binio.MA_MEM); z = np.zeros(4, np.uint64)
    out0, oo0, out1, oo1, stat = eng.cal_ov_r(e0, z, e0, z)
    assert out0.size == 0 and out1.size == 0
    eng.close()


def test_gpu_vs_oracle_noisy_repeats(hb):
    """seeded reads with errors, a repeat family and N bases (raw-read context: the
    chaining DP, secondary chains, the per-type cap and the shadow filter are all
    exercised); every read's anchors / chains / chain anchors / fake cigars equal the
    CPU oracle's on the same inputs"""
    from hifiasm_b200 import sim
    h1, h2 = sim.sim_genome(500000, 21, snp_rate=0.003, repeat_frac=0.25, repeat_len=1800, repeat_div=0.01)
    reads = sim.sim_reads(h1, h2, 24, 7000, 21, sd_len=2500, min_len=300, err=0.004, n_rate=1e-4)
    flat, boff, ln, npos, noff = sim.pack_reads(reads)
    n = len(reads)
    eng = hb.Engine(0)
    eng.upload_reads(ln, flat, boff, npos, noff)
    hom = eng.ft_gen(); eng.update_cov(hom)
    hom2, het2 = eng.pt_gen(); eng.set_opt(hom_cov=hom2, het_cov=het2)
    st = ho.Store(ln, boff, flat, noff, npos)
    opt = ho.default_opt()
    ft, ohom = ho.ft_gen(st, opt)
    ho.lib().hao_opt_update_cov(C.byref(opt), ohom)
    pt, h3, t3 = ho.pt_gen(st, ft, opt)
    assert (hom, hom2, het2) == (ohom, h3, t3)
    assert eng.get_opt().max_n_chain == opt.max_n_chain
    high, low = int(h3 * (2.0 - 0.333)), int(h3 * 0.333)
    off, mz = eng.sketch(0, n)
    aoff, an = eng.anchors(0, n)
    coff, ch, hoff, hits, foff, fc = eng.chains(0, n, 0.02)
    n_dp = 0
    for i in range(n):
        omz = ho.sketch(st.decode(i), 51, 51, i, 1, ft, 500, 1000)
        assert mz[int(off[i]):int(off[i + 1])].tobytes() == omz.tobytes(), "sketch read %d" % i
        oan = ho.anchors(st, pt, omz, high, low)
        assert an[int(aoff[i]):int(aoff[i + 1])].tobytes() == oan.tobytes(), "anchors read %d" % i
        och, ohits, ofc = ho.lchain(st, i, oan, 0.02, 51, opt.max_n_chain)
        g = ch[int(coff[i]):int(coff[i + 1])]
        assert g.size == och.size, "chain count read %d" % i
        for f, of in (("x_pos_s", "x_pos_s"), ("x_pos_e", "x_pos_e"), ("y_id", "y_id"), ("y_pos_s", "y_pos_s"), ("y_pos_e", "y_pos_e"),
                      ("y_pos_strand", "y_pos_strand"), ("shared_seed", "shared_seed"), ("first_hit", "non_homopolymer_errors"), ("fc_n", "fc_n")):
            assert (g[f] == och[of]).all(), "read %d chain field %s" % (i, f)
        assert hits[int(hoff[i]):int(hoff[i + 1])].tobytes() == ohits.tobytes(), "chain anchors read %d" % i
        gfc = fc[int(foff[i]):int(foff[i + 1])]
        ofc_cat = np.concatenate([ofc[int(c["fc_off"]):int(c["fc_off"]) + int(c["fc_n"])] for c in och]) if och.size else np.zeros(0, np.uint64)
        gfc_cat = np.concatenate([gfc[int(c["fc_off"]):int(c["fc_off"]) + int(c["fc_n"])] for c in g]) if g.size else np.zeros(0, np.uint64)
        assert gfc_cat.tobytes() == ofc_cat.tobytes(), "fake cigars read %d" % i
    c = eng.counters()
    assert c["groups_sequential"] > 0  # the general (DP) path really ran
    woff, win = eng.windows(0, n, 0.02, 0.04, 775)
    n_al = 0
    for i in range(0, n, 7):
        omz = ho.sketch(st.decode(i), 51, 51, i, 1, ft, 500, 1000)
        och, ohits, ofc = ho.lchain(st, i, ho.anchors(st, pt, omz, high, low), 0.02, 51, opt.max_n_chain)
        ow = ho.windows(st, i, och, ofc)
        assert win[int(woff[i]):int(woff[i + 1])].tobytes() == ow.tobytes(), "window pass read %d" % i
        n_al += int((ow["err"] != 2**31 - 1).sum())
    assert n_al > 0
    eng.close()
