"""GPU parity of the steps that close an error-correction round, through the C-ABI, against the states the
unmodified reference passes through in its three rounds (tests/golden/g*_rounds.npz; see tests/golden/make_rounds.py).

The rounds are chained on the device: the read store of round K + 1 is the one hb_ec_apply / hb_ec_post_rev leave in HBM in round K,
the index is rebuilt on it (hb_pt_gen), and the previous round's lists feed the exact shortcut of the alignment stage.  The edit
scripts are the reference's (the consensus, row a14, is not on the device yet)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from goldenlib import Golden  # noqa: E402
import roundlib  # noqa: E402
from hifiasm_b200 import binio  # noqa: E402


@pytest.fixture(scope="module")
def hb():
    import hifiasm_b200
    return hifiasm_b200


@pytest.mark.parametrize("name", ["g1", "g2", "g3"])
def test_round_closing_steps(hb, name):
    """rows a16-a18 with the reference's lists as input: corrected reads, remapped exact intervals, reverse complement + flipped lists"""
    g = Golden(name); rd = roundlib.Rounds(name)
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    for K in range(3):
        scc, scc_off = rd.scc(K)
        src, soff, fc, ab = rd.hap(K, "src"); rev, roff, _, _ = rd.hap(K, "rev")
        eng.ec_stage_scc(scc, scc_off)
        n_changed, tot = eng.ec_apply()
        sl = eng.download_reads()
        assert (roundlib.reads_digests(sl) == rd.digest(K, "sl_reads")).all(), "round %d: corrected reads" % K
        assert tot == int(sl.length.sum()) and 0 <= n_changed <= sl.n
        upd, n_exact, n_inexact = eng.ec_update_paf(binio.disk_to_mem(src), soff)
        assert (roundlib.list_digests(upd, soff, 0) == rd.digest(K, "upd_src")).all(), "round %d: updated paf" % K
        assert n_exact + n_inexact == src.size and n_exact == int((upd["el"] == 1).sum())
        if K < 2:
            psrc, psoff, prev, proff = eng.ec_post_rev(upd, soff, binio.disk_to_mem(rev), roff)
        else:
            psrc, psoff, prev, proff = upd, soff, binio.disk_to_mem(rev), roff
        post = eng.download_reads()
        assert (roundlib.reads_digests(post) == rd.digest(K, "post_reads")).all(), "round %d: reads after the round" % K
        assert (roundlib.list_digests(psrc, psoff, 0) == rd.digest(K, "post_src")).all(), "round %d: paf after the round" % K
        assert (roundlib.list_digests(prev, proff, 1) == rd.digest(K, "post_rev")).all(), "round %d: reverse_paf after the round" % K
    assert (roundlib.reads_digests(eng.download_reads()) == roundlib.reads_digests(g.pre)).all()
    eng.close()


def _cmp_lists(tag, K, got, goff, want, woff, is_rev, skip):
    bad = []
    for i in range(goff.size - 1):
        if skip[i]:
            continue
        a = roundlib.canon_list(got[int(goff[i]):int(goff[i + 1])], is_rev); b = roundlib.canon_list(want[int(woff[i]):int(woff[i + 1])], is_rev)
        if a.tobytes() != b.tobytes():
            bad.append(i)
    assert not bad, "round %d %s: %d reads differ, first %s" % (K, tag, len(bad), bad[:5])


@pytest.mark.parametrize("name", ["g1", "g3"])
def test_round_chained_on_device(hb, name):
    """a whole round on the device, three rounds chained: index -> alignment stage (with the previous round's exact shortcut) -> phasing -> dedup ->
    window consensus (row a14: edit scripts) -> paf[] / reverse_paf[] / is_fully_corrected / is_abnormal (row a15) -> rows a16-a18.
    The consensus includes the graph path (cns_gen_full) for the stretches the vote cannot settle.  Reads whose alignment wanted rechain_aln_hc
    (status bit 2; none in these sets) are reported by the engine and compared nowhere; after the comparison the chain continues with the reference's
    scripts and lists so that every round starts from the reference's state."""
    g = Golden(name); rd = roundlib.Rounds(name)
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)   # the filter table of the raw reads serves all rounds (Assembly.cpp:2083)
    n = g.raw.n
    prev_src = np.zeros(0, binio.MA_MEM); prev_off = np.zeros(n + 1, np.uint64)
    n_checked = 0
    for K in range(3):
        p = rd.params(K)
        hom_k, het_k = eng.pt_gen()
        assert (hom_k, het_k) == (int(p["hom_cov"]), int(p["het_cov"])), "round %d: coverage peaks" % K
        eng.set_opt(hom_cov=hom_k, het_cov=het_k)
        eng.ec_stage_prev(prev_src, prev_off)
        r = eng.ec_round(0, n, 0.02, 0.04, 775, use_prev=1)
        scc, scc_off = rd.scc(K)
        src, soff, fc, ab = rd.hap(K, "src"); rev, roff, _, _ = rd.hap(K, "rev")
        ok = r["status"] == 0
        assert int((r["status"] & (1 | 2 | 8)).sum()) == 0, "round %d: reads without an edit script (status %s)" % (K, sorted(set(int(x) for x in r["status"])))
        bad = [i for i in range(n) if ok[i] and r["scc"][int(r["scc_off"][i]):int(r["scc_off"][i + 1])].tobytes() != scc[int(scc_off[i]):int(scc_off[i + 1])].tobytes()]
        assert not bad, "round %d edit scripts: %d reads differ, first %s" % (K, len(bad), bad[:5])
        _cmp_lists("paf", K, r["src"], r["src_off"], src, soff, 0, ~ok)
        _cmp_lists("reverse_paf", K, r["rev"], r["rev_off"], rev, roff, 1, (r["status"] & 4) != 0)   # the reverse list does not depend on the script
        assert (r["is_fully_corrected"][ok] == fc[ok]).all() and (r["is_abnormal"][ok] == ab[ok]).all(), "round %d: is_fully_corrected / is_abnormal" % K
        if ok.all():
            assert r["n_corrected"] == int(p["tot_e"])
        n_checked += int(ok.sum())
        # close the round on the device from the reference's scripts and lists
        eng.ec_stage_scc(scc, scc_off)
        eng.ec_apply()
        upd, _, _ = eng.ec_update_paf(binio.disk_to_mem(src), soff)
        if K < 2:
            prev_src, prev_off, _, _ = eng.ec_post_rev(upd, soff, binio.disk_to_mem(rev), roff)
        else:
            prev_src, prev_off = upd, soff
        assert (roundlib.reads_digests(eng.download_reads()) == rd.digest(K, "post_reads")).all(), "round %d: reads after the round" % K
    assert n_checked > 2 * n
    eng.close()


def test_round_lists_from_staged_scripts(hb):
    """row a15 alone: the round's lists from edit scripts the caller stages (here the reference's of round 0), as a shard of a multi-GPU run does
    with the scripts gathered from the other ranks"""
    g = Golden("g2"); rd = roundlib.Rounds("g2")
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    hom_k, het_k = eng.pt_gen(); eng.set_opt(hom_cov=hom_k, het_cov=het_k)
    n = g.raw.n
    scc, scc_off = rd.scc(0)
    eng.ec_stage_scc(scc, scc_off)
    eng.ec_stage_prev(np.zeros(0, binio.MA_MEM), np.zeros(n + 1, np.uint64))
    poff, P = eng.ec_phase(0, n, 0.02, 0.04, 775)
    skip = np.array([bool(P[int(poff[i]):int(poff[i + 1])]["need_rechain"].any()) for i in range(n)])
    so, S, ro, Rv, f_ec, f_ab = eng.ec_round_lists(0, n, 0.02, 0.04, 775, use_prev=1)
    src, soff, fc, ab = rd.hap(0, "src"); rev, roff, _, _ = rd.hap(0, "rev")
    _cmp_lists("paf", 0, S, so, src, soff, 0, skip)
    _cmp_lists("reverse_paf", 0, Rv, ro, rev, roff, 1, skip)
    keep = ~skip
    assert keep.sum() > n // 2 and (f_ec[keep] == fc[keep]).all() and (f_ab[keep] == ab[keep]).all()
    eng.close()


@pytest.mark.parametrize("name", ["g1", "g2", "g3"])
def test_whole_stage_from_raw_reads(hb, name, tmp_path):
    """the north-star property on the golden sets: raw reads in, the whole overlap / error-correction stage on the device with NO reference state
    fed back — filter table, then three EC rounds (index, alignment, phasing, consensus, lists, apply, update, reverse), then the final index and
    cal_ov_r — and out come the reference's corrected reads and byte-identical ovlp.source.bin / ovlp.reverse.bin"""
    g = Golden(name); rd = roundlib.Rounds(name)
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    n = g.raw.n
    src = np.zeros(0, binio.MA_MEM); soff = np.zeros(n + 1, np.uint64); rev = src.copy(); roff = soff.copy()
    for K in range(3):
        hom_k, het_k = eng.pt_gen(); eng.set_opt(hom_cov=hom_k, het_cov=het_k)      # ha_ec: index of the round on the current reads (Assembly.cpp:1007)
        r = eng.cal_ec_r(K, 1 if K == 2 else 0, src, soff)                           # is_sv only in the last round (Assembly.cpp:1021)
        assert not r["status"].any(), "round %d: reads the engine could not finish: %s" % (K, np.nonzero(r["status"])[0][:8])
        p = rd.params(K)
        assert (r["tot_b"], r["tot_e"]) == (int(p["tot_b"]), int(p["tot_e"])), "round %d: the [M::pec] counters" % K
        src, soff, rev, roff = r["src"], r["src_off"], r["rev"], r["rev_off"]
        assert (roundlib.list_digests(src, soff, 0) == rd.digest(K, "post_src")).all(), "round %d: paf" % K
        assert (roundlib.list_digests(rev, roff, 1) == rd.digest(K, "post_rev")).all(), "round %d: reverse_paf" % K
        h_src, _, fc, ab = rd.hap(K, "src")
        assert (r["is_fully_corrected"] == fc).all() and (r["is_abnormal"] == ab).all(), "round %d: read flags" % K
        assert r["n_exact"] + r["n_inexact"] == h_src.size
    reads = eng.download_reads()
    assert (roundlib.reads_digests(reads) == roundlib.reads_digests(g.pre)).all(), "corrected reads (ec.bin after canonicalisation)"
    hom_f, het_f = eng.pt_gen(); eng.set_opt(hom_cov=hom_f, het_cov=het_f)
    out0, oo0, out1, oo1, stat = eng.cal_ov_r(src, soff, rev, roff)
    p0, p1 = str(tmp_path / "src.bin"), str(tmp_path / "rev.bin")
    f0, fo0, fc0, ab0 = g.fin_src; f1, fo1, fc1, ab1 = g.fin_rev
    binio.write_ovlp_bin(p0, out0, oo0, fc0, ab0); binio.write_ovlp_bin(p1, out1, oo1, fc1, ab1)   # the two read flags are those of round 3 (checked in test_round_chained_on_device)
    assert open(p0, "rb").read() == g.z["fin_ovlp_source"].tobytes(), "ovlp.source.bin"
    assert open(p1, "rb").read() == g.z["fin_ovlp_reverse"].tobytes(), "ovlp.reverse.bin"
    # ... and the files a user of the reference asks for with --write-paf --write-ec, from the DEVICE's state through the library's own
    # writers, byte-identical to what the unmodified reference binary wrote for the same reads (tests/golden/outputs.npz)
    import test_outputs
    reads.names, reads.name_blob, reads.name_index = g.raw.names, g.raw.name_blob, g.raw.name_index   # read names travel beside the store (All_reads.name)
    test_outputs.check_outputs(name, tmp_path, reads, out0, oo0, fc0, ab0, out1, oo1, fc1, ab1)
    eng.close()


def test_stale_index_is_refused(hb):
    """a pass after the read store changed (hb_ec_apply / hb_ec_post_rev / a new upload) without a fresh hb_pt_gen must fail loudly, not chain stale positions"""
    from hifiasm_b200.engine import HBError
    g = Golden("g1"); rd = roundlib.Rounds("g1")
    eng = hb.Engine(0)
    eng.upload_store(g.raw); eng.update_cov(eng.ft_gen()); h, t = eng.pt_gen(); eng.set_opt(hom_cov=h, het_cov=t)
    eng.chains(0, 8, 0.02)                       # fine: the index describes the store
    scc, scc_off = rd.scc(0); eng.ec_stage_scc(scc, scc_off); eng.ec_apply()
    with pytest.raises(HBError):
        eng.chains(0, 8, 0.02)
    eng.pt_gen(); eng.chains(0, 8, 0.02)         # rebuilt: fine again
    eng.upload_store(g.raw)
    with pytest.raises(HBError):
        eng.chains(0, 8, 0.02)
    eng.close()
