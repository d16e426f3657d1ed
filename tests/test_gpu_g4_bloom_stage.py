"""GPU tests of the paths round 1 wrote after its GPU budget was spent; all of them run green on a B200 since round 2 (profiles/r2_gpu_all_tests.log):

* golden set g4 — 40 kb reads with tandem-repeat blocks, where accepted overlaps keep unaligned windows of >= 512 bp on both reads and the
  reference re-seeds them (rechain_aln_hc, Correct.cpp:17669; hb_ecrechain.cuh / k_ecb_rechain).  Round 1's two stage-level failures on this set
  were NOT the rescue: the anchor grouping kernel keyed its groups by (target, strand), so a target with anchors on both strands (tandem / palindromic
  repeats) was chained as two groups instead of one call with both strand blocks (anchor.cpp:1928-1934) and produced extra one-anchor chains.
* the filter table behind the reference's Bloom filter (opt.bf_shift)
* stage.run_stage (FASTA in -> the reference's files out)
* dist.cal_ec_r_sharded with one rank"""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from goldenlib import Golden  # noqa: E402

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def hb():
    import hifiasm_b200
    return hifiasm_b200


def test_stages_raw_g4(hb):
    """every stage of an EC round on the raw reads of g4 through the C-ABI, incl. the rescue (no overlap may keep need_rechain)"""
    import test_gpu_parity as tp
    g = Golden("g4")
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    tp._check_stages(g, eng, "raw", g.raw)
    assert eng.profile().get("k_ecb_rechain", (0, 0.0))[0] >= 1, "the rescue kernel did not launch"
    eng.close()


def test_stages_final_g4(hb):
    """the same on the corrected reads with the previous round's lists (exact shortcut + rescue)"""
    import test_gpu_parity as tp
    g = Golden("g4")
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    tp._check_stages(g, eng, "final", g.pre)
    eng.close()


def test_whole_stage_from_raw_reads_g4(hb, tmp_path):
    """raw reads -> three EC rounds -> final pass on the device, byte-identical ovlp.*.bin, with the rescue on the path"""
    import test_gpu_round as tr
    tr.test_whole_stage_from_raw_reads(hb, "g4", tmp_path)


# ---- the filter table behind the reference's Bloom filter (opt.bf_shift = hifiasm -f): CPU side in tests/test_bloom.py
import numpy as np  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.mark.parametrize("shift", [0, 21, 22, 24])
def test_gpu_bloom(hb, shift):
    import make_bloom
    from goldenlib import GOLDEN
    from hifiasm_b200 import sim
    z = np.load(os.path.join(GOLDEN, "bloom.npz"))
    flat, boff, ln, npos, noff = sim.pack_reads(make_bloom.reads())
    eng = hb.Engine(0)
    eng.set_opt(bf_shift=shift)
    eng.upload_reads(ln, flat, boff, npos, noff)
    hom = eng.ft_gen()
    key, cnt = z["f%d_key" % shift], z["f%d_cnt" % shift]
    assert hom == int(z["f%d_hom" % shift][0]) and eng.ft_size() == key.size
    assert (eng.ft_cnt(key) == cnt).all()
    rng = np.random.default_rng(1)
    assert (eng.ft_cnt(rng.integers(0, 2**63, 1000, dtype=np.uint64)) == 0).all()
    eng.close()


# ---- FASTA in, the reference's files out: hifiasm_b200.stage.run_stage (every step a C-ABI call that is GPU-checked on its own; this glue is new)
@pytest.mark.parametrize("name", ["g1", "g3"])
def test_run_stage_fasta_to_files(hb, name, tmp_path):
    import hashlib
    import make_golden as mg
    from goldenlib import GOLDEN
    from hifiasm_b200 import sim, binio, stage
    gk, rk = mg.DATASETS[name]
    h1, h2 = sim.sim_genome(**gk)
    fa = str(tmp_path / "reads.fa"); sim.write_fasta(fa, sim.sim_reads(h1, h2, **rk))
    pfx = str(tmp_path / "asm")
    info = stage.run_stage(fa, pfx)
    z = np.load(os.path.join(GOLDEN, "outputs.npz"))
    for suf, key in (("ovlp.paf", "paf"), ("ec.fa", "ecfa"), ("ovlp.source.bin", "src"), ("ovlp.reverse.bin", "rev")):
        b = open("%s.%s" % (pfx, suf), "rb").read()
        assert len(b) == int(z["%s_%s_size" % (name, key)][0]) and (np.frombuffer(hashlib.blake2b(b, digest_size=16).digest(), np.uint8) == z["%s_%s_dg" % (name, key)]).all(), suf
    g = Golden(name); mine = binio.load_ec_bin(pfx + ".ec.bin")
    assert (mine.length == g.pre.length).all() and (binio.canonical_packed(mine) == binio.canonical_packed(g.pre)).all() and mine.name_blob == g.pre.name_blob
    assert mine.total_reads_bases == g.pre.total_reads_bases and info["reads"] == g.pre.n and os.path.getsize(pfx + ".ec.bin") == int(z["%s_ecbin_size" % name][0])


# ---- the EC rounds sharded over ranks (hifiasm_b200.dist.cal_ec_r_sharded): with one process it must equal hb_cal_ec_r; N > 1 runs under
# torchrun with tools/ec_sharded_check.py (the exchange helpers are covered by tests/test_dist_gloo.py on the CPU)
def test_sharded_round_single_rank_equals_cal_ec_r(hb):
    import roundlib
    from hifiasm_b200 import binio, dist as hdist
    g = Golden("g1"); rd = roundlib.Rounds("g1")
    eng = hb.Engine(0)
    eng.upload_store(g.raw)
    hom = eng.ft_gen(); eng.update_cov(hom)
    n = g.raw.n
    src = np.zeros(0, binio.MA_MEM); soff = np.zeros(n + 1, np.uint64)
    for K in range(3):
        hom_k, het_k = eng.pt_gen(); eng.set_opt(hom_cov=hom_k, het_cov=het_k)
        r = hdist.cal_ec_r_sharded(eng, K, 1 if K == 2 else 0, src, soff)
        assert not r["status"].any()
        p = rd.params(K)
        assert (r["tot_b"], r["tot_e"]) == (int(p["tot_b"]), int(p["tot_e"]))
        src, soff, rev, roff = r["src"], r["src_off"], r["rev"], r["rev_off"]
        assert (roundlib.list_digests(src, soff, 0) == rd.digest(K, "post_src")).all() and (roundlib.list_digests(rev, roff, 1) == rd.digest(K, "post_rev")).all()
    assert (roundlib.reads_digests(eng.download_reads()) == roundlib.reads_digests(g.pre)).all()
    eng.close()
