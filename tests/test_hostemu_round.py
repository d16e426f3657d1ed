"""CPU checks of the steps that close an error-correction round: the bodies of the product's kernels
(hifiasm_b200/csrc/hb_ecround.cuh), compiled as host code (tests/hostemu — test infrastructure), against the states the unmodified
reference passes through in each of its three rounds (tests/golden/g*_rounds.npz, made by tests/golden/make_rounds.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, os.path.join(HERE, "hostemu"))
sys.path.insert(0, HERE)

from hifiasm_b200 import binio  # noqa: E402
import goldenlib  # noqa: E402
import roundlib  # noqa: E402
import emu  # noqa: E402


@pytest.mark.parametrize("name", os.environ["HB_GOLDEN_NAMES"].split(",") if os.environ.get("HB_GOLDEN_NAMES") else ["g1", "g2", "g3", "g4"])
def test_round_closing_steps(name):
    """rounds 0, 1, 2 chained: the read store of round K+1 is the one these steps produce in round K"""
    g = goldenlib.Golden(name); rd = roundlib.Rounds(name)
    er = emu.Reads(g.raw)
    n = er.n
    for K in range(3):
        scc, scc_off = rd.scc(K)
        src, soff, fc, ab = rd.hap(K, "src"); rev, roff, _, _ = rd.hap(K, "rev")
        # a16: worker_sl_ec
        sl = emu.ec_apply(er, scc, scc_off)
        assert (roundlib.reads_digests(sl.export()) == rd.digest(K, "sl_reads")).all(), "round %d: corrected reads" % K
        # a17: worker_update_dc_ec on the corrected reads
        upd, n_exact = emu.ec_update(sl, binio.disk_to_mem(src), soff, scc, scc_off)
        assert (roundlib.list_digests(upd, soff, 0) == rd.digest(K, "upd_src")).all(), "round %d: updated paf" % K
        assert 0 < n_exact <= int((src["el"] == 1).sum())
        # a18: worker_hap_post_rev — rounds 0 and 1 only (cal_ec_r: not in the last round when the number of rounds is odd, ecovlp.cpp:6290)
        if K < 2:
            post = emu.ec_rc(sl)
            psrc, psoff = emu.ec_flip(sl, upd, soff); prev, proff = emu.ec_flip(sl, binio.disk_to_mem(rev), roff)
        else:
            post = sl; psrc, psoff = upd, soff; prev, proff = binio.disk_to_mem(rev), roff
        assert (roundlib.reads_digests(post.export()) == rd.digest(K, "post_reads")).all(), "round %d: reads after the round" % K
        assert (roundlib.list_digests(psrc, psoff, 0) == rd.digest(K, "post_src")).all(), "round %d: paf after the round" % K
        assert (roundlib.list_digests(prev, proff, 1) == rd.digest(K, "post_rev")).all(), "round %d: reverse_paf after the round" % K
        er = post
    # the reads after round 2 are the ones the final overlap pass starts from
    assert (roundlib.reads_digests(er.export()) == roundlib.reads_digests(g.pre)).all()
