"""Parity at benchmark shape, on the GPU box: the whole stage (stage.run_stage: FASTA -> filter table -> 3 EC rounds -> final pass -> files) against
the UNMODIFIED reference binary (oracle/_ref/hifiasm, prebuilt; -f0 --write-paf --write-ec) on the same seeded read set of SURVEY.md §8(d)'s shape:
diploid genome with 0.1 % SNPs, 5 % of it in 50-copy 5 kb repeat families (both orientations), 30x reads of 15 kb with 0.2 % errors, N bases.
.ovlp.paf / .ec.fa / .ovlp.source.bin / .ovlp.reverse.bin are compared byte for byte, .ec.bin up to the pad bytes the reference leaves undefined.
HB_SCALE_MB picks the genome size (default 4 Mb = 8000 reads: ~1 minute with the reference run; profiles/ holds the log of a 20 Mb run)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = [pytest.mark.gpu]


def _cmp_files(td, a, b):
    from hifiasm_b200 import binio
    diff = []
    for suf in ("ovlp.paf", "ec.fa", "ovlp.source.bin", "ovlp.reverse.bin"):
        x, y = open(os.path.join(td, a + "." + suf), "rb").read(), open(os.path.join(td, b + "." + suf), "rb").read()
        if x != y:
            diff.append("%s (%d vs %d bytes)" % (suf, len(x), len(y)))
    x, y = binio.load_ec_bin(os.path.join(td, a + ".ec.bin")), binio.load_ec_bin(os.path.join(td, b + ".ec.bin"))
    if not ((x.length == y.length).all() and (binio.canonical_packed(x) == binio.canonical_packed(y)).all() and x.name_blob == y.name_blob and (x.hom_cov, x.het_cov) == (y.hom_cov, y.het_cov)):
        diff.append("ec.bin")
    return diff


@pytest.fixture(scope="module")
def ref_run(tmp_path_factory):
    """the seeded read set and the reference binary's files for it (one run for the module)"""
    import simgen
    ref = os.path.join(ROOT, "oracle", "_ref", "hifiasm")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/hifiasm is not built (make -C oracle ref where /root/reference exists)")
    mb = float(os.environ.get("HB_SCALE_MB", "4")); td = str(tmp_path_factory.mktemp("scale"))
    fa = os.path.join(td, "reads.fa")
    rs = simgen.make(mb, 30, seed=int(os.environ.get("HB_SCALE_SEED", "77")), n_rate=0.0002, fasta=fa)
    thr = len(os.sched_getaffinity(0))
    p = subprocess.run([ref, "-o", os.path.join(td, "ref"), "-t%d" % thr, "-f0", "--write-paf", "--write-ec", fa], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return mb, td, fa, rs


# default: the batches the anchor budget gives at this size (two), on two lanes.  Then many small batches on two lanes (each lane's host thread takes
# batches in turn on its own stream; what they add to the lists is added in batch order) and on one lane: the files must not depend on either.
@pytest.mark.parametrize("budget,lanes", [(None, None), ("6000000", "2"), ("6000000", "1")])
def test_whole_stage_files_equal_reference_binary_at_scale(ref_run, monkeypatch, budget, lanes):
    from hifiasm_b200 import stage
    mb, td, fa, rs = ref_run
    if budget:
        monkeypatch.setenv("HB_ANCHOR_BUDGET", budget)
    if lanes:
        monkeypatch.setenv("HB_LANES", lanes)
    if lanes == "1":
        monkeypatch.setenv("HB_NO_SKETCH_REUSE", "1")        # (this variant also sketches the query reads again instead of reading the index build's sketch)
    out = "gpu_%s_%s" % (budget, lanes)
    info = stage.run_stage(fa, os.path.join(td, out))
    assert info["reads"] == rs.n and info["bases"] == rs.bases
    diff = _cmp_files(td, "ref", out)
    print("scale parity (budget %s, lanes %s): %g Mb genome, %d reads, %d bases, corrected per round %s, overlaps %d + %d: %s" % (budget, lanes, mb, rs.n, rs.bases, info["corrected_bases"], info["overlaps_src"], info["overlaps_rev"], "identical" if not diff else diff))
    assert not diff, diff
