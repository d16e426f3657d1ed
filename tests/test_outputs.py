"""The stage's on-disk products through the library's native writers (csrc/hostio.cu: hb_write_paf / hb_write_ec_fa / hb_write_ovlp_bin /
hb_write_ec_bin, host code — no GPU) against what the UNMODIFIED reference binary writes when a user runs it with --write-paf --write-ec
(tests/golden/outputs.npz, made by tests/golden/make_outputs.py from oracle/_ref/hifiasm): byte-identical X.ovlp.paf, X.ec.fa,
X.ovlp.source.bin, X.ovlp.reverse.bin — the files BASELINE.json's north_star names.  The inputs here are the reference's own final state
(golden reads + lists); tests/test_gpu_round.py feeds the same writers with the state the DEVICE produced from raw reads."""
import hashlib
import os

import numpy as np
import pytest

from goldenlib import Golden, GOLDEN
from hifiasm_b200 import binio


def _dg(path):
    b = open(path, "rb").read()
    return len(b), np.frombuffer(hashlib.blake2b(b, digest_size=16).digest(), dtype=np.uint8)


def check_outputs(name, tmp_path, reads, src, soff, fc0, ab0, rev, roff, fc1, ab1):
    """writes the four files from (corrected reads, final lists) and compares them with the reference binary's"""
    z = np.load(os.path.join(GOLDEN, "outputs.npz"))
    p = {k: str(tmp_path / k) for k in ("paf", "ecfa", "src", "rev")}
    binio.native_write_paf(p["paf"], reads, src, soff)
    binio.native_write_ec_fa(p["ecfa"], reads)
    binio.native_write_ovlp_bin(p["src"], src, soff, fc0, ab0)
    binio.native_write_ovlp_bin(p["rev"], rev, roff, fc1, ab1)
    for k in ("paf", "ecfa", "src", "rev"):
        size, dg = _dg(p[k])
        if k in ("paf", "ecfa"):
            head = open(p[k], "rb").read(4096)
            assert head == z["%s_%s_head" % (name, k)].tobytes(), "%s: first bytes of %s" % (name, k)
        assert size == int(z["%s_%s_size" % (name, k)][0]) and (dg == z["%s_%s_dg" % (name, k)]).all(), "%s: %s differs from the reference's file" % (name, k)


@pytest.mark.parametrize("name", os.environ["HB_GOLDEN_NAMES"].split(",") if os.environ.get("HB_GOLDEN_NAMES") else ["g1", "g2", "g3", "g4"])
def test_native_writers_match_the_reference_binary(name, tmp_path):
    g = Golden(name)
    f0, fo0, fc0, ab0 = g.fin_src; f1, fo1, fc1, ab1 = g.fin_rev
    check_outputs(name, tmp_path, g.pre, f0, fo0, fc0, ab0, f1, fo1, fc1, ab1)
    # ec.bin: the writer reproduces the dump it was loaded from byte for byte (header, N lists, lengths, packed reads incl. their pad
    # bytes, names, name index, trio flags, coverage peaks); against the BINARY's ec.bin only the never-initialised pad bytes differ
    e = str(tmp_path / "ec.bin")
    binio.native_write_ec_bin(e, g.pre)
    assert open(e, "rb").read() == g.z["pre_ec"].tobytes()
    z = np.load(os.path.join(GOLDEN, "outputs.npz"))
    assert os.path.getsize(e) == int(z["%s_ecbin_size" % name][0])
    # the numpy test tooling agrees with the native writers
    a, b = str(tmp_path / "a.paf"), str(tmp_path / "b.bin")
    binio.write_paf(a, g.pre, f0, fo0); binio.write_ovlp_bin(b, f0, fo0, fc0, ab0)
    assert open(a, "rb").read() == open(str(tmp_path / "paf"), "rb").read() and open(b, "rb").read() == open(str(tmp_path / "src"), "rb").read()


def test_writers_report_io_errors(tmp_path):
    g = Golden("g1"); f0, fo0, fc0, ab0 = g.fin_src
    with pytest.raises(IOError):
        binio.native_write_paf(str(tmp_path / "no" / "such" / "dir" / "x.paf"), g.pre, f0, fo0)
    bad = binio.disk_to_mem(f0[:1].copy()); bad["tn"] = g.pre.n + 5
    with pytest.raises(IOError):
        binio.native_write_paf(str(tmp_path / "x.paf"), g.pre, bad, np.array([0, 1] + [1] * (g.pre.n - 1), np.uint64))


@pytest.mark.parametrize("name", ["g1", "g3"])
def test_native_ingest_matches_the_reference(name, tmp_path):
    """hb_readset_load (FASTA / gzip FASTA / FASTQ -> the All_reads layout) against the reference's own R_INF right after it parsed the same
    file (the golden raw ec.bin): lengths, 2-bit reads, N sites, names, and the header fields write_All_reads stores"""
    import gzip
    import sys
    sys.path.insert(0, os.path.join(GOLDEN))
    import make_golden as mg
    from hifiasm_b200 import sim
    g = Golden(name); gk, rk = mg.DATASETS[name]
    h1, h2 = sim.sim_genome(**gk); reads = sim.sim_reads(h1, h2, **rk)
    fa = str(tmp_path / "reads.fa"); sim.write_fasta(fa, reads)
    gz = str(tmp_path / "reads.fa.gz"); open(gz, "wb").write(gzip.compress(open(fa, "rb").read()))
    fq = str(tmp_path / "reads.fq")   # FASTQ with wrapped lines and a description after the name
    with open(fq, "w") as f:
        for nm, r in zip(g.raw.names, reads):
            s = np.frombuffer(b"ACGTN", np.uint8)[r].tobytes().decode()
            f.write("@%s some description\n%s\n%s\n+\n%s\n" % (nm, s[:70], s[70:], "I" * len(s)))
    for path in (fa, gz, fq):
        rs = binio.native_load_reads(path)
        assert rs.n == g.raw.n and (rs.length == g.raw.length).all() and rs.names == g.raw.names
        assert (rs.n_off == g.raw.n_off).all() and (rs.n_pos == g.raw.n_pos).all()
        assert (binio.canonical_packed(rs) == binio.canonical_packed(g.raw)).all()
        assert (rs.index_size, rs.name_index_size, rs.total_reads_bases, rs.name_blob) == (g.raw.index_size, g.raw.name_index_size, g.raw.total_reads_bases, g.raw.name_blob)
        assert (rs.name_index[:rs.n + 1] == g.raw.name_index[:rs.n + 1]).all()
    # two files = one read set in file order; adapter trimming drops what becomes empty
    two = binio.native_load_reads([fa, gz])
    assert two.n == 2 * g.raw.n and (two.length[g.raw.n:] == g.raw.length).all()
    cut = binio.native_load_reads(fa, adapter_len=int(g.raw.length.min()) // 2 + 1)
    assert cut.n < g.raw.n and (cut.length > 0).all()
    with pytest.raises(IOError):
        binio.native_load_reads(str(tmp_path / "missing.fa"))
