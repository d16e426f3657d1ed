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


@pytest.mark.parametrize("name", ["g1", "g2", "g3", "g4"])
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
