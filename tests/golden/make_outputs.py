#!/usr/bin/env python
"""Regenerates tests/golden/outputs.npz: what the UNMODIFIED reference binary itself writes for the golden data sets when it is run the
way a user runs it — `hifiasm -o X -t4 -f0 --write-paf --write-ec reads.fa` (oracle/_ref/hifiasm) — i.e. the files BASELINE.json's
north_star names: X.ovlp.paf (Output_PAF, Assembly.cpp:1673), X.ec.fa (Output_corrected_reads, 884), X.ec.bin, X.ovlp.source.bin,
X.ovlp.reverse.bin.  Stored per data set: size + blake2b-128 digest of each file (the PAF and FASTA are megabytes), the first 4 KiB of
the PAF and of the FASTA for debugging, and a check that the binary's .bin files equal the ones refdump produced for g*.npz (same state).
Only runs in the build container (needs oracle/_ref/hifiasm)."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from hifiasm_b200 import sim, binio  # noqa: E402
import make_golden as mg  # noqa: E402

HIFIASM = os.path.join(ROOT, "oracle", "_ref", "hifiasm")


def dg16(b: bytes) -> np.ndarray:
    return np.frombuffer(hashlib.blake2b(b, digest_size=16).digest(), dtype=np.uint8)


def main():
    if not os.path.exists(HIFIASM):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    arrs = {}
    only = [a for a in sys.argv[1:] if a in mg.DATASETS]   # `make_outputs.py fz3` (tools/fuzz_vs_reference.py): only that set, into mg.OUT_DIR
    for name, (gk, rk) in mg.DATASETS.items():
        if only and name not in only:
            continue
        h1, h2 = sim.sim_genome(**gk)
        reads = sim.sim_reads(h1, h2, **rk)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "reads.fa"); sim.write_fasta(fa, reads)
            pfx = os.path.join(td, "asm")
            subprocess.run([HIFIASM, "-o", pfx, "-t4", "-f0", "--write-paf", "--write-ec", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)  # the graph stages after the overlap stage may find nothing to assemble on toy data: only the stage's files matter
            z = np.load(os.path.join(mg.OUT_DIR, name + ".npz"))
            for suf, key in (("ovlp.paf", "paf"), ("ec.fa", "ecfa"), ("ovlp.source.bin", "src"), ("ovlp.reverse.bin", "rev"), ("ec.bin", "ecbin")):
                b = open("%s.%s" % (pfx, suf), "rb").read()
                arrs["%s_%s_size" % (name, key)] = np.array([len(b)], np.uint64); arrs["%s_%s_dg" % (name, key)] = dg16(b)
                if key in ("paf", "ecfa"):
                    arrs["%s_%s_head" % (name, key)] = np.frombuffer(b[:4096], dtype=np.uint8)
            assert open(pfx + ".ovlp.source.bin", "rb").read() == z["fin_ovlp_source"].tobytes(), name
            assert open(pfx + ".ovlp.reverse.bin", "rb").read() == z["fin_ovlp_reverse"].tobytes(), name
            a = binio.load_ec_bin(pfx + ".ec.bin")
            with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
                tf.write(z["pre_ec"].tobytes()); tf.flush(); b = binio.load_ec_bin(tf.name)
            assert (a.length == b.length).all() and (binio.canonical_packed(a) == binio.canonical_packed(b)).all() and a.name_blob == b.name_blob, name  # ec.bin: identical after masking the pad bytes the reference leaves uninitialised
            print(name, {k: int(arrs["%s_%s_size" % (name, k)][0]) for k in ("paf", "ecfa", "src", "rev", "ecbin")})
    out = os.path.join(mg.OUT_DIR, "outputs.npz")
    np.savez_compressed(out, **arrs)
    print("->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
