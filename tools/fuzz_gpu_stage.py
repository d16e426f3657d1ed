#!/usr/bin/env python
"""End-to-end differential fuzzing ON THE GPU BOX: fresh read sets -> FASTA -> (a) the unmodified reference binary oracle/_ref/hifiasm
(prebuilt, travels with the snapshot) with --write-paf --write-ec on the host cores, (b) hifiasm_b200.stage.run_stage on the GPU -> the
files must be byte-identical (.ovlp.paf, .ec.fa, .ovlp.source.bin, .ovlp.reverse.bin; .ec.bin up to the reference's uninitialised pad bytes).
    usage (under gpurun): python tools/fuzz_gpu_stage.py [first_seed [n_configs]]      -> one line per configuration, exit code 1 on a mismatch"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from hifiasm_b200 import sim, binio, stage  # noqa: E402
import fuzz_vs_reference as fz  # noqa: E402


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 5000; n = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    ref = os.path.join(ROOT, "oracle", "_ref", "hifiasm"); bad = []
    if not os.path.exists(ref):
        sys.exit("oracle/_ref/hifiasm is missing (built where /root/reference exists: make -C oracle ref)")
    for i in range(n):
        g, r = fz.config(i, seed0 + i)
        h1, h2 = sim.sim_genome(**g); reads = sim.sim_reads(h1, h2, **r)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "reads.fa"); sim.write_fasta(fa, reads)
            subprocess.run([ref, "-o", os.path.join(td, "ref"), "-t%d" % min(32, os.cpu_count() or 1), "-f0", "--write-paf", "--write-ec", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
            try:
                info = stage.run_stage(fa, os.path.join(td, "gpu")); diff = []
            except Exception as ex:  # a read the engine could not finish, or an error of the library
                info = {}; diff = ["run_stage: %r" % (ex,)]
            for suf in ("ovlp.paf", "ec.fa", "ovlp.source.bin", "ovlp.reverse.bin"):
                if not diff or not diff[0].startswith("run_stage"):
                    a, b = open(os.path.join(td, "ref." + suf), "rb").read(), open(os.path.join(td, "gpu." + suf), "rb").read()
                    if a != b:
                        diff.append("%s (%d vs %d bytes)" % (suf, len(a), len(b)))
            if not diff:
                x, y = binio.load_ec_bin(os.path.join(td, "ref.ec.bin")), binio.load_ec_bin(os.path.join(td, "gpu.ec.bin"))
                if not ((x.length == y.length).all() and (binio.canonical_packed(x) == binio.canonical_packed(y)).all() and x.name_blob == y.name_blob and (x.hom_cov, x.het_cov) == (y.hom_cov, y.het_cov)):
                    diff.append("ec.bin")
            print("config %d (seed %d, kind %d): %d reads, %s" % (i, seed0 + i, i % 11, len(reads), "identical" if not diff else "DIFFERENT: " + "; ".join(diff)), flush=True)
            if diff:
                bad.append(seed0 + i)
    print("different:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
