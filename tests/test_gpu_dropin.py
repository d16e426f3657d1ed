"""The compiled drop-in on the GPU box: oracle/_ref/hifiasm_b200_dropin = the reference's own main() / ha_assemble() and all its
objects, with cal_ec_r / cal_ov_r (ecovlp.h:12,14) replaced at link time by shim/hb_shim.cpp over libhifiasm_b200.so.  Run with the stock binary's
flags on the same FASTA, it must write the same files: option parsing, ingest, R_INF ownership (malloc'ed buffers, size >= length), the [M::pec] /
ha_print_ovlp_stat_0 lines and the reference's own writers all run around the GPU stage."""
import os
import re
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, HERE)

pytestmark = [pytest.mark.gpu]


def test_dropin_binary_writes_the_reference_files(tmp_path):
    import simgen
    import test_gpu_scale as tg
    ref = os.path.join(ROOT, "oracle", "_ref", "hifiasm"); drop = os.path.join(ROOT, "oracle", "_ref", "hifiasm_b200_dropin")
    if not (os.path.exists(ref) and os.path.exists(drop)):
        pytest.skip("oracle/_ref/hifiasm[_b200_dropin] are not built (make -C oracle ref where /root/reference exists)")
    td = str(tmp_path); fa = os.path.join(td, "reads.fa")
    rs = simgen.make(float(os.environ.get("HB_DROPIN_MB", "2")), 30, seed=91, n_rate=0.0002, fasta=fa)
    thr = len(os.sched_getaffinity(0))
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = ":".join(p for p in ("/usr/local/cuda/lib64", env.get("LD_LIBRARY_PATH", "")) if p)
    try:   # the CUDA runtime the library was linked against ships with torch's wheels
        import nvidia.cuda_runtime
        env["LD_LIBRARY_PATH"] = os.path.join(os.path.dirname(nvidia.cuda_runtime.__file__), "lib") + ":" + env["LD_LIBRARY_PATH"]
    except Exception:
        pass
    a = subprocess.run([ref, "-o", os.path.join(td, "ref"), "-t%d" % thr, "-f0", "--write-paf", "--write-ec", fa], capture_output=True, text=True)
    b = subprocess.run([drop, "-o", os.path.join(td, "gpu"), "-t%d" % thr, "-f0", "--write-paf", "--write-ec", fa], capture_output=True, text=True, env=env)
    assert a.returncode == 0, a.stderr[-1500:]
    assert b.returncode == 0, b.stderr[-3000:]
    diff = tg._cmp_files(td, "ref", "gpu")
    assert not diff, diff
    # the reference's own log lines come out of the shim with the same numbers
    pick = lambda s: re.findall(r"\[M::pec::[0-9.]+\] # bases: (\d+); # corrected bases: (\d+)", s) + re.findall(r"\[M::pec::[0-9.]+\] # exact o: (\d+); # non-exact o: (\d+)", s) + [x for x in re.findall(r"\[M::ha_print_ovlp_stat_0\] # ([a-z ]+): (\d+)", s) if x[0] != "running time"]
    assert pick(a.stderr) == pick(b.stderr), (pick(a.stderr), pick(b.stderr))
    print("drop-in: %d reads, %d bases: files and log counters identical to the stock binary" % (rs.n, rs.bases))
