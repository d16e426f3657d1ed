"""The commit order of concurrently processed batches (hifiasm_b200/csrc/hb_order.h, used by the pass's lanes): a plain C++ check on the CPU —
200 batches finished out of order by 1..4 threads come out in batch order; a failing batch releases every waiter."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_pass_order(tmp_path):
    exe = str(tmp_path / "order_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "cpp", "order_test.cpp"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == "OK", p.stdout + p.stderr
