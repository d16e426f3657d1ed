"""The filter table with the reference's Bloom filter (hifiasm -f 21 / 22 / 24; -f 37 is its default): tests/golden/bloom.npz holds what the
UNMODIFIED reference's ha_ft_gen + ha_ft_cnt answer for every distinct k-mer of a read set on which the filter changes the table
(tests/golden/make_bloom.py).  CPU: the oracle restatement (hao_ft_gen_bf) and the device function of the filter (hb_bf_insert, host
emulation).  GPU (test_gpu_bloom): hb_ft_gen with opt.bf_shift through the C-ABI."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ha_oracle as ho  # noqa: E402
from goldenlib import GOLDEN  # noqa: E402
from hifiasm_b200 import sim  # noqa: E402
import make_bloom  # noqa: E402


@pytest.fixture(scope="module")
def data():
    z = np.load(os.path.join(GOLDEN, "bloom.npz"))
    rr = make_bloom.reads()
    flat, boff, ln, npos, noff = sim.pack_reads(rr)
    return z, (ln, flat, boff, npos, noff)


@pytest.mark.parametrize("shift", make_bloom.SHIFTS)
def test_oracle_bloom_counting_vs_reference(data, shift):
    z, (ln, flat, boff, npos, noff) = data
    st = ho.Store(ln, boff, flat, noff, npos)
    opt = ho.default_opt()
    ft, hom = ho.ft_gen(st, opt, shift)
    key, cnt = z["f%d_key" % shift], z["f%d_cnt" % shift]
    assert hom == int(z["f%d_hom" % shift][0])
    assert int(ho.lib().hao_ft_size(C.c_void_p(ft))) == key.size
    mine = np.array([ho.lib().hao_ft_cnt(C.c_void_p(ft), C.c_uint64(int(h))) for h in key], dtype=np.int32)
    assert (mine == cnt).all()
    # the filter matters on this set: the tables of the small filters differ from the exact one, -f 24 by a single k-mer
    if shift:
        assert key.size > z["f0_key"].size


@pytest.mark.gpu
@pytest.mark.xfail(reason="hb_ft_gen's Bloom path (opt.bf_shift > 0) was written after the round's GPU budget was spent: not yet run on a B200", strict=False)
@pytest.mark.parametrize("shift", make_bloom.SHIFTS)
def test_gpu_bloom(data, shift):
    import hifiasm_b200
    z, (ln, flat, boff, npos, noff) = data
    eng = hifiasm_b200.Engine(0)
    eng.set_opt(bf_shift=shift)
    eng.upload_reads(ln, flat, boff, npos, noff)
    hom = eng.ft_gen()
    key, cnt = z["f%d_key" % shift], z["f%d_cnt" % shift]
    assert hom == int(z["f%d_hom" % shift][0]) and eng.ft_size() == key.size
    assert (eng.ft_cnt(key) == cnt).all()
    rng = np.random.default_rng(1)
    assert (eng.ft_cnt(rng.integers(0, 2**63, 1000, dtype=np.uint64)) == 0).all()
    eng.close()
