"""The filter table with the reference's Bloom filter (hifiasm -f 21 / 22 / 24; -f 37 is its default): tests/golden/bloom.npz holds what the
UNMODIFIED reference's ha_ft_gen + ha_ft_cnt answer for every distinct k-mer of a read set on which the filter changes the table
(tests/golden/make_bloom.py).  CPU: the oracle restatement (hao_ft_gen_bf) and the device function of the filter (hb_bf_insert, host
emulation).  GPU: tests/test_zz_gpu_not_yet_run.py::test_gpu_bloom, hb_ft_gen with opt.bf_shift through the C-ABI."""
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


@pytest.mark.parametrize("shift", make_bloom.SHIFTS)
def test_device_formulation_vs_reference(data, shift):
    """hb_bloom.cuh + the way index.cu evaluates the filter (distinct k-mers grouped by filter block, walked in first-occurrence order)
    instead of streaming every occurrence through 4096 filters: same table as the reference"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu"))
    import emu
    z, (ln, flat, boff, npos, noff) = data
    st = ho.Store(ln, boff, flat, noff, npos)
    key, cnt = emu.bf_counts(ho.all_kmers(st, ho.default_opt()), shift)
    hom = int(z["f%d_hom" % shift][0]); cutoff = int(hom * 5.0)   # ha_ft_gen: cutoff = peak_hom * high_factor, table = counts in [cutoff, 4095]
    rk, rc = z["f%d_key" % shift], z["f%d_cnt" % shift]
    keep = cnt >= cutoff
    assert int(keep.sum()) == rk.size and (key[keep] == rk).all()
    c = cnt[keep].astype(np.int64); c[c > 2000] = 2**31 - 1   # gen_hh: counts above max_kmer_cnt read as INT32_MAX (htab.cpp:1038-1070)
    assert (c == rc.astype(np.int64)).all()
