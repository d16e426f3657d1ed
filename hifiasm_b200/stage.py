"""The whole overlap / error-correction stage as one call: FASTA / FASTQ in, the reference's files out.

Host-side mirror of the stage part of ha_assemble() (Assembly.cpp:2076-2108): filter table, number_of_round x ha_ec (index of the
round + cal_ec_r), --write-ec, ha_ec_ff (final index + cal_ov_r), --write-paf, and the three .bin dumps a later hifiasm run reloads
instead of recomputing the stage.  Every step is a C-ABI call of include/hifiasm_b200.h (hb_readset_load, hb_reads_upload, hb_ft_gen,
hb_pt_gen, hb_cal_ec_r, hb_reads_download, hb_cal_ov_r, hb_write_*); nothing is computed here.
"""
from __future__ import annotations

import numpy as np

from . import binio
from .engine import Engine


def run_stage(paths, out_prefix: str, device: int = 0, n_round: int = 3, write_paf: bool = True, write_ec: bool = True, bf_shift: int = 0, adapter_len: int = 0):
    """-> dict(reads, bases, corrected_bases per round, overlaps).  Files: <out>.ec.bin, <out>.ovlp.source.bin, <out>.ovlp.reverse.bin and,
    on request, <out>.ec.fa (hifiasm --write-ec) and <out>.ovlp.paf (--write-paf).  bf_shift = hifiasm's -f (0 = exact counting).
    Under torchrun (torch.distributed initialised, one process per GPU) every pass is sharded over the ranks: reads + index replicated, the
    EC rounds through dist.cal_ec_r_sharded (all-gather of edit scripts and lists), the final pass on the rank's shard with the lists
    all-gathered; rank 0 writes the files."""
    from . import dist as hdist
    rank, world = hdist._world()
    tdev = None
    if world > 1:
        import torch
        tdev = torch.device("cuda", device)
    rs = binio.native_load_reads(paths, adapter_len)
    n = rs.n
    eng = Engine(device)
    try:
        eng.set_opt(bf_shift=bf_shift)
        eng.upload_store(rs)
        hom = eng.ft_gen(); eng.update_cov(hom)                                     # Assembly.cpp:2081-2085
        src = np.zeros(0, binio.MA_MEM); soff = np.zeros(n + 1, np.uint64); rev = src.copy(); roff = soff.copy()
        fc = np.zeros(n, np.uint8); ab = np.zeros(n, np.uint8); corrected = []
        hom_k = het_k = 0
        for k in range(n_round):                                                      # ha_ec, Assembly.cpp:996-1030
            hom_k, het_k = eng.pt_gen(); eng.set_opt(hom_cov=hom_k, het_cov=het_k)
            r = eng.cal_ec_r(k, 1 if k == n_round - 1 else 0, src, soff) if world == 1 else hdist.cal_ec_r_sharded(eng, k, 1 if k == n_round - 1 else 0, src, soff, device=tdev)
            if r["status"].any():
                raise RuntimeError("round %d: %d reads could not be finished on the device (status bits %s)" % (k, int((r["status"] != 0).sum()), sorted(set(int(x) for x in r["status"][r["status"] != 0]))))
            src, soff, rev, roff, fc, ab = r["src"], r["src_off"], r["rev"], r["rev_off"], r["is_fully_corrected"], r["is_abnormal"]
            corrected.append(r["tot_e"])
        reads = eng.download_reads()
        reads.names, reads.name_blob, reads.name_index = rs.names, rs.name_blob, rs.name_index
        reads.index_size, reads.name_index_size, reads.total_reads_bases, reads.adapter_len = rs.index_size, rs.name_index_size, rs.total_reads_bases, rs.adapter_len   # total_reads_bases keeps the pre-correction total (Process_Read.cpp:79)
        reads.trio_flag = np.zeros(n, np.uint8)                                       # AMBIGU after cal_ec_r (ecovlp.cpp:6301)
        if write_ec and rank == 0:
            binio.native_write_ec_fa(out_prefix + ".ec.fa", reads)                   # Assembly.cpp:2097-2100
        hom_f, het_f = eng.pt_gen(); eng.set_opt(hom_cov=hom_f, het_cov=het_f)       # ha_ec_ff(1), Assembly.cpp:1942-1959
        if world == 1:
            out0, oo0, out1, oo1, stat = eng.cal_ov_r(src, soff, rev, roff)
        else:
            r0, r1 = hdist.shard_range(n, rank, world)
            p0, q0, p1, q1, stat = eng.cal_ov_r(src, soff, rev, roff, r0, r1)
            out0, oo0 = hdist.all_gather_ragged(p0, q0, tdev); out1, oo1 = hdist.all_gather_ragged(p1, q1, tdev)
        reads.hom_cov, reads.het_cov = hom_f, het_f
        info = dict(reads=n, bases=int(rs.length.sum()), corrected_bases=corrected, overlaps_src=int(out0.size), overlaps_rev=int(out1.size), hom_cov=hom_f, het_cov=het_f)
        if rank != 0:
            return info
        if write_paf:
            binio.native_write_paf(out_prefix + ".ovlp.paf", reads, out0, oo0)       # Assembly.cpp:2109
        binio.native_write_ec_bin(out_prefix + ".ec.bin", reads)                     # write_all_data_to_disk, Overlaps.cpp:23567
        binio.native_write_ovlp_bin(out_prefix + ".ovlp.source.bin", out0, oo0, fc, ab)
        binio.native_write_ovlp_bin(out_prefix + ".ovlp.reverse.bin", out1, oo1, None, None)
        return info
    finally:
        eng.close()
