"""The whole overlap / error-correction stage as one call: FASTA / FASTQ in, the reference's files out.

Host-side mirror of the stage part of ha_assemble() (Assembly.cpp:2076-2108): filter table, number_of_round x ha_ec (index of the
round + cal_ec_r), --write-ec, ha_ec_ff (final index + cal_ov_r), --write-paf, and the three .bin dumps a later hifiasm run reloads
instead of recomputing the stage.  Every step is a C-ABI call of include/hifiasm_b200.h (hb_readset_load, hb_reads_upload, hb_stage_run =
hb_ft_gen + n x (hb_pt_gen + hb_cal_ec_r) + hb_pt_gen + hb_cal_ov_r, hb_reads_download, hb_write_*); nothing is computed here.
"""
from __future__ import annotations

import numpy as np

from . import binio
from .engine import Engine


def run_stage(paths, out_prefix: str, device: int = 0, n_round: int = 3, write_paf: bool = True, write_ec: bool = True, bf_shift: int = 0, adapter_len: int = 0):
    """-> dict(reads, bases, corrected_bases per round, overlaps).  Files: <out>.ec.bin, <out>.ovlp.source.bin, <out>.ovlp.reverse.bin and,
    on request, <out>.ec.fa (hifiasm --write-ec) and <out>.ovlp.paf (--write-paf).  bf_shift = hifiasm's -f (0 = exact counting).
    bf_shift defaults to 0, NOT to hifiasm's own default of 37: the files equal those of `hifiasm -f0`; pass bf_shift=37 for the files of a default run.
    Under torchrun (torch.distributed initialised, one process per GPU) hb_stage_run shards every pass over the ranks: reads + index replicated, one
    all-gather of edit scripts + lists + flags per EC round and one of the final lists (csrc/stage.cu; a rank that fails says so in its blob, so the
    others raise instead of waiting); rank 0 writes the files."""
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
        # the stage itself is ONE C-ABI call (hb_stage_run, csrc/stage.cu): filter table, n_round x (index + cal_ec_r), index + cal_ov_r; with more than one
        # rank it shards every pass and exchanges edit scripts + lists through the all-gather callback (NCCL via torch.distributed)
        from .engine import torch_allgather
        r = eng.stage_run(n_round, rank, world, torch_allgather(tdev) if world > 1 else None)
        if r["n_unfinished"]:
            raise RuntimeError("%d reads could not be finished on the device (consensus arena / rescue scratch)" % r["n_unfinished"])
        out0, oo0, out1, oo1, fc, ab, corrected = r["src"], r["src_off"], r["rev"], r["rev_off"], r["is_fully_corrected"], r["is_abnormal"], r["corrected_bases"]
        hom_f, het_f = r["hom_cov"], r["het_cov"]
        reads = eng.download_reads()                                                  # the corrected reads (the final pass does not change them)
        reads.names, reads.name_blob, reads.name_index = rs.names, rs.name_blob, rs.name_index
        reads.index_size, reads.name_index_size, reads.total_reads_bases, reads.adapter_len = rs.index_size, rs.name_index_size, rs.total_reads_bases, rs.adapter_len   # total_reads_bases keeps the pre-correction total (Process_Read.cpp:79)
        reads.trio_flag = np.zeros(n, np.uint8)                                       # AMBIGU after cal_ec_r (ecovlp.cpp:6301)
        if write_ec and rank == 0:
            binio.native_write_ec_fa(out_prefix + ".ec.fa", reads)                   # Assembly.cpp:2097-2100
        reads.hom_cov, reads.het_cov = hom_f, het_f
        info = dict(reads=n, bases=int(rs.length.sum()), corrected_bases=corrected, overlaps_src=int(out0.size), overlaps_rev=int(out1.size), hom_cov=hom_f, het_cov=het_f, ms=r["ms"], device_ms=r["device_ms"])
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
