#!/usr/bin/env python
"""GPU box: chains of golden set g4 from the device (k_chain_warp + k_post_warp) against the host emulation of the per-thread bodies (which
match the reference's digests), field by field.  Prints the first differing chains per read.   usage: python tools/debug_g4_chains.py [g4] [raw|final]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hifiasm_b200  # noqa: E402
from goldenlib import Golden  # noqa: E402
import emu  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "g4"; mode = sys.argv[2] if len(sys.argv) > 2 else "raw"
    g = Golden(name); rs = g.raw if mode == "raw" else g.pre; p = g.params(mode)
    eng = hifiasm_b200.Engine(0)
    eng.upload_store(g.raw); hom = eng.ft_gen(); eng.update_cov(hom)
    eng.upload_store(rs); hom, het = eng.pt_gen(); eng.set_opt(hom_cov=hom, het_cov=het)
    n = rs.n; bw = float(p["bw_thres"])
    aoff, an = eng.anchors(0, n)
    coff, ch, hoff, hits, foff, fc = eng.chains(0, n, bw)
    R = emu.Reads(rs); k = int(eng.get_opt().k_mer_length); mnc = int(eng.get_opt().max_n_chain)
    nbad = 0
    for i in range(n):
        e_ch, e_hits, e_fc, _ = emu.chains(R, i, an[int(aoff[i]):int(aoff[i + 1])], bw, k, mnc)
        d_ch = ch[int(coff[i]):int(coff[i + 1])]; d_fc = fc[int(foff[i]):int(foff[i + 1])]; d_hits = hits[int(hoff[i]):int(hoff[i + 1])]
        bad = []
        if e_ch.size != d_ch.size:
            bad.append("chain count %d (device) vs %d (emu)" % (d_ch.size, e_ch.size))
        for j in range(min(e_ch.size, d_ch.size)):
            for f in ("x_pos_s", "x_pos_e", "y_id", "y_pos_s", "y_pos_e", "y_pos_strand", "shared_seed", "first_hit", "n_hits", "fc_n"):
                if int(e_ch[j][f]) != int(d_ch[j][f]):
                    bad.append("chain %d field %s: device %d emu %d  (device %s | emu %s)" % (j, f, int(d_ch[j][f]), int(e_ch[j][f]), d_ch[j], e_ch[j]))
                    break
            else:
                a = d_fc[int(d_ch[j]["fc_off"]):int(d_ch[j]["fc_off"]) + int(d_ch[j]["fc_n"])]; b = e_fc[int(e_ch[j]["fc_off"]):int(e_ch[j]["fc_off"]) + int(e_ch[j]["fc_n"])]
                if a.size != b.size or (a != b).any():
                    bad.append("chain %d fake cigar differs: device %s emu %s" % (j, a[:6], b[:6]))
        if d_hits.size != e_hits.size or (d_hits.tobytes() != e_hits.tobytes()):
            bad.append("chain anchors differ (%d vs %d)" % (d_hits.size, e_hits.size))
        if bad and nbad < 3:   # every chain slot (kept or not) from the compacted anchor lists: (ordinal -> first anchor's target-side offset, anchors)
            def slots(h):
                o = h["id_strand"] & 0x7fffffff; out = []
                for k in np.unique(o):
                    m = h[o == k]; out.append((int(k), int(m.size), int(m["self_offset"][0]), int(m["offset"][0]), int(m["self_offset"][-1]), int(m["offset"][-1]), int(m["id_strand"][0] >> 31)))
                return out
            sd, se = slots(d_hits), slots(e_hits)
            print("   slots: device %d, emu %d" % (len(sd), len(se)))
            ke = {(x[2], x[3], x[6]): x for x in se}; kd = {(x[2], x[3], x[6]): x for x in sd}
            for x in sd:
                y = ke.get((x[2], x[3], x[6]))
                if y is None or y[1] != x[1] or y[4:6] != x[4:6]:
                    print("   device slot %s  <-> emu %s" % (x, y))
            for y in se:
                if (y[2], y[3], y[6]) not in kd:
                    print("   emu-only slot %s" % (y,))
        if bad:
            nbad += 1
            if nbad <= 6:
                print("read %d (%d anchors, %d chains):" % (i, int(aoff[i + 1] - aoff[i]), d_ch.size))
                for b in bad[:5]:
                    print("   ", b)
    print("%s/%s: %d of %d reads differ; profile %s; counters %s" % (name, mode, nbad, n, {k: v[0] for k, v in eng.profile().items()}, eng.counters()))
    eng.close()


if __name__ == "__main__":
    main()
