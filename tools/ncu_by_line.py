#!/usr/bin/env python
"""Stall samples / executed instructions / active lanes of ONE kernel of an ncu report, summed per source line.
ncu's CSV source page is SASS-only; the line of every SASS instruction comes from `nvdisasm -g` of the cubin inside the object file.
    python tools/ncu_by_line.py gpurun_out/X.ncu-rep hifiasm_b200/csrc/build/kern_ecb.o _Z9k_ecb_seg9EcCigArgs [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    amap, cur, on = {}, None, False
    for l in dis.splitlines():
        if l.startswith(".text."):
            on = l.startswith(".text." + mangled + ":")
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            amap[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isamp, iex, ith = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    agg, tot, base = collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0], None
    for r in rows[2:]:
        try:
            a = int(r[ia], 16)
        except (ValueError, IndexError):
            continue
        base = a if base is None else base
        v = [int(float(r[i] or 0)) for i in (isamp, iex, ith)]
        k = amap.get(a - base)
        for i in range(3):
            agg[k][i] += v[i]
            tot[i] += v[i]
    print("%d SASS instructions mapped; samples %d, warp instructions %d, lanes per instruction %.1f" % (len(amap), tot[0], tot[1], tot[2] / max(1, tot[1])))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-26s samples %5.1f %%   instructions %5.1f %%   lanes %4.1f" % ("%s:%d" % k if k else "?", 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(1, v[1])))


if __name__ == "__main__":
    main()
