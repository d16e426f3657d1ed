#!/bin/bash
# One ncu --set full capture of a kernel on a small whole stage (one lane, 4 Mb):  gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh k_ecb_seg$ r2_prof_ecb_seg'
#   read here with `ncu -i gpurun_out/NAME.ncu-rep --page raw --csv` and tools/ncu_by_line.py (per source line); never quote a time measured under ncu.
K=${1:?kernel regex}; O=${2:-prof}
mkdir -p gpurun_out
HB_LANES=1 HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/$O python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/$O.log 2>&1
ls -la gpurun_out/$O.ncu-rep
