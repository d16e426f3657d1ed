#!/bin/bash
mkdir -p gpurun_out
HB_LANES=1 HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ecb_seg$ -c 1 -f -o gpurun_out/r2_prof_ecb_seg2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_seg2.log 2>&1
HB_LANES=1 HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_windows -c 1 -f -o gpurun_out/r2_prof_windows2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_win2.log 2>&1
ls -la gpurun_out/*.ncu-rep
