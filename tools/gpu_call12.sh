#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_all_9.log 2>&1; tail -4 gpurun_out/r2_gpu_all_9.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1c.json 2> gpurun_out/r2_bench100_n1c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1c.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:20])
PY
tail -2 gpurun_out/r2_bench100_n1c.err
HB_BENCH_GENOME_MB=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_windows -c 1 -o gpurun_out/r2_prof_windows python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_win.log 2>&1
HB_BENCH_GENOME_MB=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_post_warp -c 1 -o gpurun_out/r2_prof_post python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_post.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
