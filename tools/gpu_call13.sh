#!/bin/bash
mkdir -p gpurun_out
for B in 150994944 201326592; do
HB_ANCHOR_BUDGET=$B timeout 900 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_bench100_b$B.json 2> gpurun_out/r2_bench100_b$B.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench100_b$B.json').read().strip().splitlines()[-1])
    print('budget $B value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:12])
except Exception as e: print('budget $B failed', e)
PY
tail -2 gpurun_out/r2_bench100_b$B.err; nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
timeout 600 python -m pytest tests/test_gpu_round.py -m gpu -q -x -k stale > gpurun_out/r2_gpu_stale.log 2>&1; tail -2 gpurun_out/r2_gpu_stale.log
