#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_all_10.log 2>&1; tail -4 gpurun_out/r2_gpu_all_9.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1d.json 2> gpurun_out/r2_bench100_n1d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1d.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:20])
PY
tail -2 gpurun_out/r2_bench100_n1d.err
