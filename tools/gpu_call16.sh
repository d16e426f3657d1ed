#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_round.py -m gpu -q -x -s > gpurun_out/r2_gpu_sub_16.log 2>&1; grep "scale parity" gpurun_out/r2_gpu_sub_16.log; tail -3 gpurun_out/r2_gpu_sub_16.log
for L in 2 1; do
HB_LANES=$L timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_lanes$L.json 2> gpurun_out/r2_bench100_lanes$L.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench100_lanes$L.json').read().strip().splitlines()[-1])
print('lanes $L value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:14])
PY
tail -2 gpurun_out/r2_bench100_lanes$L.err
done
nvidia-smi --query-gpu=memory.used --format=csv
