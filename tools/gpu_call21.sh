#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_round.py tests/test_gpu_dropin.py -m gpu -q -x -s > gpurun_out/r2_gpu_sub_21.log 2>&1; grep "scale parity" gpurun_out/r2_gpu_sub_21.log; tail -3 gpurun_out/r2_gpu_sub_21.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1g.json 2> gpurun_out/r2_bench100_n1g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1g.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:16])
PY
tail -2 gpurun_out/r2_bench100_n1g.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; cat gpurun_out/r2_bench_reference_n1.json | cut -c1-400
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
