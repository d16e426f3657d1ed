#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_all_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_all_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'e2e',round(d['e2e']['value'],4),'ms/step',round(d['ms_per_step']), d['config']['step_device_ms'], d['config']['last_step_host_ms'], d['config']['result_digest'], d['clocks']); print(d['cpu_baseline']['value']); print(list(d['roofline']['kernels_ms_per_step'].items())[:12])
PY
tail -2 gpurun_out/r2_bench_n1.err
HB_LANES=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n1_lanes1.json 2> gpurun_out/r2_bench_n1_lanes1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_lanes1.json').read().strip().splitlines()[-1])
print('lanes1 value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms']); print(list(d['roofline']['kernels_ms_per_step'].items())[:24]); print(d['roofline'].get('byte_bound_kernel'))
PY
