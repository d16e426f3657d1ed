#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value',round(d['value'],4),'e2e',round(d['e2e']['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest'])
PY
grep "warm-up" gpurun_out/r2_bench_n8.err | head -3 | cut -c1-200; tail -2 gpurun_out/r2_bench_n8.err | cut -c1-300
