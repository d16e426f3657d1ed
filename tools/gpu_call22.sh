#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/stage_vs_ref.py 4 > gpurun_out/r2_stage_vs_ref_n2.log 2>&1; tail -3 gpurun_out/r2_stage_vs_ref_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value'],4),'e2e',round(d['e2e']['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest'])
PY
tail -3 gpurun_out/r2_bench_n2.err | cut -c1-300
