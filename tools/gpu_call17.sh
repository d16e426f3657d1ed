#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 100663296" "4 100663296" "4 67108864" "3 150994944"; do
set -- $cfg; L=$1; B=$2
HB_LANES=$L HB_ANCHOR_BUDGET=$B timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_l${L}_b$B.json 2> gpurun_out/r2_bench100_l${L}_b$B.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench100_l${L}_b$B.json').read().strip().splitlines()[-1])
print('lanes $L budget $B value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest'])
PY
tail -1 gpurun_out/r2_bench100_l${L}_b$B.err
done
