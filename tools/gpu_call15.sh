#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_round.py -m gpu -q -x > gpurun_out/r2_gpu_sub_15.log 2>&1; tail -3 gpurun_out/r2_gpu_sub_15.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1e.json 2> gpurun_out/r2_bench100_n1e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1e.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:20])
PY
HB_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_bench100_trace.json 2> gpurun_out/r2_bench100_trace.err; python - <<'PY'
import re, collections
agg=collections.OrderedDict()
for l in open('gpurun_out/r2_bench100_trace.err'):
    m=re.match(r'\[hb trace\] (cal_ec_r )?(\S+)\s+\+?\s*([0-9.]+) ms', l)
    if m:
        k=(m.group(1) or '')+m.group(2); e=agg.setdefault(k,[0,0.0]); e[0]+=1; e[1]+=float(m.group(3))
for k,v in agg.items(): print('%-24s n=%5d  %10.1f ms'%(k,v[0],v[1]))
PY
grep -v "hb trace" gpurun_out/r2_bench100_trace.err | tail -3
