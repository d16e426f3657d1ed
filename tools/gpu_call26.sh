#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_round.py -m gpu -q -x > gpurun_out/r2_gpu_sub_26.log 2>&1; tail -3 gpurun_out/r2_gpu_sub_26.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1i.json 2> gpurun_out/r2_bench100_n1i.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1i.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:12])
PY
HB_FT_CHUNK_BITS=2 HB_BENCH_GENOME_MB=20 timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20 Mb chunked ft', d['config']['result_digest'], d['config']['last_step_host_ms']['ft'])"
HB_BENCH_GENOME_MB=20 timeout 600 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20 Mb one-pass ft', d['config']['result_digest'], d['config']['last_step_host_ms']['ft'])"
