#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_round.py -m gpu -q -x > gpurun_out/r2_gpu_sub_19.log 2>&1; tail -3 gpurun_out/r2_gpu_sub_19.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench100_n1f.json 2> gpurun_out/r2_bench100_n1f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1f.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest']); print(list(d['roofline']['kernels_ms_per_step'].items())[:16])
PY
tail -2 gpurun_out/r2_bench100_n1f.err
HB_LANES=1 HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ecb_seg$ -c 1 -f -o gpurun_out/r2_prof_ecb_seg3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_seg3.log 2>&1
HB_LANES=1 HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ec_overlap$ -c 1 -f -o gpurun_out/r2_prof_ec_overlap python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_eco.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
