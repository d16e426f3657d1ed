#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_bench100_n1.json 2> gpurun_out/r2_bench100_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench100_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value'],4),'e2e',round(d['e2e']['value'],4),'ms/step',round(d['ms_per_step']), d['config']['last_step_host_ms'], d['config']['result_digest'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value']); print(list(d['roofline']['kernels_ms_per_step'].items())[:18]); print(d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])
PY
tail -4 gpurun_out/r2_bench100_n1.err
# ncu: launch list of one small step (shares), then full captures of the kernels that top the step
HB_BENCH_GENOME_MB=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; wc -l gpurun_out/r2_launches.csv
HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ph_decide -c 1 -o gpurun_out/r2_prof_ph_decide python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_ph.log 2>&1; ls -la gpurun_out/*.ncu-rep
HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ecb_seg$ -c 1 -o gpurun_out/r2_prof_ecb_seg python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_seg.log 2>&1
HB_BENCH_GENOME_MB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -c 1 -o gpurun_out/r2_prof_expand python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_expand.log 2>&1
ls -la gpurun_out/*.ncu-rep
