#!/bin/bash
# round 2, GPU call 1: failure text of the two g4 stage tests, the whole GPU suite, whole-stage profiles with the reference beside them
mkdir -p gpurun_out
for t in test_stages_raw_g4 test_stages_final_g4; do
	timeout 300 python -m pytest tests/test_zz_gpu_not_yet_run.py -m gpu -q -x --runxfail -k "$t" > gpurun_out/r2_zz_$t.log 2>&1
	echo "== $t: $(tail -1 gpurun_out/r2_zz_$t.log)"
done
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/r2_gpu_all_0.log 2>&1; tail -3 gpurun_out/r2_gpu_all_0.log
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
timeout 600 python tools/profile_stage.py 20 --ref --out gpurun_out/r2_stage20_0.json > gpurun_out/r2_stage20_0.log 2>&1; tail -5 gpurun_out/r2_stage20_0.log
timeout 1500 python tools/profile_stage.py 100 --ref --out gpurun_out/r2_stage100_0.json > gpurun_out/r2_stage100_0.log 2>&1; tail -5 gpurun_out/r2_stage100_0.log
