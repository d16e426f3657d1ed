#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_g4_chains.py g4 raw > gpurun_out/r2_g4_chains_raw.log 2>&1; tail -25 gpurun_out/r2_g4_chains_raw.log
timeout 900 python -m pytest tests/test_gpu_round.py -m gpu -q -x > gpurun_out/r2_gpu_round_1.log 2>&1; tail -5 gpurun_out/r2_gpu_round_1.log
timeout 600 python tools/profile_stage.py 20 --out gpurun_out/r2_stage20_1.json > gpurun_out/r2_stage20_1.log 2>&1; tail -3 gpurun_out/r2_stage20_1.log
HB_TRACE_WS=1 timeout 900 python tools/profile_stage.py 100 --out gpurun_out/r2_stage100_1.json > gpurun_out/r2_stage100_1.log 2>&1; grep "hb ws" gpurun_out/r2_stage100_1.log | sort | uniq -c | sort -k4 -n -r | head -30
tail -3 gpurun_out/r2_stage100_1.log
