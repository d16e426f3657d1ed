#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_g4_chains.py g4 raw > gpurun_out/r2_g4_chains_raw.log 2>&1; tail -25 gpurun_out/r2_g4_chains_raw.log
timeout 1200 python -m pytest tests -m gpu -q -x -k "not g4 or whole" > gpurun_out/r2_gpu_all_2.log 2>&1; tail -5 gpurun_out/r2_gpu_all_2.log
HB_TRACE_EC=1 timeout 600 python tools/profile_stage.py 20 --out gpurun_out/r2_stage20_2.json > gpurun_out/r2_stage20_2.log 2>&1; tail -3 gpurun_out/r2_stage20_2.log; grep "EC base" gpurun_out/r2_stage20_2.log | head -5
