#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_g4_chains.py g4 raw > gpurun_out/r2_g4_chains_raw.log 2>&1; tail -40 gpurun_out/r2_g4_chains_raw.log
HB_CHAIN_THREAD=1 HB_POST_THREAD=1 timeout 300 python tools/debug_g4_chains.py g4 raw > gpurun_out/r2_g4_chains_raw_thread.log 2>&1; tail -4 gpurun_out/r2_g4_chains_raw_thread.log
timeout 1200 python -m pytest tests -m gpu -q -k "not g4 or whole" > gpurun_out/r2_gpu_all_3.log 2>&1; tail -5 gpurun_out/r2_gpu_all_3.log
timeout 600 python tools/profile_stage.py 20 --out gpurun_out/r2_stage20_3.json > gpurun_out/r2_stage20_3.log 2>&1; tail -3 gpurun_out/r2_stage20_3.log
