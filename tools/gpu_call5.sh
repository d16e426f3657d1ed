#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_g4_chains.py g4 raw > gpurun_out/r2_g4_chains_raw.log 2>&1; tail -4 gpurun_out/r2_g4_chains_raw.log
timeout 1200 python -m pytest tests -m gpu -q --runxfail > gpurun_out/r2_gpu_all_4.log 2>&1; tail -5 gpurun_out/r2_gpu_all_4.log
timeout 600 python tools/profile_stage.py 20 --out gpurun_out/r2_stage20_4.json > gpurun_out/r2_stage20_4.log 2>&1; tail -3 gpurun_out/r2_stage20_4.log
timeout 900 python tools/profile_stage.py 100 --out gpurun_out/r2_stage100_4.json > gpurun_out/r2_stage100_4.log 2>&1; tail -3 gpurun_out/r2_stage100_4.log
