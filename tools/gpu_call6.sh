#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_all_5.log 2>&1; tail -5 gpurun_out/r2_gpu_all_5.log
timeout 900 python tools/profile_stage.py 100 --out gpurun_out/r2_stage100_5.json > gpurun_out/r2_stage100_5.log 2>&1; tail -3 gpurun_out/r2_stage100_5.log
