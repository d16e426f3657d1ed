#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stage_vs_ref.py 4 > gpurun_out/r2_stage_vs_ref_n1.log 2>&1; tail -3 gpurun_out/r2_stage_vs_ref_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/stage_vs_ref.py 4 > gpurun_out/r2_stage_vs_ref_n2.log 2>&1; tail -3 gpurun_out/r2_stage_vs_ref_n2.log
HB_BENCH_GENOME_MB=20 timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/r2_bench20_n1.json 2> gpurun_out/r2_bench20_n1.err; tail -c 1500 gpurun_out/r2_bench20_n1.json; tail -3 gpurun_out/r2_bench20_n1.err
HB_BENCH_GENOME_MB=20 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_bench20_n2.json 2> gpurun_out/r2_bench20_n2.err; tail -c 600 gpurun_out/r2_bench20_n2.json; tail -3 gpurun_out/r2_bench20_n2.err
