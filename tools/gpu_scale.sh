#!/bin/bash
# Strong scaling on N GPUs of one box:  gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scale.sh N [check]'
#   the same 100 Mb stage under torchrun (1 warm-up + 1 timed step); with "check" first the 4 Mb stage on N ranks against the reference binary's files.
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" = "check" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/stage_vs_ref.py 4 > gpurun_out/r2_stage_vs_ref_n$N.log 2>&1; tail -2 gpurun_out/r2_stage_vs_ref_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"], 4), "Gbp/s", round(d["ms_per_step"]), "ms/step", d["config"]["last_step_host_ms"], d["config"]["result_digest"])
PY
