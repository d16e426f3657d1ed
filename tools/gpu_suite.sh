#!/bin/bash
# The GPU-box call the round's evidence comes from (one B200):  gpurun --timeout 3000 -- 'bash tools/gpu_suite.sh'
#   every -m gpu test, the default bench line (3 warm-up + 2 timed whole stages, cpu_baseline), the reference arm, and one step on one lane
#   (exclusive per-kernel times).  Copy gpurun_out/r2_* into profiles/ afterwards.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_all_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_all_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
HB_LANES=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n1_lanes1.json 2> gpurun_out/r2_bench_n1_lanes1.err
python - <<'PY'
import json
for f in ("r2_bench_n1", "r2_bench_reference_n1", "r2_bench_n1_lanes1"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "unavailable")}, (d.get("config") or {}).get("result_digest"))
PY
