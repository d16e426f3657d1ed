# Sanity subset on the GPU box (2 minutes): round tests, the stage against the reference binary at 4 Mb, the drop-in binary, a 20 Mb stage digest.
#   gpurun --timeout 480 -- "bash tools/gpu_sanity.sh"
mkdir -p gpurun_out; timeout 400 python -m pytest tests/test_gpu_round.py "tests/test_gpu_scale.py::test_whole_stage_files_equal_reference_binary_at_scale[None-None]" tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2_gpu_last_sanity.log 2>&1; tail -3 gpurun_out/r2_gpu_last_sanity.log; HB_BENCH_GENOME_MB=20 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20 Mb', d['config']['result_digest'], round(d['ms_per_step']), [(k,v) for k,v in d['roofline']['kernels_ms_per_step'].items() if k=='k_ph_decide'])"
