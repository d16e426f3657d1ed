#!/bin/bash
# First GPU call of a round (run under gpurun from the repo root):
#   1. the device paths written after round 1's GPU budget was spent (tests/test_zz_gpu_not_yet_run.py), each in its own process with --runxfail,
#      so a failure shows its traceback and a crash cannot poison the next one
#   2. the whole GPU suite
#   3. end-to-end differential fuzzing against the reference binary on the box (tools/fuzz_gpu_stage.py)
#   4. one default bench run (JSON line -> gpurun_out/bench_n1.json, stderr -> gpurun_out/bench_n1.err)
# Usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# then, on 2 GPUs:  /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ec_sharded_check.py g1'
mkdir -p gpurun_out
export HB_TRACE_EC=1
for t in test_whole_stage_from_raw_reads_g4 test_stages_raw_g4 test_stages_final_g4 test_gpu_bloom test_run_stage_fasta_to_files test_sharded_round_single_rank_equals_cal_ec_r; do
	timeout 300 python -m pytest tests/test_zz_gpu_not_yet_run.py -m gpu -q -x --runxfail -k "$t" > gpurun_out/zz_$t.log 2>&1
	echo "== $t: $(tail -1 gpurun_out/zz_$t.log)"
done
timeout 900 python tools/fuzz_gpu_stage.py 5000 11 > gpurun_out/fuzz_gpu_stage.log 2>&1; tail -4 gpurun_out/fuzz_gpu_stage.log
unset HB_TRACE_EC
timeout 600 python -m pytest tests -m gpu -q -rxX > gpurun_out/gpu_all.log 2>&1; tail -3 gpurun_out/gpu_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
