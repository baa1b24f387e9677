set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py -q -m gpu -s -x > gpurun_out/r2_t12.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t12.log
for c in cfg2 cfg1 cfg3; do timeout 300 python tools/ls_stats.py run $c 200000 >> gpurun_out/r2_ls_stats.jsonl 2>> gpurun_out/r2_ls_stats.err; done
grep -E "run_reps|genotype [0-9]|passed|failed|FAILED|rc=|Error|assert" gpurun_out/r2_t12.log | tail -30
cat gpurun_out/r2_ls_stats.jsonl; tail -5 gpurun_out/r2_ls_stats.err
