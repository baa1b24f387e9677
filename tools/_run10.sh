set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_stats.py tests/test_gpu_edges.py -q -m gpu -s -x > gpurun_out/r2_t8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8.log
grep -E "genotype [0-9]|passed|failed|FAILED|rc=|Error|assert" gpurun_out/r2_t8.log | tail -40
