set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2_t7_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_all.log
grep -E "worst|KS|passed|failed|FAILED|rc=" gpurun_out/r2_t7_all.log | tail -120
