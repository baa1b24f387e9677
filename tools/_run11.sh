set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -s --durations=25 > gpurun_out/r2_t11_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t11_all.log
timeout 600 python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc=$?"
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2_t11_all.log | tail -40
tail -c 3000 gpurun_out/r2_bench11.json
