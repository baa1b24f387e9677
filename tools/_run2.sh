set -x
mkdir -p gpurun_out
timeout 500 python tools/_diag_ls.py > gpurun_out/r2_diag2.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2_t2_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2_all.log
cat gpurun_out/r2_diag2.log | grep -v "^   g="
tail -n 40 gpurun_out/r2_t2_all.log
