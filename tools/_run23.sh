set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_edges.py -q -m gpu -x > gpurun_out/r2_t23.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t23.log
tail -3 gpurun_out/r2_t23.log
timeout 200 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3 2>/dev/null | grep lockstep | cut -c1-200
timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes | cut -c1-90
timeout 200 python tools/genotyping_probe.py 2500 32 0.1 2>&1 | grep genotypes | cut -c1-90
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg2.csv python tools/prof_k1.py cfg2 4000000 lockstep > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg1.csv python tools/prof_k1.py cfg1 4000000 lockstep > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg3.csv python tools/prof_k1.py cfg3 500000 lockstep > /dev/null 2>&1
