set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2_t22_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t22_all.log
tail -5 gpurun_out/r2_t22_all.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg2.csv python tools/prof_k1.py cfg2 4000000 lockstep > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg1.csv python tools/prof_k1.py cfg1 4000000 lockstep > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_counts_cfg3.csv python tools/prof_k1.py cfg3 500000 lockstep > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcr_ls -c 1 -f -o gpurun_out/r2_prof_ls7 python tools/prof_k1.py cfg2 1000000 lockstep > gpurun_out/r2_ncu_ls7.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench22.json
