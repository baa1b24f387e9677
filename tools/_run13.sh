set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_ls_stats.jsonl
for c in cfg2 cfg1 cfg3; do timeout 60 python tools/ls_stats.py run $c 100000 >> gpurun_out/r2_ls_stats.jsonl 2>> gpurun_out/r2_ls_stats.err; done
cat gpurun_out/r2_ls_stats.jsonl; tail -5 gpurun_out/r2_ls_stats.err
