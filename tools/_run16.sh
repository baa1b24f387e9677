set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_ls_stats_vcf.jsonl gpurun_out/r2_ls_stats.err
for m in pair slot; do timeout 120 python tools/ls_stats.py run vcf:10000:$m 100 >> gpurun_out/r2_ls_stats_vcf.jsonl 2>> gpurun_out/r2_ls_stats.err; done
for m in pair slot; do SONICS_SORT_MODE=$m timeout 200 python tools/genotyping_probe.py > gpurun_out/r2_geno16_$m.txt 2>&1; done
for m in pair slot; do SONICS_SORT_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_geno16_launches_$m.csv python tools/genotyping_probe.py > /dev/null 2>&1; done
cat gpurun_out/r2_geno16_*.txt
