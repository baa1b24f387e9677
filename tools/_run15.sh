set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_ls_stats_vcf.jsonl
for m in pair slot; do timeout 120 python tools/ls_stats.py run vcf:10000:$m 100 >> gpurun_out/r2_ls_stats_vcf.jsonl 2>> gpurun_out/r2_ls_stats.err; done
for m in pair slot; do timeout 120 python tools/ls_stats.py run vcf:10000:$m 400 >> gpurun_out/r2_ls_stats_vcf.jsonl 2>> gpurun_out/r2_ls_stats.err; done
SONICS_AB_LOCI=10000 timeout 300 python tools/k1_ab.py 4000000 vcf_100,vcf_400 > gpurun_out/r2_ab15.jsonl 2> gpurun_out/r2_ab15.err
cat gpurun_out/r2_ab15.jsonl; tail -3 gpurun_out/r2_ab15.err gpurun_out/r2_ls_stats.err
