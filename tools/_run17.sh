set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lockstep.py -q -m gpu -x > gpurun_out/r2_t17.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t17.log
tail -3 gpurun_out/r2_t17.log
rm -f gpurun_out/r2_ls_stats_vcf.jsonl gpurun_out/r2_ls_stats.err
for n in 100 400; do timeout 120 python tools/ls_stats.py run vcf:10000:pair $n >> gpurun_out/r2_ls_stats_vcf.jsonl 2>> gpurun_out/r2_ls_stats.err; done
SONICS_AB_LOCI=10000 timeout 300 python tools/k1_ab.py 4000000 vcf_100,vcf_400 > gpurun_out/r2_ab17.jsonl 2> gpurun_out/r2_ab17.err
cut -c1-250 gpurun_out/r2_ab17.jsonl
for m in pair slot; do SONICS_SORT_MODE=$m timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes; done
timeout 200 python tools/genotyping_probe.py 2500 32 0.1 2>&1 | grep genotypes
SONICS_SORT_MODE=slot timeout 200 python tools/genotyping_probe.py 2500 32 0.1 2>&1 | grep genotypes
