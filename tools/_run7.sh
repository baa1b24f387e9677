set -x
mkdir -p gpurun_out
SONICS_AB_BITS="5,5,5" timeout 300 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3,vcf_100,vcf_400 > gpurun_out/r2_ab5.jsonl 2> gpurun_out/r2_ab5.err
for v in mb5 mb7 mb8; do SONICS_B200_LIB=$PWD/sonics_b200/variants/$v.so SONICS_AB_BITS="5,5,5" timeout 300 python tools/k1_ab.py 4000000 cfg2,vcf_400 | grep lockstep | sed "s/^/$v /" >> gpurun_out/r2_ab5.jsonl; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcr_ls -c 1 -f -o gpurun_out/r2_prof_ls5 python tools/prof_k1.py cfg2 1000000 lockstep > gpurun_out/r2_ncu_ls5.log 2>&1
timeout 600 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_primitives.py -q -m gpu -x > gpurun_out/r2_t6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t6.log
cat gpurun_out/r2_ab5.jsonl; tail -n 4 gpurun_out/r2_t6.log
