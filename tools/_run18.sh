set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcr_ls -c 1 -f -o gpurun_out/r2_prof_ls6 python tools/prof_k1.py cfg2 1000000 lockstep > gpurun_out/r2_ncu_ls6.log 2>&1
ls -la gpurun_out/
