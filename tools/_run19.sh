set -x
mkdir -p gpurun_out
for mag in "0,0" "2,7" "3,6" "4,5"; do echo "mag $mag"; SONICS_SORT_MAG=$mag timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes | cut -c1-90;  SONICS_SORT_MAG=$mag timeout 200 python tools/genotyping_probe.py 2500 32 0.1 2>&1 | grep genotypes | cut -c1-90; done
