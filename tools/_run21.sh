set -x
for mb in 3 4 6 7 9; do echo "morton $mb"; SONICS_SORT_MORTON=$mb SONICS_SORT_PAIR_ALWAYS=1 timeout 200 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3 2>/dev/null | grep lockstep | cut -c1-200;  SONICS_SORT_MORTON=$mb timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes | cut -c1-90; done
