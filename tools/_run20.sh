set -x
mkdir -p gpurun_out
timeout 200 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3 2>/dev/null | grep lockstep | cut -c1-200
SONICS_SORT_PAIR_ALWAYS=1 timeout 200 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3 2>/dev/null | grep lockstep | cut -c1-200
