set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_ab30.txt
for v in base p17 p20 p24 p30 p40; do
  L=$PWD/sonics_b200/variants/$v.so; [ $v = base ] && L=$PWD/sonics_b200/libsonics_b200.so
  echo "== $v" >> gpurun_out/r2_ab30.txt
  SONICS_B200_LIB=$L timeout 200 python tools/k1_ab.py 4000000 cfg1,cfg2,cfg3 2>/dev/null | grep lockstep | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'], '%.4g'%d['reactions_per_s'], d['same_as_first'])" >> gpurun_out/r2_ab30.txt
  SONICS_B200_LIB=$L timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes | cut -c1-60 >> gpurun_out/r2_ab30.txt
done
cat gpurun_out/r2_ab30.txt
