set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fast_samplers.py -q -m gpu -s > gpurun_out/fast_tests.log 2>&1; echo "rc=$?" >> gpurun_out/fast_tests.log
grep -E "npq=|group |fast n=|passed|failed|rc=|Error|assert" gpurun_out/fast_tests.log | cut -c1-200 | tail -50
for t in 20 50 100 300 1000; do python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import k1_ab, torch
from sonics_b200 import engine as eng, driver
import numpy as np
e=eng.Engine(0); e.set_samplers("fast", $t)
pre={"one_allele_threshold": 45, "noise_threshold": 0.05}
for name,spec,kw,n in (("cfg1",k1_ab.CFG1,{},4000000),("cfg2",k1_ab.CFG2,{},4000000),("cfg3",k1_ab.CFG3,dict(n_cycles=30,capture_cycle=16),500000)):
    kind,(reads,nc)=driver.prefilter(("s","b",".",spec),pre,{"monte_carlo":True})
    p=eng.make_params(**kw); t=e.build_table_from_arrays(p,[reads],[float(np.float32(nc))])
    ms,_=k1_ab.timed(e,p,t,n)
    print("fast min_npq=$t", name, "%.4g"%(n/ms*1e3))
PY
done
SONICS_FAST_SAMPLERS=100 timeout 200 python tools/genotyping_probe.py 2>&1 | grep genotypes | cut -c1-200
