set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lockstep.py -q -m gpu -x > gpurun_out/r2_t14.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t14.log
tail -5 gpurun_out/r2_t14.log
timeout 300 python tools/k1_ab.py 4000000 cfg2,vcf_100,vcf_400 > gpurun_out/r2_ab14.jsonl 2> gpurun_out/r2_ab14.err
cat gpurun_out/r2_ab14.jsonl; tail -3 gpurun_out/r2_ab14.err
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench14.json'))
print(d['value'], d['e2e']['value'])
g=d['genotyping']; print({k:g[k] for k in ('value','seconds','reactions_per_s')}, g['file_to_file'], g['strong_scaling']['value'], g['strong_scaling']['seconds'], g['strong_scaling']['timings_rank0'])
PY
