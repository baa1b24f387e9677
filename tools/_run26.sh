set -x
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo "rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_final_n$N.json 2> gpurun_out/r2_bench_final_n$N.err; echo "rc=$?"
fi
tail -c 900 gpurun_out/r2_bench_final_n$N.json
