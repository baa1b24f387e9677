#!/bin/bash
# One measurement round on a B200 box (run under gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh tests counters profile bench'
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_round.sh scale 8'
# Everything lands in gpurun_out/; copy what should be kept into profiles/ (tools/make_k1_counters.py,
# tools/ncu_summary.py, tools/ncu_regions.py turn the ncu files into the committed summaries).
set -x
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
while [ $# -gt 0 ]; do
  case "$1" in
    tests)     # the -m gpu suite with the printed p-values / agreement rates (profiles/rNN_gates_report.txt)
      timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" ;;
    counters)  # ncu metric pass over every kernel of one K1 step per configuration -> profiles/k1_counters.json
      for c in cfg2:4000000 cfg1:4000000 cfg3:500000; do
        timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/counts_${c%%:*}.csv \
          python tools/prof_k1.py ${c%%:*} ${c##*:} lockstep > /dev/null 2>&1
      done ;;
    profile)   # one full capture of the dominant kernel (after the program ran clean without ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcr_ls -c 1 -f \
        -o gpurun_out/prof_k_pcr_ls python tools/prof_k1.py cfg2 1000000 lockstep > gpurun_out/ncu_profile.log 2>&1 ;;
    launches)  # launch list of the genotyping loop (cold-cache, serialised: shares of the step, not absolutes)
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_genotyping.csv python tools/genotyping_probe.py > /dev/null 2>&1 ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" ;;
    scale)     # bench.py on N GPUs of this box, as the driver launches it
      N=$2; shift
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
      echo "scale rc=$?" ;;
  esac
  shift
done
