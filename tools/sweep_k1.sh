#!/bin/bash
# Compile k_pcr variants on the GPU box and time each with bench.py (3 steps).  usage: sweep_k1.sh "<-D flags>" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  (cd sonics_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wno-deprecated-declarations $flags -shared -o ../libsonics_b200.so sonics_b200.cu >/dev/null 2>&1) || { echo "BUILD FAILED: $flags"; continue; }
  python bench.py --no-cpu-baseline --no-genotyping --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-70s %.3e reactions/s  %.1f ms' % ('$flags', d['value'], d['ms_per_step']))"
done
