#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1
nproc; lscpu | grep "Model name"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_n1.json
python scripts/perm_bench.py | tee gpurun_out/${TAG}_perm_bench.txt
