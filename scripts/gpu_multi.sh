#!/bin/bash
# Usage (gpurun --gpus N): bash scripts/gpu_multi.sh <tag> <N>  -- the C++ multi-device caller on 1..N GPUs + the multi tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1; N=${2:-2}
DEV=$(seq -s, 0 $((N-1)))
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for d in 0 $(seq -s, 0 1) $(seq -s, 0 3) $DEV; do
  cnt=$(echo $d | tr ',' '\n' | wc -l)
  if [ $cnt -le $N ]; then tests/cpp/build/multi_gpu_b200 100000 200 28 $d | tee -a gpurun_out/${TAG}_multi_cpp.jsonl; fi
done
