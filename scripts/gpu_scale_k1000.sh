#!/bin/bash
# BASELINE.json configs[2]: the SAME 100 000-problem batch at k = 1000, sharded over N GPUs (strong scaling).
# Usage (via gpurun --gpus N): bash scripts/gpu_scale_k1000.sh <tag> <N>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1; N=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --k 1000 --problems $((100000 / N)) --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_k1000_n$N.json | cut -c1-300
