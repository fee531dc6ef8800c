#!/bin/bash
# Usage (via gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1; N=$2
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/${TAG}_scale_n$N.json
