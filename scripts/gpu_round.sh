#!/bin/bash
# Usage (on the GPU box, via gpurun): bash scripts/gpu_round.sh <tag> [tests] [bench] [launches] [murty] [perm]
# Runs the selected stages and leaves their outputs under gpurun_out/<tag>_*.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1; shift
for stage in "$@"; do
  case $stage in
    tests) timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json ;;
    benchq) timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1 ;;
    murty) ncu --set full --clock-control none --import-source on -k regex:murty_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_murty python bench.py --steps 2 --warmup 1 --no-cpu --problems 20000 > gpurun_out/${TAG}_prof_murty.log 2>&1; tail -2 gpurun_out/${TAG}_prof_murty.log | cut -c1-300 ;;
    perm) ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_perm python bench.py --steps 1 --warmup 1 --no-cpu --problems 2000 > gpurun_out/${TAG}_prof_perm.log 2>&1; tail -2 gpurun_out/${TAG}_prof_perm.log | cut -c1-300 ;;
  esac
done
