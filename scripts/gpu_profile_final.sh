#!/bin/bash
# Final-of-round profiling set (via gpurun): launch list of the default bench, full ncu captures of the headline
# kernel at the BASELINE size, of the permanent kernel and of the one-CTA-per-problem latency kernel.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=$1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:murty_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_murty python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_prof_murty.log 2>&1
tail -2 gpurun_out/${TAG}_prof_murty.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_perm python bench.py --steps 1 --warmup 1 --no-cpu --problems 2000 > gpurun_out/${TAG}_prof_perm.log 2>&1
tail -2 gpurun_out/${TAG}_prof_perm.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:murty_cta -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_cta python scripts/ncu_cta_case.py > gpurun_out/${TAG}_prof_cta.log 2>&1
tail -2 gpurun_out/${TAG}_prof_cta.log | cut -c1-200
ls -la gpurun_out | tail -8
