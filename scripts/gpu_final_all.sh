#!/bin/bash
# One gpurun call for the end-of-round evidence: profiles, bench lines (both arms), latency table, configs report,
# sanitizers.  Usage: bash scripts/gpu_final_all.sh <tag>
cd "${GRAFT_REPO_ROOT:-.}"
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_profile_final.sh $TAG
bash scripts/gpu_final_bench.sh $TAG
timeout 300 python scripts/latency_bench.py > gpurun_out/${TAG}_latency.json 2> gpurun_out/${TAG}_latency.err
PDA_B200_LIB=$PWD/probabilisticsemslam_b200/libpda_b200_prof.so python scripts/cta_phase_profile.py > gpurun_out/${TAG}_cta_phases.txt 2>&1
CFG3_PROBLEMS=100000 timeout 900 python scripts/configs_report.py > gpurun_out/${TAG}_configs_1_3_4.json 2> gpurun_out/${TAG}_configs_err.log
bash scripts/sanitize.sh > gpurun_out/${TAG}_sanitize.log 2>&1
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_*.txt
