cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PDA_B200_LIB=$PWD/probabilisticsemslam_b200/libpda_b200_prof.so python scripts/cta_phase_profile.py 2>&1 | grep -E "EVENTS|ns"
timeout 300 python scripts/latency_bench.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err; tail -3 gpurun_out/r02_latency.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_latency.json"))
print({k: round(v) for k, v in d["host_call_us"].items()}); print({k: round(v) for k, v in d["cpu_us"].items()})
print({k: round(v) for k, v in d["kernel_us"].items()})
PY
