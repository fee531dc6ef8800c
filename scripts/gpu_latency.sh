cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/latency_bench.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err; tail -3 gpurun_out/r02_latency.err; cat gpurun_out/r02_latency.json | head -80
