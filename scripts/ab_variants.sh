#!/bin/bash
# runs scripts/fast_path_probe.py once per build_dbg/libpda_*.so (and once for the in-tree library)
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-100000}; K=${2:-200}
echo "== in-tree"; PROBE_PATHS=warp,fast python scripts/fast_path_probe.py $N $K | grep path
for f in build_dbg/libpda_*.so; do echo "== $f"; PDA_B200_LIB=$PWD/$f PROBE_PATHS=warp,fast python scripts/fast_path_probe.py $N $K | grep path; done
