#!/bin/bash
# compute-sanitizer passes over a small run of every kernel (via gpurun): memcheck + racecheck + synccheck.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import numpy as np
from probabilisticsemslam_b200 import api, synth
pb = synth.g1_dense(24, first=777)
for path in ("warp", "cta"):
    api.set_murty_path(path)
    r = api.murty_batch(pb, 40, weight_mode=api.WEIGHTS_GATED)
    print("murty", path, "nFound", r.n_found[:6], float(r.probs.sum()))
api.set_murty_path("auto")
fr = synth.quadric_frames(6, first=3)
print("moments->weights", [float(t.sum()) for t in api.association_from_moments_batch(fr, 10.0, 40)])
g2 = synth.g2_gated(6, first=5)
cond, _ = api.condition_costs_batch(g2)
keep = [p for p in range(len(cond)) if cond.matrix(p).shape[0] <= 14]
sub = synth.pack([cond.matrix(p) for p in keep], [int(cond.nL[p]) for p in keep])
tabs, st = api.permanent_prob_batch(sub, 1)
print("permprob", len(tabs), st.tolist())
A = synth.dense_square(3, 10)
print("perm", api.permanent_batch([a.reshape(10, 10, order="F") for a in A])[0])
print("range", api.permanent_range(A[0].reshape(10, 10, order="F"), 0, 512))
print("lap", api.assign2D(pb.matrix(0))[0])
print("approx", api.permanent_approx_batch([a.reshape(10, 10, order="F") for a in A], 40)[0])
print("permprob approx", len(api.permanent_prob_batch(sub, 0)[0]))
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  echo "=== $tool"
  PYTHONPATH=$PWD timeout 1500 compute-sanitizer --tool $tool --print-limit 30 python /tmp/san_case.py 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/sanitizer_$tool.txt
done
