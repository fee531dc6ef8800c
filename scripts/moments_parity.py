"""Moments -> weights at scale: N gated frames through pda_association_from_moments_batch_host against the CPU chain
(oracle: cost matrix -> conditionCosts -> assignmentProb(k) -> un-compaction).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probabilisticsemslam_b200 import api, synth
from oracle.loader import load_oracle
N, K = int(os.environ.get("MOM_N", "4000")), 200
orc = load_oracle()
frames = synth.quadric_frames(N, first=700_000)
t0 = time.time(); got = api.association_from_moments_batch(frames, 10.0, K); t_gpu = time.time() - t0
costs = api.quadric_cost_batch(frames, 10.0)
worst = 0.0; bad_cost = 0
t0 = time.time()
for f, g, c in zip(frames, got, costs):
    want = orc.association_from_moments(*f, 10.0, K)
    d = np.abs(g - want) / np.maximum(np.abs(want), 1e-300)
    worst = max(worst, float(np.max(np.where(g == want, 0.0, d))))
    bad_cost += 0 if np.array_equal(c.view(np.int64), orc.quadric_cost_matrix(*f, 10.0).view(np.int64)) else 1
print(json.dumps({"frames": N, "k": K, "worst_relative_weight_difference": worst, "cost_matrices_not_bit_identical": bad_cost,
                  "gpu_call_seconds": round(t_gpu, 3), "cpu_seconds_one_thread": round(time.time() - t0, 2)}))
