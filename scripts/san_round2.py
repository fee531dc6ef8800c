"""compute-sanitizer case for the kernels that are new in round 2: the pruning Murty kernel (bound tightening included) and its
exact fallback on tied problems, the fused finalisation of the NW permanent (several CTAs per matrix), the subset-DP
permanent.  usage: compute-sanitizer --tool racecheck python scripts/san_round2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probabilisticsemslam_b200 import api, synth
api.set_murty_path("fast")
pb = synth.g1_dense(8, first=777)
r = api.murty_batch(pb, 120, weight_mode=api.WEIGHTS_GATED)
print("murty fast nFound", r.n_found.tolist(), float(r.probs.sum()))
pi = synth.g1_dense(4, first=5, integer=True)
r = api.murty_batch(pi, 60, weight_mode=api.WEIGHTS_GATED)
print("murty fast (ties -> exact fallback) nFound", r.n_found.tolist())
api.set_murty_path("auto")
A = synth.dense_square(1, 18, first=3)[0].reshape(18, 18, order="F")
print("perm n=18 (fused finalisation over several CTAs)", api.permanentExact(A))
print("range", api.permanent_range(A, 0, 1 << 16))
R = [synth.dense_square(1, 9, first=40 + i)[0].reshape(9, 9, order="F")[:4 + i % 3, :] for i in range(6)]
print("perm DP", api.permanent_batch(R)[0])
