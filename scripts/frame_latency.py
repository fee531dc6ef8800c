"""Latency of ONE gated frame (the SLAM per-frame call): C-ABI host calls from quadric moments and from the cost matrix,
beside the reference's own code (strict build; the -Ofast builds can loop forever on gated +inf matrices)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from probabilisticsemslam_b200 import api, synth, _lib
from oracle.loader import load_oracle, load_reference, reference_available
o = load_oracle()
ref = load_reference("strict") if reference_available("strict") else o   # the -Ofast builds can loop forever on gated (+inf) matrices
L = _lib.lib()
fr = synth.quadric_frames(8, first=50)
def med(f, reps=40, warm=5):
    for _ in range(warm): f()
    t=[]
    for _ in range(reps):
        t0=time.perf_counter(); f(); t.append(time.perf_counter()-t0)
    return 1e6*float(np.median(t))
for i, f in enumerate(fr[:6]):
    C = o.quadric_cost_matrix(*f, 10.0)
    cond, idx = o.condition_costs(C)
    cL = cond.shape[0] - cond.shape[1]
    pk = api._pack_moments([f])
    nL, nM = f[0].shape[0], f[2].shape[0]
    out = np.zeros(nM * (nL + 1))
    ptrs = [a.ctypes.data for a in pk[:6]]
    g_call = med(lambda: L.pda_association_from_moments_batch_host(*ptrs, 1, 10.0, 200, out.ctypes.data, 0))
    # from the cost matrix on (what getAssignmentProbsFromCosts does)
    flat = np.ascontiguousarray(C.reshape(-1, order="F")); off = np.zeros(1, np.int64); poff = np.zeros(1, np.int64)
    nl32 = np.array([nL], np.int32); nm32 = np.array([nM], np.int32)
    g_cost = med(lambda: L.pda_association_probs_batch_host(flat.ctypes.data, off.ctypes.data, nl32.ctypes.data, nm32.ctypes.data, 1, 200, out.ctypes.data, poff.ctypes.data, 0))
    c_ref = med(lambda: ref.association_probs(C, nL, 200), 10, 2)
    nf = api.murty_batch(synth.pack([cond], [cL]), 200).n_found[0]
    print("frame", i, "nM", nM, "cond rows", cond.shape[0], "nFound", int(nf), "gpu C-call from moments us", round(g_call), "from costs us", round(g_cost), "cpu reference (strict -O2) from costs us", round(c_ref))
