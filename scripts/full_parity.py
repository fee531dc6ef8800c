"""Full-size parity: ALL 100 000 problems of BASELINE.json configs[1] (k = 200), every hypothesis, against the
reference's own code compiled IEEE-strict (oracle/_ref/libpda_ref_strict.so; falls back to the C restatement).
Index lists and gains bit for bit, weights to 1e-9 relative.  Prints one JSON line; ~2 minutes on the GPU box."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probabilisticsemslam_b200 import api, synth
from oracle.loader import load_oracle, load_reference, reference_available

N, K, CHUNK = int(os.environ.get("FULL_N", "100000")), int(os.environ.get("FULL_K", "200")), int(os.environ.get("FULL_CHUNK", "10000"))
chk, kind = (load_reference("strict"), "reference (oracle/_ref strict build)") if reference_available("strict") else (load_oracle(), "oracle restatement")
bad_lists = bad_gain = bad_found = 0
worst_w = 0.0
hyps = 0
t0 = time.time()
for first in range(0, N, CHUNK):
    pb = synth.g1_dense(min(CHUNK, N - first), first=first, integer=bool(int(os.environ.get("FULL_INT", "0"))))  # FULL_INT=1: exact-tie stress
    got = api.murty_batch(pb, K, weight_mode=api.WEIGHTS_GATED)
    want = chk.batch(pb, K, threads=os.cpu_count(), want_probs=True, want_lists=True)
    bad_found += int(np.count_nonzero(got.n_found != want["n_found"]))
    if not np.array_equal(got.row4col, want["row4col"]) or not np.array_equal(got.col4row, want["col4row"]):
        for p in range(len(pb)):
            r, c, g = got.lists(pb, p)
            n, nc, nr = int(want["n_found"][p]), int(pb.nM[p]), int(pb.nL[p] + pb.nM[p])
            ok = np.array_equal(r.reshape(-1), want["row4col"][want["r4c_off"][p]:want["r4c_off"][p] + n * nc]) and \
                np.array_equal(c.reshape(-1), want["col4row"][want["c4r_off"][p]:want["c4r_off"][p] + n * nr])
            bad_lists += 0 if ok else 1
    gg = got.gain.reshape(-1).view(np.int64)
    wg = want["gain"].view(np.int64)
    valid = (np.arange(K)[None, :] < want["n_found"][:, None]).reshape(-1)
    bad_gain += int(np.count_nonzero((gg != wg) & valid))
    hyps += int(want["n_found"].sum())
    d = np.abs(got.probs - want["probs"]) / np.maximum(np.abs(want["probs"]), 1e-300)
    worst_w = max(worst_w, float(np.max(np.where(want["probs"] == got.probs, 0.0, d))))
print(json.dumps({"problems": N, "k": K, "integer_costs": bool(int(os.environ.get("FULL_INT", "0"))), "hypotheses_compared": hyps, "checker": kind, "problems_with_different_nFound": bad_found,
                  "problems_with_different_lists": bad_lists, "gains_with_different_bits": bad_gain,
                  "worst_relative_weight_difference": worst_w, "seconds": round(time.time() - t0, 1)}))
