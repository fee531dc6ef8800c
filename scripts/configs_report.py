"""BASELINE.json configs 1, 3 and 4 on the current GPU (config 2 and 5 are bench.py itself).
Writes one JSON document to stdout; run through gpurun and keep the output under profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probabilisticsemslam_b200 import api, synth, device as dev
from oracle.loader import load_oracle, load_reference, reference_available

out = {}
chk = load_reference("timing") if (reference_available("fast") or reference_available("native")) else load_oracle()

# ---- config 1: ONE 5x30 problem, k = 200, through the reference-shaped call (host buffers, batch of one) --------
c1 = synth.g1_dense(1, nM=5)
C = c1.matrix(0)
for _ in range(5):
    api.assignmentProb(C, 30, 200)
t = []
for _ in range(50):
    t0 = time.perf_counter(); p_gpu = api.assignmentProb(C, 30, 200); t.append(time.perf_counter() - t0)
tc = []
for _ in range(20):
    t0 = time.perf_counter(); p_cpu = chk.assignment_prob(C, 30, 200); tc.append(time.perf_counter() - t0)
lat = {}
for k in (1, 20, 100, 200, 1000):
    for _ in range(3): api.assignmentProb(C, 30, k)
    tk = []
    for _ in range(20):
        t0 = time.perf_counter(); api.assignmentProb(C, 30, k); tk.append(time.perf_counter() - t0)
    tk2 = []
    for _ in range(5):
        t0 = time.perf_counter(); chk.assignment_prob(C, 30, k); tk2.append(time.perf_counter() - t0)
    lat[k] = {"gpu_us": 1e6 * float(np.median(tk)), "cpu_us": 1e6 * float(np.median(tk2))}
out["config1_single_5x30"] = {"call": "assignmentProb (batch of one, host buffers, includes H2D/D2H and launch)",
                              "k200_gpu_us_median": 1e6 * float(np.median(t)), "k200_cpu_us_median": 1e6 * float(np.median(tc)),
                              "max_abs_diff_vs_cpu": float(np.max(np.abs(p_gpu - p_cpu))), "by_k": lat,
                              "note": "one problem occupies one warp; the GPU is built for batches (config 2)"}

# ---- config 3: 100k problems at k = 1000 (1 GPU share of it) ---------------------------------------------------------
n3 = int(os.environ.get("CFG3_PROBLEMS", "100000"))
pb = synth.g1_dense(n3)
plan = dev.MurtyPlan(pb, k=1000, weights=True)
for _ in range(2): plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): plan.run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nf = plan.n_found.cpu().numpy()
out["config3_k1000"] = {"problems": n3, "ms_per_pass": ms, "problems_per_s": n3 / (ms * 1e-3), "mean_found": float(nf.mean()),
                        "algorithmic_GB": plan.algorithmic_bytes() / 1e9, "workspace_GB": plan.workspace_bytes / 1e9}
cpu = chk.batch(synth.g1_dense(800), 1000, threads=os.cpu_count(), want_probs=True, want_lists=True)
out["config3_k1000"]["cpu_problems_per_s_all_threads"] = 800 / cpu["seconds"]
del plan
torch.cuda.empty_cache()

# ---- config 4: accuracy sweep -- batched exact permanents n = 12..20, and permanent weights vs k-best weights -----------
sweep = {}
for n in range(12, 21):
    A = synth.dense_square(64, n, first=100 * n)
    got, st = api.permanent_batch([a.reshape(n, n, order="F") for a in A])
    _, want = chk.permanent_batch(A, n, threads=os.cpu_count())
    sweep[n] = {"max_rel_diff_vs_cpu": float(np.max(np.abs(got - want) / np.abs(want)))}
g2 = synth.g2_gated(400, first=10_000)
cond, _ = api.condition_costs_batch(g2)
keep = [p for p in range(len(cond)) if cond.matrix(p).shape[0] - 1 <= 20]
sub = synth.pack([cond.matrix(p) for p in keep], [int(cond.nL[p]) for p in keep])
t0 = time.perf_counter(); tabs, st = api.permanent_prob_batch(sub, 1); t_perm = time.perf_counter() - t0
errs = {}
for k in (1, 20, 100, 200, 1000):
    t0 = time.perf_counter(); r = api.assignment_prob_batch(sub, k); tk = time.perf_counter() - t0
    e = [float(np.max(np.abs(r.prob_table(sub, i) - tabs[i]))) for i in range(len(sub))]
    errs[k] = {"median_max_abs_err_vs_permanent": float(np.median(e)), "p95": float(np.percentile(e, 95)), "worst": float(np.max(e)),
               "batch_seconds": tk}
out["config4_accuracy_sweep"] = {"permanent_batch_n12_20": sweep, "gated_problems": len(sub),
                                 "dims": [int(np.min(sub.num_row)), int(np.max(sub.num_row))],
                                 "permanent_prob_batch_seconds": t_perm, "kbest_vs_permanent_weights": errs}
print(json.dumps(out, indent=1))
