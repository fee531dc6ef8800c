"""Latency of small batches (BASELINE.json configs[0]: the per-frame call of the SLAM loop) for both Murty kernels:
the reference-shaped host call (H2D + launch + D2H included, wall clock) and the kernel alone (CUDA events).
Prints one JSON document; run through gpurun and keep the output under profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probabilisticsemslam_b200 import api, synth, device as dev
from oracle.loader import load_oracle, load_reference, reference_available

chk = load_reference("timing") if (reference_available("fast") or reference_available("native")) else load_oracle()
out = {"host_call_us": {}, "kernel_us": {}, "cpu_us": {}}
C = synth.g1_dense(1, nM=5).matrix(0)


def med(f, reps, warm=3):
    for _ in range(warm): f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(t))


for k in (1, 20, 200, 1000):
    out["cpu_us"][k] = med(lambda: chk.assignment_prob(C, 30, k), 5 if k > 200 else 15, 1)
    for path in ("warp", "cta"):
        api.set_murty_path(path)
        out["host_call_us"][f"{path}_k{k}"] = med(lambda: api.assignmentProb(C, 30, k), 30)
api.set_murty_path("auto")
ref = None
for path in ("warp", "cta"):
    api.set_murty_path(path)
    for n in (1, 8, 32, 148, 296, 592, 1184, 2368):
        for k in (200,) if n > 1 else (1, 20, 200, 1000):
            pb = synth.g1_dense(n, first=77)
            plan = dev.MurtyPlan(pb, k=k, weights=True)
            for _ in range(3): plan.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps): plan.run()
            e1.record(); torch.cuda.synchronize()
            out["kernel_us"][f"{path}_n{n}_k{k}"] = 1e3 * e0.elapsed_time(e1) / reps
            del plan
api.set_murty_path("auto")
print(json.dumps(out, indent=1))
