"""Times the permanent kernels on the current GPU: python scripts/perm_bench.py (prints one line per case)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probabilisticsemslam_b200 import synth, device as dev, _lib

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

peak = _lib.lib().pda_diag_dfma_tflops()
print("fp64 peak TFLOP/s", peak)
for n, m, reps in [(24, 1, 50), (28, 1, 10), (20, 2000, 5), (16, 20000, 5), (12, 100000, 5), (8, 100000, 5)]:
    A = synth.dense_square(m, n, first=n)
    plan = dev.PermanentPlan(A, n)
    ms = timeit(plan.run, reps)
    tf = plan.flops() / (ms * 1e-3) / 1e12
    print(json.dumps({"n": n, "mats": m, "ms": round(ms, 4), "tflops": round(tf, 2), "frac_of_fp64_peak": round(tf / peak, 3)}))
