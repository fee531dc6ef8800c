import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from probabilisticsemslam_b200 import api, synth, device as dev, _lib
api.set_murty_path("cta")
L = _lib.lib()
names = "ns root commit select tasks final total rounds ntasks commits pop push".split()
for nM, k in ((5, 200),):
    pb = synth.g1_dense(1, nM=nM)
    plan = dev.MurtyPlan(pb, k=k, weights=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(4): plan.run()
    torch.cuda.synchronize()
    e0.record()
    for rep in range(4): plan.run()
    e1.record(); torch.cuda.synchronize()
    print("EVENTS nM", nM, "k", k, "us per run", 1e3 * e0.elapsed_time(e1) / 4, flush=True)
    if hasattr(L, "pda_debug_read_prof"):
        buf = (ctypes.c_longlong * 16)()
        L.pda_debug_read_prof(buf)
        print("   ", {n: int(buf[i]) for i, n in enumerate(names)}, flush=True)

    if hasattr(L, "pda_debug_read_trace"):
        tr = (ctypes.c_longlong * 256)()
        L.pda_debug_read_trace(tr)
        for r in range(30):
            print("   round", r, "w0", tr[4*r], "tasks_cycles", tr[4*r+1], "ntasks", tr[4*r+2], "sweep", tr[4*r+3])
