"""Times one device-resident pass of the Murty batch under each kernel path and reports how many problems the
pruning kernel handed to the exact one (workspace header, bytes 8..11).  usage: python scripts/fast_path_probe.py [n] [k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from probabilisticsemslam_b200 import api, synth, device as dev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 200
integer = len(sys.argv) > 3 and sys.argv[3] == "int"
def dump_stats(tag):
    if not os.environ.get("PDA_B200_LIB"):
        return
    import ctypes
    from probabilisticsemslam_b200 import _lib
    st = (ctypes.c_ulonglong * 16)()
    try:
        _lib.lib().pda_debug_fast_stats(st, 1)
    except AttributeError:
        return
    names = ["children", "abandoned", "dropped_done", "kept", "tighten_calls", "slots_at_tighten", "pops", "-",
             "argmins", "ff_tried", "ff_applied", "loop_trips", "flip_hops", "searches_done", "relax_real", "relax_pad"]
    runs = 5 * n
    print(tag, {k: round(int(v) / runs, 2) for k, v in zip(names, st)})


pb = synth.g1_dense(n, integer=integer)
dump_stats("reset")
ref = None
for path in os.environ.get("PROBE_PATHS", "warp,fast,auto").split(","):
    api.set_murty_path(path)
    plan = dev.MurtyPlan(pb, k=k, weights=True)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fb = int(plan.workspace[8:12].view(torch.int32).item())
    out = (plan.n_found.cpu().numpy(), plan.row4col.cpu().numpy(), plan.col4row.cpu().numpy(), plan.gain.cpu().numpy().view(np.int64), plan.probs.cpu().numpy())
    same = None
    if ref is None:
        ref = out
    else:
        nf = ref[0]
        same = bool(np.array_equal(out[0], nf) and np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2]))
        g_ok = all(np.array_equal(out[3][p * k:p * k + nf[p]], ref[3][p * k:p * k + nf[p]]) for p in range(0, n, max(1, n // 2000)))
        same = same and g_ok
        wdiff = float(np.nanmax(np.abs(out[4] - ref[4])))
    print(f"path {path:5s}: {ms:8.3f} ms  {n / ms * 1e3:12.0f} problems/s  fallback {fb}  same_as_warp {same}" + ("" if same is None else f" wdiff {wdiff:.2e}"))
    dump_stats(path)
    del plan
    torch.cuda.empty_cache()
api.set_murty_path("auto")
