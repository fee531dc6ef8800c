"""Generates tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):

    make -C oracle ref && python tests/golden/make_golden.py

It drives oracle/_ref/libpda_ref_strict.so -- the reference's own shortestPathCPP.cpp,
assignment.cpp and nwPerm.cpp code compiled IEEE-strict (see oracle/Makefile) -- on the
seeded inputs of probabilisticsemslam_b200/synth.py and stores inputs + outputs.  The
reference ships no golden vectors for this path; these files are
what pins the oracle (tests/test_oracle_golden.py) and, through it and directly, the
CUDA path (tests/test_gpu_*.py) on machines where the reference is not present.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.loader import load_reference  # noqa: E402
from probabilisticsemslam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
INF = np.inf


def kbest_case(R, name, mats, k, *, cutoff=42.0, use_cutoff=True, maximize=False):
    """Stores, per problem, nFound and the first nFound hypotheses (int16 lists, float64 gains)."""
    rec = {"k": np.int64(k), "cutoff": np.float64(cutoff), "use_cutoff": np.int64(use_cutoff),
           "maximize": np.int64(maximize), "n": np.int64(len(mats))}
    for i, m in enumerate(mats):
        m = np.asarray(m, np.float64)
        if use_cutoff:
            n, r4c, c4r, g = R.kbest2d_cutoff(k, m, cutoff, maximize)
        else:
            n, r4c, c4r, g = R.kbest2d(k, m, maximize)
        rec[f"C{i}"] = m
        rec[f"n{i}"] = np.int64(n)
        rec[f"r{i}"] = r4c[:n].astype(np.int16)
        rec[f"c{i}"] = c4r[:n].astype(np.int16)
        rec[f"g{i}"] = g[:n].copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "problems", len(mats), "k", k)


def main():
    R = load_reference("strict")
    print("reference build flags:", R.flags)

    # --- SURVEY appendix A known-answer vector -----------------------------------------------------
    appA = np.array([[1, 2], [4, 3], [6, 7], [10, INF], [INF, 10]], dtype=np.float64)
    kbest_case(R, "kbest_appendixA", [appA], 30)
    np.savez_compressed(os.path.join(OUT, "probs_appendixA.npz"), C=appA, nL=np.int64(3), k=np.int64(30),
                        probs=R.assignment_prob(appA, 3, 30))

    # --- config 1: one G1 5x30 problem, k = 200 ------------------------------------------------------
    c1 = synth.g1_dense(1, nM=5)
    kbest_case(R, "kbest_config1", [c1.matrix(0)], 200)
    np.savez_compressed(os.path.join(OUT, "probs_config1.npz"), C=c1.matrix(0), nL=np.int64(30),
                        **{f"k{k}": R.assignment_prob(c1.matrix(0), 30, k) for k in (1, 20, 100, 200, 1000)})

    # --- slices of config 2 / 3 ----------------------------------------------------------------------
    g1 = synth.g1_dense(16)
    kbest_case(R, "kbest_g1_k200", [g1.matrix(p) for p in range(16)], 200)
    kbest_case(R, "kbest_g1_k1000", [g1.matrix(p) for p in range(3)], 1000)
    g1i = synth.g1_dense(16, integer=True)
    kbest_case(R, "kbest_g1int_k200", [g1i.matrix(p) for p in range(16)], 200)
    kbest_case(R, "kbest_g1int_nocut_k300", [g1i.matrix(p) for p in range(4)], 300, use_cutoff=False)
    g2 = synth.g2_gated(24)
    cond = [R.condition_costs(g2.matrix(p)) for p in range(24)]
    kbest_case(R, "kbest_g2cond_k200", [c for c, _ in cond], 200)
    np.savez_compressed(os.path.join(OUT, "condition_g2.npz"), n=np.int64(24),
                        **{f"in{p}": g2.matrix(p) for p in range(24)},
                        **{f"out{p}": cond[p][0] for p in range(24)},
                        **{f"idx{p}": cond[p][1] for p in range(24)})

    # --- edge cases ----------------------------------------------------------------------------------
    edge = []
    e = g1.matrix(0).copy(); e[:, 1] = INF; edge.append(e)                       # infeasible column -> 0 found
    edge.append(np.array([[3.0], [7.5], [1.25], [10.0]]))                       # nM = 1
    edge.append(np.array([[10.0, INF], [INF, 10.0]]))                           # nL = 0: one assignment only
    edge.append(np.array([[1.0, 2.0], [2.0, 1.0], [10.0, INF], [INF, 10.0]]))   # k larger than #assignments
    e = synth.g1_dense(1, nM=3).matrix(0).copy(); e[:30, :] *= 5.0; edge.append(e)  # spread costs: cutoff break
    edge.append(np.zeros((6, 3)))                                                # every gain ties
    edge.append(np.array([[5.0, 5.0, 5.0], [5.0, 5.0, 5.0], [5.0, 5.0, 5.0]]))   # square, all ties
    kbest_case(R, "kbest_edges_k50", edge, 50)
    kbest_case(R, "kbest_edges_nocut_k50", edge, 50, use_cutoff=False)
    # maximisation: the asgnBB usage (k = 1, IoU-like scores, -inf off-diagonal dummies; assignment.cpp:749-750, 781)
    mx = []
    for p in range(6):
        m = synth.g1_dense(1, nM=3 + p % 3, nL=4 + p, first=100 + p).matrix(0)
        s = np.where(np.isfinite(m), m / 40.0, -INF)
        nl = 4 + p
        for c in range(s.shape[1]):
            s[nl + c, c] = 0.6
        mx.append(s)
    kbest_case(R, "kbest_maximize_k1", mx, 1, use_cutoff=False, maximize=True)
    kbest_case(R, "kbest_maximize_k40", mx, 40, use_cutoff=False, maximize=True)
    kbest_case(R, "kbest_maximize_cut_k40", mx, 40, cutoff=0.5, use_cutoff=True, maximize=True)

    # sticky ScratchSpace: kBest2D after kBest2DCutoff on the same workspace
    st = {}
    for i in range(4):
        m = g1i.matrix(i)
        n, r4c, c4r, g = R.kbest2d_after_cutoff(120, m, False, m, False, 6.0)
        st.update({f"C{i}": m, f"n{i}": np.int64(n), f"r{i}": r4c[:n].astype(np.int16),
                   f"c{i}": c4r[:n].astype(np.int16), f"g{i}": g[:n].copy()})
    np.savez_compressed(os.path.join(OUT, "kbest_sticky_k120.npz"), n=np.int64(4), k=np.int64(120),
                        first_cutoff=np.float64(6.0), **st)

    # --- plain LAP: assign2D and shortestPathCPP with duals ------------------------------------------
    lap = {}
    for i in range(8):
        m = synth.g1_dense(1, nM=3 + i % 6, nL=6 + 3 * i, first=200 + i).matrix(0)
        ret, r4c, c4r, u, v, g = R.assign2d(m, False)
        lap.update({f"C{i}": m, f"ret{i}": np.int64(ret), f"r{i}": r4c, f"c{i}": c4r, f"u{i}": u, f"v{i}": v, f"g{i}": np.float64(g)})
        safe = np.where(np.isfinite(m), m - m[np.isfinite(m)].min(), INF)
        ret, r4c, c4r, u, v, g, fb = R.shortest_path(safe)
        lap.update({f"S{i}": safe, f"sret{i}": np.int64(ret), f"sr{i}": r4c, f"sc{i}": c4r, f"su{i}": u, f"sv{i}": v,
                    f"sg{i}": np.float64(g), f"sf{i}": fb})
    np.savez_compressed(os.path.join(OUT, "lap.npz"), n=np.int64(8), **lap)

    # --- weights --------------------------------------------------------------------------------------
    w = {"n": np.int64(12)}
    for p in range(12):
        c, _ = cond[p]
        nl = c.shape[0] - c.shape[1]
        w[f"C{p}"] = c
        w[f"nL{p}"] = np.int64(nl)
        w[f"k200_{p}"] = R.assignment_prob(c, nl, 200)
        w[f"k20_{p}"] = R.assignment_prob(c, nl, 20)
        if c.shape[0] <= 18:
            w[f"bf_{p}"] = R.brute_force_prob(c, nl)
            st_, pp = R.permanent_prob(c, nl, 1)
            assert st_ == 0
            w[f"pp_{p}"] = pp
    np.savez_compressed(os.path.join(OUT, "weights_g2cond.npz"), **w)
    w = {"n": np.int64(8)}
    for p in range(8):
        w[f"C{p}"] = g1.matrix(p)
        w[f"k200_{p}"] = R.assignment_prob(g1.matrix(p), 30, 200)
    np.savez_compressed(os.path.join(OUT, "weights_g1.npz"), **w)
    v = synth.u01(synth.stream(np.arange(3, dtype=np.uint64), 40, synth.SEED + 5)) * 60.0
    np.savez_compressed(os.path.join(OUT, "toprobs.npz"), n=np.int64(3), **{f"in{i}": v[i] for i in range(3)},
                        **{f"out{i}": R.to_probs(v[i]) for i in range(3)})

    # --- getAssignmentProbs from the cost matrix on (assignment.cpp:57-74) ---------------------------------------
    a = {"n": np.int64(16)}
    for p in range(16):
        a[f"C{p}"] = g2.matrix(p)
        st_, pr = R.association_probs(g2.matrix(p), 30, 200, False)
        assert st_ == 0
        a[f"k200_{p}"] = pr
        if cond[p][0].shape[0] <= 20:
            st_, pr = R.association_probs(g2.matrix(p), 30, 200, True)
            assert st_ == 0
            a[f"perm_{p}"] = pr
    np.savez_compressed(os.path.join(OUT, "association_g2.npz"), **a)

    # --- stereo box association: asgnBB + computeBBCostMatrix + boundBox::IoU (reference's own code) ---------------
    L, Rt = synth.stereo_boxes(40)
    bb = {"n": np.int64(40), "nonassign": np.float64(0.6)}
    for i in range(40):
        bb[f"L{i}"], bb[f"R{i}"] = L[i], Rt[i]
        bb[f"a{i}"] = R.asgn_bb(L[i], Rt[i], 0.6)
        if len(L[i]) and len(Rt[i]):
            bb[f"C{i}"] = R.bb_cost_matrix(L[i], Rt[i], 0.6)
    np.savez_compressed(os.path.join(OUT, "asgn_bb.npz"), **bb)

    # --- permanents ------------------------------------------------------------------------------------
    pm = {}
    dims = list(range(1, 21)) + [22]
    for n in dims:
        A = synth.dense_square(2, n)
        for i in range(2):
            a = A[i].reshape(n, n, order="F")
            val, st_ = R.permanent_exact_square(a)
            pm[f"A_{n}_{i}"] = a
            pm[f"p_{n}_{i}"] = np.float64(val)
    rect = [(2, 3), (3, 7), (7, 3), (4, 12), (12, 5), (1, 6), (5, 18)]
    for (r, c) in rect:
        a = synth.dense_square(1, max(r, c), first=50 + r)[0][:r * c].reshape(r, c, order="F")
        pm[f"R_{r}_{c}"] = a
        pm[f"rp_{r}_{c}"] = np.float64(R.permanent_exact(a)[0])
    np.savez_compressed(os.path.join(OUT, "permanent.npz"), dims=np.asarray(dims), rect=np.asarray(rect), **pm)
    # conditionedPermanent on sparse likelihood-shaped matrices
    cp = {"n": np.int64(10)}
    for i in range(10):
        c, _ = cond[i]
        P = R.to_probs(c.reshape(-1, order="F")).reshape(c.shape, order="F")
        sub = P[1:, 1:]
        if max(np.count_nonzero(sub.max(axis=1) > 0), sub.shape[1]) > 32:
            sub = sub[:20]
        cp[f"A{i}"] = sub
        val, st_ = R.conditioned_permanent(sub, 1)
        cp[f"v{i}"] = np.float64(val)
        cp[f"s{i}"] = np.int64(st_)
    np.savez_compressed(os.path.join(OUT, "conditioned_permanent.npz"), **cp)
    print("done")


if __name__ == "__main__":
    main()
