"""Generates the golden files of the two rows whose reference side cannot be compiled in the build container:

  quadric_costs.npz   computeQuadricCostMatrix (assignment.cpp:705-722) -- the reference calls Eigen's ldlt().solve
                      (Eigen is not installed).  The stored costs are a 50-digit TRUTH (mpmath: d^T S^-1 d solved
                      exactly to working precision, rounded once to double), so they pin the restatement in
                      oracle/oracle_quadric.c and the CUDA kernel independently of any double-precision solver; the
                      stored weights are the restatement's (getAssignmentProbs chain, k = 200).
  perm_approx.npz     Huber's approximation (nwPerm.cpp:126-211) -- the reference draws from an unseeded rand(); the
                      stored estimates and success counts come from oracle/oracle_perm_approx.c on the counter-based
                      stream (seed 20260217, matrix index = position), the exact permanents from the oracle's NW walk.

    python tests/golden/make_golden_restated.py
"""
from __future__ import annotations

import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.loader import load_oracle  # noqa: E402
from probabilisticsemslam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NONASSIGN, K = 10.0, 200


def truth_costs(lm, lc, mm, mc):
    mp.mp.dps = 50
    nL, nM = lm.shape[0], mm.shape[0]
    C = np.full((nL + nM, nM), np.inf)
    for c in range(nM):
        for r in range(nL):
            d = mp.matrix([mp.mpf(float(lm[r, i])) - mp.mpf(float(mm[c, i])) for i in range(3)])
            S = mp.matrix(3, 3)
            for i in range(3):
                for j in range(3):
                    S[i, j] = mp.mpf(float(lc[r, i, j])) + mp.mpf(float(mc[c, i, j]))
            x = mp.lu_solve(S, d)
            C[r, c] = float(sum(d[i] * x[i] for i in range(3)))
        C[nL + c, c] = NONASSIGN
    return C


def main():
    orc = load_oracle()
    frames = synth.quadric_frames(12, first=2024)
    rec = {"n": np.int64(len(frames)), "nonassign": np.float64(NONASSIGN), "k": np.int64(K)}
    for i, (lm, lc, mm, mc) in enumerate(frames):
        rec[f"lm{i}"], rec[f"lc{i}"], rec[f"mm{i}"], rec[f"mc{i}"] = lm, lc, mm, mc
        rec[f"cost{i}"] = truth_costs(lm, lc, mm, mc)
        rec[f"probs{i}"] = orc.association_from_moments(lm, lc, mm, mc, NONASSIGN, K)
    np.savez_compressed(os.path.join(OUT, "quadric_costs.npz"), **rec)

    mats = [synth.dense_square(1, n, first=500 + n)[0].reshape(n, n, order="F") for n in (3, 6, 9, 12, 15, 18)]
    rng = np.random.default_rng(4)
    mats += [rng.random((3, 7)), rng.random((8, 5))]
    pa = {"n": np.int64(len(mats)), "iterations": np.int64(300), "seed": np.int64(20260217)}
    for i, A in enumerate(mats):
        est, succ = orc.permanent_approx(A, 300, 20260217, i)
        pa[f"A{i}"], pa[f"est{i}"], pa[f"succ{i}"], pa[f"exact{i}"] = A, np.float64(est), np.int64(succ), np.float64(orc.permanent_exact(A)[0])
    np.savez_compressed(os.path.join(OUT, "perm_approx.npz"), **pa)
    print("wrote quadric_costs.npz, perm_approx.npz")


if __name__ == "__main__":
    main()
