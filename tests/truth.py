"""Cancellation-free "truth" for permanent-based quantities (test infrastructure).

The NW/Ryser formula is an alternating sum and can lose many digits on sparse, wide-dynamic-range
matrices (the reference itself is then only ~1e-7 accurate).  For non-negative
matrices the permanent can instead be accumulated with non-negative terms only: a subset DP over the
SHORT side, visiting the long side's lines one by one.  perm(pad_with_ones(A)) / (|m-n|)!  of the
reference (nwPerm.cpp:217-231) equals the sum over injective maps short -> long, which is what the DP
returns, to ~1e-15 relative.
"""
import numpy as np


def perm_injective(A: np.ndarray) -> float:
    """Sum over injective maps from the short side of A into the long side of prod A[i, f(i)]."""
    A = np.asarray(A, dtype=np.float64)
    if A.shape[0] > A.shape[1]:
        A = A.T
    m, n = A.shape  # m <= n: every row gets a distinct column
    if m == 0:
        return 1.0
    f = np.zeros(1 << m)
    f[0] = 1.0
    masks = np.arange(1 << m)
    for c in range(n):
        g = f.copy()
        for i in range(m):
            bit = 1 << i
            src = masks[(masks & bit) == 0]
            g[src | bit] += f[src] * A[i, c]
        f = g
    return float(f[(1 << m) - 1])


def permanent_prob_truth(C: np.ndarray, nL: int) -> np.ndarray:
    """permanentProb (assignment.cpp:145-290) with every sub-permanent evaluated cancellation-free."""
    C = np.asarray(C, dtype=np.float64)
    nR, nM = C.shape
    lo = C.min()
    P = np.where(lo + 42.0 > C, np.exp(lo - C), 0.0)
    if nM == 1:
        return (P[:, 0] / P[:, 0].sum())[None, :]
    probs = np.zeros((nM, nL + 1))
    best = 0.0
    for m in range(nM):
        cols = [c for c in range(nM) if c != m]
        tot = 0.0
        for l in range(nL + 1):
            row = l if l < nL else nL + m
            if l < nL and P[l, m] == 0:
                continue
            rows = [r for r in range(nR) if r != row]
            sub = P[np.ix_(rows, cols)]
            sub = sub[sub.max(axis=1) > 0]          # all-zero rows cannot be used and are dropped by the reference too
            val = P[row, m] * perm_injective(sub)
            probs[m, l] = val
            tot += val
        best = max(best, tot)
    return probs / best
