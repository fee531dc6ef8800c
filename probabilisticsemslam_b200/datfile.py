"""The reference's cost-matrix wire format.

`saveAssignmentProb` (assignment.cpp:821-831, called from system.cpp:271-272) writes one cost matrix per
frame to generatedData/<seq>/costMatrices/<ID>_frame<N>.dat: one CSV line per ROW (landmarks, then one
dummy row per detection), values printed with std::to_string (fixed, 6 decimals), +inf printed as "inf".
`getCosts` (comparison.cpp:32-57) reads it back: a token starting with 'i' is +inf, anything else std::stod.
The 6-decimal quantisation makes exact cost ties common, which is what the tie-break emulation is for.
"""
from __future__ import annotations

import math
import os

import numpy as np


def frame_path(directory: str, run_id: str, frame: int) -> str:
    """<dir>/<ID>_frame<N>.dat (comparison.cpp:33-34)."""
    return os.path.join(directory, f"{run_id}_frame{frame}.dat")


def _to_string(x: float) -> str:
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    if math.isnan(x):
        return "nan"
    return f"{x:f}"  # std::to_string(double) == printf("%f")


def write_dat(path: str, C: np.ndarray) -> None:
    C = np.asarray(C, dtype=np.float64)
    with open(path, "w") as f:
        for r in range(C.shape[0]):
            f.write(",".join(_to_string(float(v)) for v in C[r]) + "\n")


def read_dat(path: str) -> np.ndarray:
    rows = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            rows.append([np.inf if tok[0] == "i" else float(tok) for tok in line.split(",")])
    return np.asarray(rows, dtype=np.float64)
