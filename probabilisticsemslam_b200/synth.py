"""Seeded synthetic association problems.

One counter-based splitmix64 stream per problem, so the same inputs can be
regenerated bit-for-bit from Python, C or CUDA:

    seed_p  = (0x9E3779B97F4A7C15 * (p + 1)) XOR SEED          (mod 2^64)
    x_j     = mix(seed_p + (j + 1) * 0x9E3779B97F4A7C15)       j = 0, 1, 2, ...
    u01     = (x >> 11) * 2^-53

Problem layout follows the reference's cost matrices (assignment.cpp:705-722):
column-major ``C[row + col * numRow]``, one column per detection, ``nL`` landmark
rows followed by one dummy ("missed detection") row per detection; the dummy block
is +inf except its diagonal, which holds NONASSIGN_QUADRIC = 10 (runOpts/calibSample.txt:7).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 20260217
_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
NONASSIGN = 10.0


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= _M1
    z ^= z >> np.uint64(27)
    z *= _M2
    z ^= z >> np.uint64(31)
    return z


def stream(problem_ids: np.ndarray, n_draws: int, seed: int = SEED) -> np.ndarray:
    """uint64 draws x_0..x_{n_draws-1} for every problem id -> shape (len(ids), n_draws)."""
    with np.errstate(over="ignore"):
        p = np.asarray(problem_ids, dtype=np.uint64)
        base = (_GAMMA * (p + np.uint64(1))) ^ np.uint64(seed)
        j = np.arange(1, n_draws + 1, dtype=np.uint64)
        return _mix(base[:, None] + j[None, :] * _GAMMA)


def u01(x: np.ndarray) -> np.ndarray:
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


@dataclass
class ProblemBatch:
    """A ragged batch of association problems in the flat layout the C-ABI takes."""

    costs: np.ndarray     # float64, concatenated column-major matrices
    cost_off: np.ndarray  # int64[n]   start of problem p in ``costs``
    nL: np.ndarray        # int32[n]   landmark rows
    nM: np.ndarray        # int32[n]   detections (columns); numRow = nL + nM

    def __len__(self) -> int:
        return int(self.nL.shape[0])

    @property
    def num_row(self) -> np.ndarray:
        return (self.nL + self.nM).astype(np.int32)

    def matrix(self, p: int) -> np.ndarray:
        """(numRow, numCol) view-copy of problem p (Fortran order == the wire layout)."""
        nr, nc = int(self.nL[p] + self.nM[p]), int(self.nM[p])
        o = int(self.cost_off[p])
        return self.costs[o:o + nr * nc].reshape((nr, nc), order="F")

    def flat(self, p: int) -> np.ndarray:
        nr, nc = int(self.nL[p] + self.nM[p]), int(self.nM[p])
        o = int(self.cost_off[p])
        return self.costs[o:o + nr * nc]

    def slice(self, lo: int, hi: int) -> "ProblemBatch":
        o0 = int(self.cost_off[lo])
        o1 = int(self.cost_off[hi]) if hi < len(self) else int(self.costs.shape[0])
        return ProblemBatch(self.costs[o0:o1].copy(), (self.cost_off[lo:hi] - o0).copy(),
                            self.nL[lo:hi].copy(), self.nM[lo:hi].copy())

    def take(self, idx: np.ndarray) -> "ProblemBatch":
        return pack([self.matrix(int(i)) for i in idx], [int(self.nL[int(i)]) for i in idx])


def pack(mats, nLs) -> ProblemBatch:
    """Pack a list of (numRow, numCol) matrices into a ProblemBatch."""
    nL = np.asarray(nLs, dtype=np.int32)
    nM = np.asarray([m.shape[1] for m in mats], dtype=np.int32)
    sizes = np.asarray([m.size for m in mats], dtype=np.int64)
    off = np.zeros(len(mats), dtype=np.int64)
    if len(mats) > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    costs = np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1, order="F") for m in mats]) if mats else np.zeros(0)
    return ProblemBatch(np.ascontiguousarray(costs), off, nL, nM)


def g1_dense(n: int, *, nL: int = 30, nM: int | None = None, integer: bool = False,
             first: int = 0, seed: int = SEED, scale: float = 40.0) -> ProblemBatch:
    """G1 "dense30": nM = 3 + (x_0 mod 6) detections (or ``nM`` forced), ``nL`` landmarks,
    landmark costs ``scale * u01`` drawn column by column (x_1, x_2, ...), dummy diagonal 10.
    ``integer=True`` floors the landmark costs (G1-int: the exact-tie stress)."""
    ids = np.arange(first, first + n, dtype=np.uint64)
    max_m = 8 if nM is None else nM
    x = stream(ids, 1 + max_m * nL, seed)
    m = (np.uint64(3) + x[:, 0] % np.uint64(6)).astype(np.int32) if nM is None else np.full(n, nM, np.int32)
    land = scale * u01(x[:, 1:]).reshape(n, max_m, nL)
    if integer:
        land = np.floor(land)
    num_row = nL + m.astype(np.int64)
    sizes = num_row * m
    off = np.zeros(n, dtype=np.int64)
    if n > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    costs = np.full(int(sizes.sum()), np.inf, dtype=np.float64)
    for mm in np.unique(m):
        sel = np.nonzero(m == mm)[0]
        nr = nL + int(mm)
        blk = np.full((sel.size, int(mm), nr), np.inf)
        blk[:, :, :nL] = land[sel, :int(mm), :]
        d = np.arange(int(mm))
        blk[:, d, nL + d] = NONASSIGN
        idx = off[sel][:, None] + np.arange(int(mm) * nr)[None, :]
        costs[idx.reshape(-1)] = blk.reshape(-1)
    return ProblemBatch(costs, off, np.full(n, nL, np.int32), m)


def g2_gated(n: int, *, nL: int = 30, first: int = 0, seed: int = SEED + 1) -> ProblemBatch:
    """G2 "gated": squared distances between 3-D landmark and detection positions over a
    100 m x 100 m x 4 m scene (unit covariance), so most entries fall outside the
    colMin+42 gate and conditionCosts shrinks the problem, as on real KITTI frames.
    Each detection is a landmark plus N(0, 0.7 m) noise with probability ~0.8, else clutter."""
    ids = np.arange(first, first + n, dtype=np.uint64)
    draws = 1 + 3 * nL + 8 * 8
    x = stream(ids, draws, seed)
    m = (np.uint64(3) + x[:, 0] % np.uint64(6)).astype(np.int32)
    u = u01(x[:, 1:])
    land = u[:, :3 * nL].reshape(n, nL, 3) * np.array([100.0, 100.0, 4.0])
    d = u[:, 3 * nL:].reshape(n, 8, 8)
    mats = []
    for p in range(n):
        mm = int(m[p])
        pts = np.empty((mm, 3))
        for c in range(mm):
            if d[p, c, 0] < 0.8:
                src = land[p, int(d[p, c, 1] * nL) % nL]
                # Box-Muller-free bounded noise: sum of three uniforms, sd ~ 0.7 m
                pts[c] = src + (d[p, c, 2:5] + d[p, c, 5:8] - 1.0) * 1.7
            else:
                pts[c] = d[p, c, 2:5] * np.array([100.0, 100.0, 4.0])
        C = np.full((nL + mm, mm), np.inf)
        diff = land[p][:, None, :] - pts[None, :, :]
        C[:nL, :] = np.sum(diff * diff, axis=2)
        C[nL + np.arange(mm), np.arange(mm)] = NONASSIGN
        mats.append(C)
    return pack(mats, [nL] * n)


def quadric_frames(n: int, *, nL: int = 30, first: int = 0, seed: int = SEED + 4):
    """Frames for the cost-matrix builder (computeQuadricCostMatrix, assignment.cpp:705-722): per frame ``nL`` landmark
    quadrics scattered over a 100 m x 100 m x 4 m scene and 3-8 detected quadrics (a landmark plus ~0.7 m of noise
    with probability 0.8, else clutter).  Every quadric is a centroid and a 3x3 shape matrix R diag(r^2) R^T with
    radii 0.3-2.3 m and a random orientation.  Returns a list of (land_mean[nL,3], land_cov[nL,3,3],
    meas_mean[nM,3], meas_cov[nM,3,3])."""
    ids = np.arange(first, first + n, dtype=np.uint64)
    per = 3 + 3 + 4  # centroid, radii, quaternion
    x = stream(ids, 1 + per * (nL + 8) + 8 * 2, seed)
    m = (np.uint64(3) + x[:, 0] % np.uint64(6)).astype(np.int32)
    u = u01(x[:, 1:])

    def shapes(v):  # v[..., 7] -> cov[..., 3, 3]
        r2 = (0.3 + 2.0 * v[..., 0:3]) ** 2
        q = v[..., 3:7] * 2.0 - 1.0 + 1e-3
        q = q / np.linalg.norm(q, axis=-1, keepdims=True)
        w, a, b, c = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        R = np.stack([np.stack([1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)], -1),
                      np.stack([2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)], -1),
                      np.stack([2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)], -1)], -2)
        cov = np.einsum("...ij,...j,...kj->...ik", R, r2, R)
        return 0.5 * (cov + np.swapaxes(cov, -1, -2))  # exactly symmetric, as getCovs builds it

    frames = []
    scale = np.array([100.0, 100.0, 4.0])
    for p in range(n):
        q = u[p, :per * (nL + 8)].reshape(nL + 8, per)
        pick = u[p, per * (nL + 8):].reshape(8, 2)
        land_mean = q[:nL, 0:3] * scale
        land_cov = shapes(q[:nL, 3:10])
        mm = int(m[p])
        meas_mean = np.empty((mm, 3))
        for c in range(mm):
            if pick[c, 0] < 0.8 and nL > 0:
                meas_mean[c] = land_mean[int(pick[c, 1] * nL) % nL] + (q[nL + c, 0:3] - 0.5) * 2.4
            else:
                meas_mean[c] = q[nL + c, 0:3] * scale
        meas_cov = shapes(q[nL:nL + mm, 3:10])
        frames.append((land_mean, land_cov, meas_mean, meas_cov))
    return frames


def dense_square(n_mats: int, dim: int, *, first: int = 0, seed: int = SEED + 2) -> np.ndarray:
    """Permanent inputs: ``n_mats`` dense dim x dim matrices, entries u01, column-major,
    returned as float64[n_mats, dim*dim]."""
    ids = np.arange(first, first + n_mats, dtype=np.uint64)
    return np.ascontiguousarray(u01(stream(ids, dim * dim, seed)))


def stereo_boxes(n: int, *, first: int = 0, seed: int = SEED + 3):
    """Stereo detection pairs for the box association (asgnBB): per problem a list of right-image boxes in a
    1242 x 375 frame and the left-image boxes they came from (shifted by a disparity, jittered, some dropped, some
    spurious), as [count, 5] arrays (xmin, ymin, xmax, ymax, xOffset).  xOffset is the per-box stereo offset estimate
    the reference adds to the box the IoU is called on (boundBox.h:63-64).  Returns (list_left, list_right)."""
    ids = np.arange(first, first + n, dtype=np.uint64)
    u = u01(stream(ids, 2 + 12 * 8, seed))
    lefts, rights = [], []
    for p in range(n):
        nr = 1 + int(u[p, 0] * 10) % 10
        disp = 5.0 + 60.0 * u[p, 1]
        r = u[p, 2:].reshape(12, 8)
        R, L = [], []
        for i in range(nr):
            x0, y0 = 1100.0 * r[i, 0], 300.0 * r[i, 1]
            w, h = 30.0 + 140.0 * r[i, 2], 20.0 + 55.0 * r[i, 3]
            R.append([x0, y0, x0 + w, y0 + h, disp * (1.0 + 0.1 * (r[i, 4] - 0.5))])
            if r[i, 5] < 0.85:  # seen in the left image too
                j = 6.0 * (r[i, 6:8] - 0.5)
                L.append([x0 + disp + j[0], y0 + j[1], x0 + w + disp + j[0], y0 + h + j[1], -disp * (1.0 + 0.1 * (r[i, 4] - 0.5))])
        if r[10, 0] < 0.3:  # a spurious left detection
            L.append([1000.0 * r[10, 1], 250.0 * r[10, 2], 1000.0 * r[10, 1] + 60.0, 250.0 * r[10, 2] + 40.0, -disp])
        lefts.append(np.asarray(L, dtype=np.float64).reshape(-1, 5))
        rights.append(np.asarray(R, dtype=np.float64).reshape(-1, 5))
    return lefts, rights
