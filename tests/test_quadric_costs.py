"""Cost matrices from quadric moments (computeQuadricCostMatrix, assignment.cpp:705-722; getCovs,
:693-703; getAssignmentProbs from the moments on, :38-74).

Eigen is not installed, so the reference's own `ldlt().solve` cannot be compiled here: the oracle restates Eigen 3.4's
pivoted LDLT (oracle/oracle_quadric.c, "parity unpinned").  The CPU test pins that restatement against an independent
solver (numpy.linalg.solve); the GPU tests compare the CUDA path with the oracle and with numpy.  Tolerance for this
floating-point row: 1e-9 relative (BASELINE.json north_star); observed ~1e-15."""
import numpy as np
import pytest

from probabilisticsemslam_b200 import synth

RTOL = 1e-9
NONASSIGN = 10.0


def _numpy_costs(lm, lc, mm, mc, nonassign):
    nL, nM = lm.shape[0], mm.shape[0]
    C = np.full((nL + nM, nM), np.inf)
    for c in range(nM):
        for r in range(nL):
            d = lm[r] - mm[c]
            C[r, c] = d @ np.linalg.solve(lc[r] + mc[c], d)
        C[nL + c, c] = nonassign
    return C


def _hard_frames():
    """Shapes that exercise every pivot order of the 3x3 LDLT, wide dynamic range, d == 0, and empty sides."""
    rng = np.random.default_rng(5)
    frames = []
    for perm in ([0, 1, 2], [1, 0, 2], [2, 1, 0], [0, 2, 1], [2, 0, 1], [1, 2, 0]):
        diag = np.array([9.0, 4.0, 1.0])[perm]
        A = np.diag(diag) + 0.1 * np.ones((3, 3))
        lm = rng.normal(size=(4, 3)) * 3
        frames.append((lm, np.repeat(A[None] * 0.5, 4, 0), lm[:2] + 0.3, np.repeat(A[None] * 0.5, 2, 0)))
    Q = rng.normal(size=(5, 3, 3))
    S = np.einsum("nij,nj,nkj->nik", Q, 10.0 ** rng.uniform(-4, 4, size=(5, 3)), Q)
    S = 0.5 * (S + S.transpose(0, 2, 1))
    frames.append((rng.normal(size=(5, 3)), S, rng.normal(size=(3, 3)), S[:3] * 0.25))
    lm = rng.normal(size=(3, 3))
    frames.append((lm, np.repeat(np.eye(3)[None], 3, 0), lm.copy(), np.repeat(np.eye(3)[None], 3, 0)))  # d == 0 on the diagonal
    frames.append((np.zeros((0, 3)), np.zeros((0, 3, 3)), rng.normal(size=(2, 3)), np.repeat(np.eye(3)[None], 2, 0)))  # nL == 0
    frames.append((rng.normal(size=(3, 3)), np.repeat(np.eye(3)[None], 3, 0), np.zeros((0, 3)), np.zeros((0, 3, 3))))  # nM == 0
    return frames


def test_oracle_ldlt_against_numpy(oracle):
    worst = 0.0
    for f in synth.quadric_frames(40, first=11) + _hard_frames():
        got, want = oracle.quadric_cost_matrix(*f, NONASSIGN), _numpy_costs(*f, NONASSIGN)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        if fin.any():
            cond = max(np.linalg.cond(a + b) for a in f[1] for b in f[3]) if len(f[1]) and len(f[3]) else 1.0
            err = np.max(np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300) * (want[fin] != 0))
            assert err <= max(RTOL, 1e-15 * cond), (err, cond)
            worst = max(worst, err)
    assert worst < 1e-6


def _golden_frames():
    from helpers import golden
    z = golden("quadric_costs")
    for i in range(int(z["n"])):
        yield (z[f"lm{i}"], z[f"lc{i}"], z[f"mm{i}"], z[f"mc{i}"]), z[f"cost{i}"], z[f"probs{i}"], float(z["nonassign"]), int(z["k"])


def test_oracle_against_50_digit_truth(oracle):
    """tests/golden/quadric_costs.npz holds the costs computed with 50 significant digits (mpmath) and rounded once:
    a pin of the LDLT restatement that does not depend on any double-precision solver."""
    for f, cost, probs, na, k in _golden_frames():
        got = oracle.quadric_cost_matrix(*f, na)
        fin = np.isfinite(cost)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], cost[fin], rtol=1e-12)
        np.testing.assert_allclose(oracle.association_from_moments(*f, na, k), probs, rtol=1e-12, atol=1e-300)


@pytest.mark.gpu
def test_gpu_against_50_digit_truth(gpu_api):
    frames, costs, probs = [], [], []
    for f, cost, pr, na, k in _golden_frames():
        frames.append(f); costs.append(cost); probs.append(pr)
    got = gpu_api.quadric_cost_batch(frames, na)
    for g, want in zip(got, costs):
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(g), fin)
        np.testing.assert_allclose(g[fin], want[fin], rtol=RTOL)          # observed ~1e-15
    for g, want in zip(gpu_api.association_from_moments_batch(frames, na, k), probs):
        np.testing.assert_allclose(g, want, rtol=RTOL, atol=1e-300)


def test_oracle_getcovs(oracle):
    rng = np.random.default_rng(1)
    Q = rng.normal(size=(7, 4, 4)); Q = Q + Q.transpose(0, 2, 1)
    want = Q[:, :3, :3] + Q[:, :3, 3:4] * Q[:, None, :3, 3]
    np.testing.assert_allclose(oracle.quadric_covs(Q), want, rtol=1e-15)


@pytest.mark.gpu
def test_gpu_cost_matrices_vs_oracle_and_numpy(gpu_api, oracle):
    frames = synth.quadric_frames(300, first=500) + _hard_frames()
    got = gpu_api.quadric_cost_batch(frames, NONASSIGN)
    exact = 0
    for f, g in zip(frames, got):
        want = oracle.quadric_cost_matrix(*f, NONASSIGN)
        assert g.shape == want.shape
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(g), fin)
        np.testing.assert_allclose(g[fin], want[fin], rtol=1e-12, atol=0)
        exact += int(np.array_equal(g.view(np.int64), want.view(np.int64)))
    assert exact == len(frames), f"only {exact} of {len(frames)} matrices are bit-identical to the C restatement"
    for f, g in list(zip(frames, got))[:40]:
        want = _numpy_costs(*f, NONASSIGN)
        fin = np.isfinite(want) & (want != 0)
        np.testing.assert_allclose(g[fin], want[fin], rtol=RTOL)


@pytest.mark.gpu
def test_gpu_getcovs(gpu_api, oracle):
    rng = np.random.default_rng(2)
    Q = rng.normal(size=(100, 4, 4)); Q = Q + Q.transpose(0, 2, 1)
    np.testing.assert_allclose(gpu_api.getCovs(Q), oracle.quadric_covs(Q), rtol=1e-15)


@pytest.mark.gpu
@pytest.mark.usefixtures("murty_path")
def test_gpu_association_from_moments(gpu_api, oracle):
    """Moments in, weights out, one device pipeline, against the CPU chain cost matrix -> conditionCosts ->
    assignmentProb -> un-compaction."""
    frames = synth.quadric_frames(120, first=900) + _hard_frames()
    got = gpu_api.association_from_moments_batch(frames, NONASSIGN, 200)
    for i, (f, g) in enumerate(zip(frames, got)):
        want = oracle.association_from_moments(*f, NONASSIGN, 200)
        assert g.shape == want.shape
        np.testing.assert_allclose(g, want, rtol=RTOL, atol=1e-300, err_msg=f"frame {i}")
    one = gpu_api.getAssignmentProbs(*frames[0], NONASSIGN, 200)
    np.testing.assert_array_equal(one, got[0])
