"""CPU-only: pins the oracle restatement (oracle/*.c) against vectors produced by the
reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import KBEST_FILES, assert_kbest_equal, bits, golden, kbest_cases


@pytest.mark.parametrize("name", KBEST_FILES)
def test_kbest_bit_exact(oracle, name):
    z = golden(name)
    k, cutoff, use_cut, maxi = int(z["k"]), float(z["cutoff"]), bool(z["use_cutoff"]), bool(z["maximize"])
    for i, (C, n, r, c, g) in enumerate(kbest_cases(z)):
        got = oracle.kbest2d_cutoff(k, C, cutoff, maxi) if use_cut else oracle.kbest2d(k, C, maxi)
        assert_kbest_equal(got, (n, r, c, g), f"{name}[{i}]")


def test_appendix_a_known_answer(oracle):
    # SURVEY.md appendix A, typed in by hand (not read from a file)
    inf = np.inf
    C = np.array([[1, 2], [4, 3], [6, 7], [10, inf], [inf, 10]], dtype=np.float64)
    n, r4c, c4r, g = oracle.kbest2d_cutoff(30, C, 42.0)
    assert n == 13
    assert g[:13].tolist() == [4, 6, 8, 8, 9, 11, 11, 12, 13, 14, 16, 17, 20]
    assert r4c[:13].tolist() == [[0, 1], [1, 0], [0, 2], [2, 0], [2, 1], [1, 2], [0, 4], [3, 0], [3, 1], [1, 4], [2, 4], [3, 2], [3, 4]]
    assert c4r[6].tolist() == [0, 2, 4, 3, 1] and c4r[12].tolist() == [2, 3, 4, 0, 1]
    p = oracle.assignment_prob(C, 3, 30)
    np.testing.assert_allclose(p[0], [0.8629907580527344, 0.11540036124767361, 0.02121833941220368, 0.00039054128738825371], rtol=1e-14)
    np.testing.assert_allclose(p[1], [0.13038190608964614, 0.85252019571967197, 0.016282059796615539, 0.00081583839406631196], rtol=1e-14)


def test_sticky_scratchspace(oracle):
    z = golden("kbest_sticky_k120")
    for i in range(int(z["n"])):
        C = z[f"C{i}"]
        got = oracle.kbest2d_after_cutoff(int(z["k"]), C, False, C, False, float(z["first_cutoff"]))
        assert_kbest_equal(got, (int(z[f"n{i}"]), z[f"r{i}"].astype(np.int64), z[f"c{i}"].astype(np.int64), z[f"g{i}"]), f"sticky[{i}]")


def test_lap_with_duals(oracle):
    z = golden("lap")
    for i in range(int(z["n"])):
        ret, r4c, c4r, u, v, g = oracle.assign2d(z[f"C{i}"])
        assert ret == int(z[f"ret{i}"])
        np.testing.assert_array_equal(r4c, z[f"r{i}"]); np.testing.assert_array_equal(c4r, z[f"c{i}"])
        np.testing.assert_array_equal(bits(u), bits(z[f"u{i}"])); np.testing.assert_array_equal(bits(v), bits(z[f"v{i}"]))
        assert bits([g])[0] == bits([float(z[f"g{i}"])])[0]
        ret, r4c, c4r, u, v, g, fb = oracle.shortest_path(z[f"S{i}"])
        assert ret == int(z[f"sret{i}"])
        np.testing.assert_array_equal(r4c, z[f"sr{i}"]); np.testing.assert_array_equal(c4r, z[f"sc{i}"])
        np.testing.assert_array_equal(bits(u), bits(z[f"su{i}"])); np.testing.assert_array_equal(bits(v), bits(z[f"sv{i}"]))
        np.testing.assert_array_equal(fb, z[f"sf{i}"])


def test_condition_costs(oracle):
    z = golden("condition_g2")
    for p in range(int(z["n"])):
        out, idx = oracle.condition_costs(z[f"in{p}"])
        np.testing.assert_array_equal(bits(out), bits(z[f"out{p}"]))
        np.testing.assert_array_equal(idx, z[f"idx{p}"])


def test_weights(oracle):
    # exp() comes from the host libm in both the reference and the oracle: allow last-bit drift only
    z = golden("weights_g2cond")
    for p in range(int(z["n"])):
        C, nL = z[f"C{p}"], int(z[f"nL{p}"])
        np.testing.assert_allclose(oracle.assignment_prob(C, nL, 200), z[f"k200_{p}"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(oracle.assignment_prob(C, nL, 20), z[f"k20_{p}"], rtol=1e-13, atol=0)
        if f"bf_{p}" in z:
            np.testing.assert_allclose(oracle.brute_force_prob(C, nL), z[f"bf_{p}"], rtol=1e-13, atol=0)
            st, pp = oracle.permanent_prob(C, nL, 1)
            assert st == 0
            np.testing.assert_allclose(pp, z[f"pp_{p}"], rtol=1e-13, atol=0)
    z = golden("weights_g1")
    for p in range(int(z["n"])):
        np.testing.assert_allclose(oracle.assignment_prob(z[f"C{p}"], 30, 200), z[f"k200_{p}"], rtol=1e-13, atol=0)
    z = golden("probs_config1")
    for k in (1, 20, 100, 200, 1000):
        np.testing.assert_allclose(oracle.assignment_prob(z["C"], 30, k), z[f"k{k}"], rtol=1e-13, atol=0)
    z = golden("toprobs")
    for i in range(int(z["n"])):
        np.testing.assert_allclose(oracle.to_probs(z[f"in{i}"]), z[f"out{i}"], rtol=1e-14, atol=0)


def test_association_probs(oracle):
    z = golden("association_g2")
    for p in range(int(z["n"])):
        st, pr = oracle.association_probs(z[f"C{p}"], 30, 200, False)
        assert st == 0
        np.testing.assert_allclose(pr, z[f"k200_{p}"], rtol=1e-13, atol=0)
        if f"perm_{p}" in z:
            st, pr = oracle.association_probs(z[f"C{p}"], 30, 200, True)
            assert st == 0
            np.testing.assert_allclose(pr, z[f"perm_{p}"], rtol=1e-13, atol=0)
    # no landmarks: every detection is "not assigned" with probability one (assignment.cpp:51-53)
    st, pr = oracle.association_probs(np.array([[10.0, np.inf], [np.inf, 10.0]]), 0, 5)
    assert st == 0 and pr.tolist() == [[1.0], [1.0]]


def test_stereo_box_association(oracle):
    z = golden("asgn_bb")
    for i in range(int(z["n"])):
        np.testing.assert_array_equal(oracle.asgn_bb(z[f"L{i}"], z[f"R{i}"], float(z["nonassign"])), z[f"a{i}"])
        if f"C{i}" in z:
            np.testing.assert_array_equal(bits(np.nan_to_num(oracle.bb_cost_matrix(z[f"L{i}"], z[f"R{i}"], float(z["nonassign"])), neginf=-1.0)),
                                          bits(np.nan_to_num(z[f"C{i}"], neginf=-1.0)))


def test_permanents(oracle):
    z = golden("permanent")
    for n in z["dims"]:
        for i in range(2):
            val, st = oracle.permanent_exact_square(z[f"A_{n}_{i}"])
            assert st == 0 and bits([val])[0] == bits([float(z[f"p_{n}_{i}"])])[0], (n, i)
    for r, c in z["rect"]:
        val, st = oracle.permanent_exact(z[f"R_{r}_{c}"])
        assert st == 0 and bits([val])[0] == bits([float(z[f"rp_{r}_{c}"])])[0], (r, c)
    assert oracle.permanent_exact_square(np.ones((33, 33)))[1] == 1  # the reference throws above 32
    assert oracle.permanent_exact_square(np.array([[1.0, 2.0], [3.0, 4.0]]))[0] == 10.0
    assert oracle.permanent_exact(np.ones((2, 3)))[0] == 6.0
    z = golden("conditioned_permanent")
    for i in range(int(z["n"])):
        val, st = oracle.conditioned_permanent(z[f"A{i}"], 1)
        assert st == int(z[f"s{i}"])
        np.testing.assert_allclose(val, float(z[f"v{i}"]), rtol=1e-14)


def test_permanent_small_vs_definition(oracle):
    # independent of any file: permanent by brute-force enumeration of permutations
    import itertools
    rng = np.random.default_rng(5)
    for n in range(1, 8):
        A = rng.random((n, n))
        want = sum(np.prod([A[i, p[i]] for i in range(n)]) for p in itertools.permutations(range(n)))
        np.testing.assert_allclose(oracle.permanent_exact_square(A)[0], want, rtol=1e-12)


def test_kbest_marginals_approach_brute_force(oracle):
    # the compMethods relation (comparison.cpp:261-275, 319-324): k-best marginals ~ brute-force marginals
    z = golden("weights_g2cond")
    for p in range(int(z["n"])):
        if f"bf_{p}" in z:
            assert np.max(np.abs(z[f"k200_{p}"] - z[f"bf_{p}"])) < 0.1
            assert np.max(np.abs(z[f"pp_{p}"] - z[f"bf_{p}"])) < 1e-6


def test_kbest_is_the_sorted_list_of_all_assignments(oracle):
    """Independent of the reference: on tiny problems the k-best list must be exactly the cheapest feasible assignments
    (every detection to a distinct landmark or to its own missed-detection row), in non-decreasing order of cost, each
    listed once, with gain == the sum of the chosen entries."""
    import itertools
    rng = np.random.default_rng(77)
    for trial in range(60):
        nL, nM = int(rng.integers(0, 5)), int(rng.integers(1, 4))
        C = np.full((nL + nM, nM), np.inf)
        C[:nL, :] = np.where(rng.random((nL, nM)) < 0.8, rng.uniform(0, 20, size=(nL, nM)), np.inf)
        C[nL + np.arange(nM), np.arange(nM)] = 10.0
        every = []
        for rows in itertools.permutations(range(nL + nM), nM):
            cost = sum(C[r, c] for c, r in enumerate(rows))
            if np.isfinite(cost):
                every.append((cost, rows))
        every.sort(key=lambda t: t[0])
        k = 40
        n, r4c, c4r, g = oracle.kbest2d(k, C)
        assert n == min(k, len(every)), (trial, n, len(every))
        np.testing.assert_allclose(g[:n], [e[0] for e in every[:n]], rtol=1e-12, atol=1e-12)
        seen = set()
        for i in range(n):
            rows = tuple(int(x) for x in r4c[i])
            assert rows not in seen
            seen.add(rows)
            assert abs(sum(C[r, c] for c, r in enumerate(rows)) - g[i]) <= 1e-12 * max(1.0, abs(g[i]))
            assert all(c4r[i][r] == c for c, r in enumerate(rows))
