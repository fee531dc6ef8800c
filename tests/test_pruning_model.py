"""CPU: the pruning rules of the CUDA fast path (murty_kernel<R, true>) are sound.

oracle/oracle_murty.c carries a CPU model of the DECISIONS that kernel takes -- the bound T on the k-th gain, when and how
it is tightened (32-bucket histogram, re-checked count), which child searches are abandoned (distance beyond T - parent
gain, 1e-7 relative margin) or dropped when finished, and when a selection counts as tied and the problem is handed to
the exact kernel -- on top of the reference arithmetic.  Whenever the model does not bail out, its k-best lists, their
order and the gains must be bit-identical to the plain enumeration; on tie-heavy input it must bail out rather than
guess.  The same counters come out of the kernel itself (-DPDA_FAST_STATS build, scripts/fast_path_probe.py): on the first
20 000 G1 problems at k = 200 both give 404.31 children, 133.58 abandoned, 270.73 kept and 9.01 tightenings per problem."""
import numpy as np
import pytest

from probabilisticsemslam_b200 import synth


def _same(a, b, tag):
    n = a[0]
    assert n == b[0], f"{tag}: nFound {a[0]} vs {b[0]}"
    assert np.array_equal(a[1][:n], b[1][:n]), f"{tag}: row4col"
    assert np.array_equal(a[2][:n], b[2][:n]), f"{tag}: col4row"
    assert np.array_equal(a[3][:n].view(np.int64), b[3][:n].view(np.int64)), f"{tag}: gains"


def test_pruned_enumeration_equals_exact_on_continuous_costs(oracle):
    pb = synth.g1_dense(250, first=0)
    tot = dict(children=0, abandoned=0, kept=0, tightenings=0)
    for p in range(len(pb)):
        C = pb.matrix(p)
        got = oracle.kbest2d_cutoff_pruned(200, C, 42.0, max_col=8)
        assert got[0] != -2, f"problem {p}: continuous costs must not tie"
        _same(got, oracle.kbest2d_cutoff(200, C, 42.0), f"G1[{p}]")
        for key in tot:
            tot[key] += got[4][key]
    n = len(pb)
    # the shape of the saving: about a third of all children is abandoned, ~9 tightenings per problem
    assert 0.28 < tot["abandoned"] / tot["children"] < 0.38
    assert 7 < tot["tightenings"] / n < 12
    assert tot["kept"] < 0.72 * tot["children"]


def test_pruned_enumeration_bails_out_on_ties(oracle):
    """Integer costs: ~96 % of neighbouring hypotheses have equal gains.  The model -- like the kernel -- must
    either reproduce the reference's order exactly or decline; it must never emit a different list."""
    pb = synth.g1_dense(120, first=300, integer=True)
    bailed = 0
    for p in range(len(pb)):
        C = pb.matrix(p)
        got = oracle.kbest2d_cutoff_pruned(150, C, 42.0, max_col=8)
        if got[0] == -2:
            bailed += 1
        else:
            _same(got, oracle.kbest2d_cutoff(150, C, 42.0), f"G1-int[{p}]")
    assert bailed > 100


@pytest.mark.parametrize("k", [2, 3, 17, 1000])
def test_pruned_enumeration_other_k(oracle, k):
    pb = synth.g1_dense(12, first=5000 + k)
    for p in range(len(pb)):
        C = pb.matrix(p)
        got = oracle.kbest2d_cutoff_pruned(k, C, 42.0, max_col=8)
        assert got[0] != -2
        _same(got, oracle.kbest2d_cutoff(k, C, 42.0), f"k={k}[{p}]")


def test_pruned_enumeration_on_gated_and_small_problems(oracle):
    """Conditioned (gated) KITTI-like problems: 5-23 rows, +inf entries, often fewer than k feasible hypotheses, early
    stop by the cutoff; and a tight cutoff that ends the enumeration long before k."""
    g2 = synth.g2_gated(150, first=8000)
    done = 0
    for p in range(len(g2)):
        cond, _ = oracle.condition_costs(g2.matrix(p))
        if cond.shape[1] < 2:
            continue
        got = oracle.kbest2d_cutoff_pruned(200, cond, 42.0, max_col=8)
        if got[0] != -2:
            _same(got, oracle.kbest2d_cutoff(200, cond, 42.0), f"G2[{p}]")
            done += 1
    assert done > 100
    pb = synth.g1_dense(30, first=777)
    for p in range(len(pb)):
        for cutoff in (0.5, 3.0):
            got = oracle.kbest2d_cutoff_pruned(200, pb.matrix(p), cutoff, max_col=8)
            assert got[0] != -2
            _same(got, oracle.kbest2d_cutoff(200, pb.matrix(p), cutoff), f"cutoff {cutoff}[{p}]")


def test_pruned_enumeration_maximize(oracle):
    rng = np.random.default_rng(5)
    for i in range(20):
        C = rng.uniform(0.0, 1.0, size=(9, 4))
        got = oracle.kbest2d_cutoff_pruned(60, C, 42.0, maximize=True, max_col=4)
        assert got[0] != -2
        _same(got, oracle.kbest2d_cutoff(60, C, 42.0, maximize=True), f"max[{i}]")
