"""CPU-only: the N>1 host logic (problem sharding, the one gather, Gray-range split + partial-sum
exchange) on world_size 2 and 3 over gloo.  The compute on each rank is a stand-in (the CPU oracle, or a
numpy Gray-range walk) -- what is under test is the partitioning and the collectives."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from probabilisticsemslam_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nw_partial(a, begin, end):
    """Plain-python NW partial sum over Gray indices [begin, end): (hi, lo) with lo = 0."""
    n = a.shape[0]
    base = a[:, n - 1] - a.sum(axis=1) / 2
    tot = 0.0
    for i in range(begin, end):
        g = i ^ (i >> 1)
        x = base.copy()
        for b in range(n - 1):
            if (g >> b) & 1:
                x += a[:, b]
        tot += (-1.0 if i & 1 else 1.0) * float(np.prod(x))
    return tot, 0.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.loader import load_oracle
        orc = load_oracle()
        pb = synth.g1_dense(23, nL=8, first=5)          # 23 problems: uneven split
        cpu = lambda part, k: orc.batch(part, k, threads=1, want_probs=True, want_lists=False)["probs"]  # noqa: E731
        probs = shard.sharded_assignment_prob(pb, 40, compute=cpu)
        a = synth.dense_square(1, 9, first=3)[0].reshape(9, 9, order="F")
        perm = shard.sharded_permanent(a, partial=_nw_partial)
        frames = synth.quadric_frames(7, nL=6, first=40)     # 7 frames: uneven split, ragged tables
        cpu_frames = lambda fr, na, k: [orc.association_from_moments(*f, na, k) for f in fr]  # noqa: E731
        tabs = shard.sharded_association_from_moments(frames, 10.0, 30, compute=cpu_frames)
        if rank == 0:
            np.save(out + ".probs.npy", probs)
            np.save(out + ".perm.npy", np.array([perm]))
            np.save(out + ".tabs.npy", np.concatenate([t.reshape(-1) for t in tabs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_match_single_process(tmp_path, oracle, world):
    out = str(tmp_path / f"w{world}")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    pb = synth.g1_dense(23, nL=8, first=5)
    want = oracle.batch(pb, 40, threads=1, want_probs=True, want_lists=False)["probs"]
    np.testing.assert_array_equal(np.load(out + ".probs.npy"), want)      # same per-problem code, only re-assembled
    a = synth.dense_square(1, 9, first=3)[0].reshape(9, 9, order="F")
    np.testing.assert_allclose(np.load(out + ".perm.npy")[0], oracle.permanent_exact_square(a)[0], rtol=1e-12)
    frames = synth.quadric_frames(7, nL=6, first=40)
    want_tabs = np.concatenate([oracle.association_from_moments(*f, 10.0, 30).reshape(-1) for f in frames])
    np.testing.assert_array_equal(np.load(out + ".tabs.npy"), want_tabs)


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 100_000):
        for world in (1, 2, 3, 8):
            cuts = [shard.shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_gray_ranges_are_aligned_and_disjoint():
    for n in (1, 2, 9, 24, 28):
        for world in (1, 2, 3, 4, 8):
            total = 1 << (n - 1)
            cuts = [shard.gray_range(n, world, r) for r in range(world)]
            covered = sum(hi - lo for lo, hi in cuts)
            assert covered == total
            nonempty = [c for c in cuts if c[1] > c[0]]
            assert nonempty[0][0] == 0 and nonempty[-1][1] == total
            size = nonempty[0][1] - nonempty[0][0]
            assert all(hi - lo == size and lo % size == 0 for lo, hi in nonempty)   # aligned power-of-two pieces


def test_combine_partials_is_exact_for_cancelling_terms():
    parts = np.array([[1e30, 1.0], [-1e30, 2.0], [3.0, 0.0]])
    assert shard.combine_partials(parts, 2) == -2.0 * 6.0
