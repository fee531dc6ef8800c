"""GPU parity: the CUDA k-best path (through the C ABI) against the reference-generated golden
vectors and against the CPU oracle on larger seeded batches.  Index lists, enumeration order and
gains must be bit-identical."""
import numpy as np
import pytest

from helpers import KBEST_FILES, assert_kbest_equal, bits, golden, kbest_cases
from probabilisticsemslam_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("murty_path")]


def _run_cases(api, mats, k, use_cut, cutoff, maxi):
    pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
    res = api.murty_batch(pb, k, cut_mode=api.CUT_RELATIVE if use_cut else api.CUT_NONE, cutoff=cutoff, maximize=maxi)
    return pb, res


@pytest.mark.parametrize("name", KBEST_FILES)
def test_kbest_golden_bit_exact(gpu_api, name):
    z = golden(name)
    k, cutoff, use_cut, maxi = int(z["k"]), float(z["cutoff"]), bool(z["use_cutoff"]), bool(z["maximize"])
    cases = list(kbest_cases(z))
    pb, res = _run_cases(gpu_api, [c[0] for c in cases], k, use_cut, cutoff, maxi)
    for i, (C, n, r, c, g) in enumerate(cases):
        r4c, c4r, gain = res.lists(pb, i)
        assert_kbest_equal((int(res.n_found[i]), r4c, c4r, gain), (n, r, c, g), f"{name}[{i}]")


def test_single_problem_functions(gpu_api):
    """The reference's call signatures, batch of one (config 1)."""
    z = golden("kbest_config1")
    C, n, r, c, g = next(kbest_cases(z))
    got = gpu_api.kBest2DCutoff(200, C, 42.0)
    assert_kbest_equal(got, (n, r, c, g), "kBest2DCutoff")
    z = golden("kbest_g1int_nocut_k300")
    C, n, r, c, g = next(kbest_cases(z))
    assert_kbest_equal(gpu_api.kBest2D(300, C), (n, r, c, g), "kBest2D")


def test_sticky_scratchspace(gpu_api, oracle):
    z = golden("kbest_sticky_k120")
    for i in range(int(z["n"])):
        C = z[f"C{i}"]
        # the first call (k = 1, cutoff 6) leaves cutoffGain = shifted root gain + 6 behind (shortestPathCPP.cpp:681)
        shifted = C - C.min()
        ret, r4c, c4r, u, v, g0, fb = oracle.shortest_path(np.hstack([shifted, np.zeros((C.shape[0], C.shape[0] - C.shape[1]))]), C.shape[1])
        got = gpu_api.kBest2D_sticky(int(z["k"]), C, False, g0 + float(z["first_cutoff"]), False)
        assert_kbest_equal(got, (int(z[f"n{i}"]), z[f"r{i}"].astype(np.int64), z[f"c{i}"].astype(np.int64), z[f"g{i}"]), f"sticky[{i}]")


def test_lap_with_duals(gpu_api):
    z = golden("lap")
    for i in range(int(z["n"])):
        ret, r4c, c4r, u, v, g = gpu_api.assign2D(z[f"C{i}"])
        assert ret == int(z[f"ret{i}"])
        np.testing.assert_array_equal(r4c, z[f"r{i}"]); np.testing.assert_array_equal(c4r, z[f"c{i}"])
        np.testing.assert_array_equal(bits(u), bits(z[f"u{i}"])); np.testing.assert_array_equal(bits(v), bits(z[f"v{i}"]))
        assert bits([g])[0] == bits([float(z[f"g{i}"])])[0]
        ret, r4c, c4r, u, v, g, fb = gpu_api.shortestPathCPP(z[f"S{i}"])
        assert ret == int(z[f"sret{i}"])
        np.testing.assert_array_equal(r4c, z[f"sr{i}"]); np.testing.assert_array_equal(c4r, z[f"sc{i}"])
        np.testing.assert_array_equal(bits(u), bits(z[f"su{i}"])); np.testing.assert_array_equal(bits(v), bits(z[f"sv{i}"]))
        np.testing.assert_array_equal(fb, z[f"sf{i}"])
    # infeasible: a column of +inf
    C = z["C0"].copy(); C[:, 0] = np.inf
    assert gpu_api.assign2D(C)[0] == 0


def _compare_batch(api, oracle, pb, k, **kw):
    res = api.murty_batch(pb, k, **kw)
    cut = kw.get("cut_mode", api.CUT_RELATIVE) == api.CUT_RELATIVE
    want = oracle.batch(pb, k, cutoff=kw.get("cutoff", 42.0), threads=8, want_probs=False, want_lists=True) if cut else None
    bad = []
    for p in range(len(pb)):
        r4c, c4r, gain = res.lists(pb, p)
        if cut:
            n = int(want["n_found"][p]); nc = int(pb.nM[p]); nr = nc + int(pb.nL[p])
            wr = want["row4col"][want["r4c_off"][p]:want["r4c_off"][p] + n * nc].reshape(n, nc)
            wc = want["col4row"][want["c4r_off"][p]:want["c4r_off"][p] + n * nr].reshape(n, nr)
            wg = want["gain"][p * k:p * k + n]
        else:
            n, wr, wc, wg = oracle.kbest2d(k, pb.matrix(p), kw.get("maximize", False))
        try:
            assert_kbest_equal((int(res.n_found[p]), r4c, c4r, gain), (n, wr, wc, wg), f"problem {p}")
        except AssertionError as e:
            bad.append(str(e)[:200])
    assert not bad, f"{len(bad)} of {len(pb)} problems differ; first: {bad[0]}"


def test_g1_batch_vs_oracle(gpu_api, oracle):
    """A 1 500-problem slice of config 2 (k = 200), every hypothesis compared."""
    _compare_batch(gpu_api, oracle, synth.g1_dense(1500, first=1000), 200)


def test_g1int_tie_stress_vs_oracle(gpu_api, oracle):
    """Integer costs: ~96 % of neighbouring hypotheses have exactly equal gains, so this pins the
    first-minimum tie-break of the row scan and the libstdc++ heap mechanics."""
    _compare_batch(gpu_api, oracle, synth.g1_dense(600, first=5000, integer=True), 200)


def test_g1_k1000_vs_oracle(gpu_api, oracle):
    """Slice of config 3 (k = 1000): deeper heaps, larger node arenas."""
    _compare_batch(gpu_api, oracle, synth.g1_dense(120, first=9000), 1000)
    _compare_batch(gpu_api, oracle, synth.g1_dense(60, first=9500, integer=True), 1000)


def test_ragged_shapes_vs_oracle(gpu_api, oracle):
    """Everything from 1x1 to 60x9 in one batch, incl. square problems and a single detection."""
    mats = []
    rng = np.random.default_rng(11)
    for nL, nM in [(0, 1), (0, 3), (1, 1), (2, 2), (5, 1), (7, 7), (12, 4), (31, 1), (30, 2), (33, 3), (51, 9), (60, 4), (20, 12)]:
        C = np.full((nL + nM, nM), np.inf)
        C[:nL, :] = np.floor(rng.random((nL, nM)) * 30) / 2
        C[nL + np.arange(nM), np.arange(nM)] = 10.0
        mats.append(C)
    pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
    _compare_batch(gpu_api, oracle, pb, 64)
    _compare_batch(gpu_api, oracle, pb, 64, cut_mode=gpu_api.CUT_NONE)


def test_large_dimension_vs_oracle(gpu_api, oracle):
    """numRow up to 128 exercises the 4-rows-per-lane instantiation."""
    mats = []
    for i, (nL, nM) in enumerate([(70, 5), (100, 8), (118, 10), (90, 3)]):
        mats.append(synth.g1_dense(1, nL=nL, nM=nM, first=300 + i).matrix(0))
    pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
    _compare_batch(gpu_api, oracle, pb, 40)


def test_rounding_stress_vs_oracle(gpu_api, oracle):
    """Costs spread over twelve decades and costs that are exact multiples of 2^-20: the row scan's fast-forward
    over no-op padding hops must reproduce every last-bit effect of stepping through them one by one."""
    rng = np.random.default_rng(2024)
    mats = []
    for i in range(120):
        nL, nM = int(rng.integers(5, 40)), int(rng.integers(2, 9))
        C = np.full((nL + nM, nM), np.inf)
        if i % 3 == 0:
            C[:nL, :] = 10.0 ** rng.uniform(-6, 6, size=(nL, nM))
        elif i % 3 == 1:
            C[:nL, :] = np.round(rng.uniform(0, 30, size=(nL, nM)) * 2 ** 20) / 2 ** 20
        else:
            C[:nL, :] = rng.uniform(0, 40, size=(nL, nM)) * (rng.random((nL, nM)) < 0.6) + 1e-9 * rng.random((nL, nM))
        C[nL + np.arange(nM), np.arange(nM)] = 10.0 if i % 2 else 1e-3
        mats.append(C)
    pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
    _compare_batch(gpu_api, oracle, pb, 150)
    _compare_batch(gpu_api, oracle, pb, 150, cut_mode=gpu_api.CUT_NONE)


def test_maximize_vs_oracle(gpu_api, oracle):
    mats = []
    for p in range(20):
        m = synth.g1_dense(1, nM=3 + p % 4, nL=5 + p, first=700 + p).matrix(0)
        s = np.where(np.isfinite(m), m / 40.0, -np.inf)
        for c in range(s.shape[1]):
            s[5 + p + c, c] = 0.6
        mats.append(s)
    pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
    _compare_batch(gpu_api, oracle, pb, 30, cut_mode=gpu_api.CUT_NONE, maximize=True)


def test_small_workspace_still_correct(gpu_api, oracle):
    """A workspace with room for a single arena must give the same answers (one problem in flight)."""
    import ctypes as C
    import torch
    from probabilisticsemslam_b200 import device as dev
    pb = synth.g1_dense(40, first=123)
    plan = dev.MurtyPlan(pb, k=50, max_arenas=1)
    plan.run()
    torch.cuda.synchronize()
    res = plan.result()
    want = oracle.batch(pb, 50, threads=4, want_probs=False)
    np.testing.assert_array_equal(res.n_found, want["n_found"])
    for p in range(len(pb)):
        r4c, c4r, g = res.lists(pb, p)
        n = int(want["n_found"][p]); nc = int(pb.nM[p])
        np.testing.assert_array_equal(r4c, want["row4col"][want["r4c_off"][p]:want["r4c_off"][p] + n * nc].reshape(n, nc))


def test_full_size_properties(gpu_api):
    """100 000 KITTI-shaped problems at k = 200 (BASELINE.json configs[1]) through the device API;
    size-independent properties checked on the device, no oracle involved:
      gains non-decreasing; every hypothesis an injective map; its gain equals the sum of the selected
      costs; row4col and col4row are mutually inverse; hypotheses pairwise distinct (via a checksum);
      weights sum to one per detection."""
    import torch
    from probabilisticsemslam_b200 import device as dev
    n, k = 100_000, 200
    pb = synth.g1_dense(n)
    plan = dev.MurtyPlan(pb, k=k, weights=True)
    plan.run()
    torch.cuda.synchronize()
    nf = plan.n_found.cpu().numpy()
    assert nf.min() >= 1 and nf.max() <= k
    gain = plan.gain.view(n, k)
    valid = torch.arange(k, device="cuda")[None, :] < plan.n_found[:, None]
    d = gain[:, 1:] - gain[:, :-1]
    assert bool(((d >= 0) | ~valid[:, 1:]).all()), "gains must be non-decreasing"
    costs = plan.costs
    for m in range(3, 9):
        sel = torch.nonzero(plan.num_col == m).flatten()
        if sel.numel() == 0:
            continue
        nr = 30 + m
        r4c = torch.stack([plan.row4col[int(plan.r4c_off_h[p]):int(plan.r4c_off_h[p]) + k * m].view(k, m) for p in sel[:200].tolist()])
        c4r = torch.stack([plan.col4row[int(plan.c4r_off_h[p]):int(plan.c4r_off_h[p]) + k * nr].view(k, nr) for p in sel[:200].tolist()])
        v = valid[sel[:200]]
        assert bool((((r4c >= 0) & (r4c < nr)) | ~v[..., None]).all())
        # injective: sorted rows strictly increasing
        s, _ = torch.sort(r4c, dim=2)
        assert bool(((s[..., 1:] > s[..., :-1]) | ~v[..., None]).all())
        # inverse maps
        back = torch.gather(c4r, 2, r4c.clamp(min=0))
        assert bool(((back == torch.arange(m, device="cuda")) | ~v[..., None]).all())
        # gain == sum of selected costs (same summation order as calcGain; compare to 1e-9 relative)
        C = torch.stack([costs[int(plan.cost_off_h[p]):int(plan.cost_off_h[p]) + nr * m].view(m, nr) for p in sel[:200].tolist()])
        picked = torch.gather(C[:, None].expand(-1, k, -1, -1), 3, r4c.clamp(min=0)[..., None]).squeeze(3)
        tot = picked.sum(dim=2)
        g = gain[sel[:200]]
        assert bool((((tot - g).abs() <= 1e-9 * g.abs().clamp(min=1.0)) | ~v).all())
        # distinct hypotheses: rows < 64 and m <= 8, so sum(row * 64^col) encodes the map injectively
        w = 64 ** torch.arange(m, device="cuda", dtype=torch.int64)
        code = (r4c * w).sum(dim=2)
        code = torch.where(v, code, -torch.arange(k, device="cuda")[None, :] - 1)
        srt, _ = torch.sort(code, dim=1)
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "a hypothesis was enumerated twice"
    # weights: rows of every table sum to one
    probs = plan.probs
    for m in range(3, 9):
        sel = torch.nonzero(plan.num_col == m).flatten()[:500].tolist()
        for p in sel[:50]:
            t = probs[int(plan.prob_off_h[p]):int(plan.prob_off_h[p]) + m * 31].view(m, 31)
            assert bool(((t.sum(dim=1) - 1.0).abs() < 1e-12).all())


def test_c_abi_error_paths(gpu_api):
    """Bad arguments come back as negative status codes with a message, never as a crash or a silent fallback."""
    import ctypes as C
    from probabilisticsemslam_b200 import _lib
    lib = _lib.lib()
    pb = synth.g1_dense(4)
    nr, nc = pb.num_row, pb.nM.astype(np.int32)
    found = np.zeros(4, np.int32)
    p = lambda a: a.ctypes.data
    # numRow < numCol
    bad_nr = nc.copy() - 1
    rc = lib.pda_murty_batch_host(p(pb.costs), p(pb.cost_off), p(bad_nr), p(nc), 4, 5, 0, 0.0, 0, 0, None, None, None, None, None,
                                  p(found), 0, None, None, None, 0)
    assert rc == -1 and b"numRow" in lib.pda_last_error()
    # k < 1
    rc = lib.pda_murty_batch_host(p(pb.costs), p(pb.cost_off), p(nr), p(nc), 4, 0, 0, 0.0, 0, 0, None, None, None, None, None,
                                  p(found), 0, None, None, None, 0)
    assert rc == -1
    # weights requested without outputs
    rc = lib.pda_murty_batch_host(p(pb.costs), p(pb.cost_off), p(nr), p(nc), 4, 5, 1, 42.0, 0, 0, None, None, None, None, None,
                                  p(found), 1, None, None, None, 0)
    assert rc == -1
    # dimension above PDA_MAX_DIM
    big = synth.pack([np.zeros((200, 3))], [197])
    big_nr, big_nc = big.num_row, big.nM.astype(np.int32)  # keep the arrays alive across the call
    rc = lib.pda_murty_batch_host(p(big.costs), p(big.cost_off), p(big_nr), p(big_nc), 1, 5, 0, 0.0, 0, 0,
                                  None, None, None, None, None, p(found), 0, None, None, None, 0)
    assert rc == -3
    # device out of range
    rc = lib.pda_murty_batch_host(p(pb.costs), p(pb.cost_off), p(nr), p(nc), 4, 5, 0, 0.0, 0, 0, None, None, None, None, None,
                                  p(found), 0, None, None, None, 99)
    assert rc == -1
    # a workspace too small for a single arena
    import torch
    d = lambda t: t.data_ptr()
    costs = torch.from_numpy(pb.costs).cuda(); off = torch.from_numpy(pb.cost_off).cuda()
    tnr = torch.from_numpy(nr).cuda(); tnc = torch.from_numpy(nc).cuda(); tf = torch.zeros(4, dtype=torch.int32, device="cuda")
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = lib.pda_murty_batch(d(costs), d(off), d(tnr), d(tnc), 4, 38, 8, 50, 0, 0.0, 0, 0, None, None, None, None, None, d(tf),
                             0, None, None, None, d(ws), 1024, None)
    assert rc == -4 and b"workspace" in lib.pda_last_error()
    with pytest.raises(_lib.PdaError):
        gpu_api.kBest2D(5, np.zeros((2, 3)))          # more columns than rows


def test_fuzz_shapes_cutmodes_vs_oracle(gpu_api, oracle):
    """Seeded fuzz over everything the two kernels branch on: 1x1 ... 24x9 problems, sparse +inf patterns (including
    infeasible ones), exact-tie grids, tiny and huge k, relative cut-offs from 0 to 60, minimise and maximise."""
    rng = np.random.default_rng(20260217)
    for trial in range(12):
        mats = []
        for _ in range(160):
            nM = int(rng.integers(1, 10))
            nL = int(rng.integers(0, 16))
            C = np.full((nL + nM, nM), np.inf)
            kind = int(rng.integers(0, 4))
            if kind == 0:
                land = rng.uniform(0, 40, size=(nL, nM))
            elif kind == 1:
                land = np.floor(rng.uniform(0, 6, size=(nL, nM)))                      # many exact ties
            elif kind == 2:
                land = np.where(rng.random((nL, nM)) < 0.5, np.inf, rng.uniform(0, 30, size=(nL, nM)))  # gated entries
            else:
                land = rng.uniform(0, 1, size=(nL, nM)) * 10.0 ** rng.integers(-3, 4)
            C[:nL, :] = land
            dummy = rng.random(nM) < 0.85                                              # some detections cannot be missed
            C[nL + np.arange(nM), np.arange(nM)] = np.where(dummy, float(rng.choice([10.0, 3.0, 0.5])), np.inf)
            mats.append(C)
        pb = synth.pack(mats, [m.shape[0] - m.shape[1] for m in mats])
        k = int(rng.choice([1, 2, 7, 33, 150]))
        mode = trial % 3
        if mode == 0:
            _compare_batch(gpu_api, oracle, pb, k, cutoff=float(rng.choice([0.0, 1.5, 42.0, 60.0])))
        elif mode == 1:
            _compare_batch(gpu_api, oracle, pb, k, cut_mode=gpu_api.CUT_NONE)
        else:
            neg = synth.pack([np.where(np.isfinite(m), -m, -np.inf) for m in mats], [m.shape[0] - m.shape[1] for m in mats])
            _compare_batch(gpu_api, oracle, neg, k, cut_mode=gpu_api.CUT_NONE, maximize=True)


def test_full_size_every_hypothesis_vs_reference(gpu_api, oracle):
    """BASELINE.json configs[1] in full: all 100 000 problems, all 200 hypotheses each (2e7 index lists and gains), bit
    for bit against the reference's own code compiled IEEE-strict when oracle/_ref travelled to this box, else against
    the C restatement; weights to 1e-9.  ~15 s per kernel (the CPU side runs on all host threads)."""
    import os
    from oracle.loader import load_reference, reference_available
    chk = load_reference("strict") if reference_available("strict") else oracle
    n_total, k, chunk = 100_000, 200, 10_000
    for first in range(0, n_total, chunk):
        pb = synth.g1_dense(chunk, first=first)
        got = gpu_api.murty_batch(pb, k, weight_mode=gpu_api.WEIGHTS_GATED)
        want = chk.batch(pb, k, threads=os.cpu_count(), want_probs=True, want_lists=True)
        np.testing.assert_array_equal(got.n_found, want["n_found"])
        assert np.array_equal(got.row4col, want["row4col"]), f"row4col differs in problems {first}..{first + chunk}"
        assert np.array_equal(got.col4row, want["col4row"]), f"col4row differs in problems {first}..{first + chunk}"
        valid = (np.arange(k)[None, :] < want["n_found"][:, None]).reshape(-1)
        assert not np.any((got.gain.reshape(-1).view(np.int64) != want["gain"].view(np.int64)) & valid), "gain bits differ"
        np.testing.assert_allclose(got.probs, want["probs"], rtol=1e-9, atol=0)


def test_malformed_problem_is_reported_not_solved(gpu_api):
    """Device-pointer entry: a problem whose dimensions are malformed, or larger than the maxima the call declared, is
    reported as nFound = -1 (0 would mean "infeasible") and does not disturb its neighbours."""
    import torch
    from probabilisticsemslam_b200 import device as dev
    pb = synth.g1_dense(6, first=4321)
    good = dev.MurtyPlan(pb, k=10, weights=True)
    good.run()
    torch.cuda.synchronize()
    want = good.n_found.cpu().numpy().copy()
    want_probs = good.probs.cpu().numpy().copy()
    bad = dev.MurtyPlan(pb, k=10, weights=True)
    bad.num_col[2] = 0                                   # no detections
    bad.num_row[4] = int(bad.max_row) + 3                # more rows than the launch was sized for
    bad.run()
    torch.cuda.synchronize()
    got = bad.n_found.cpu().numpy()
    assert got[2] == -1 and got[4] == -1
    keep = [0, 1, 3, 5]
    assert np.array_equal(got[keep], want[keep])
    for p in keep:
        o, sz = int(good.prob_off_h[p]), int(pb.nM[p]) * (int(pb.nL[p]) + 1)
        assert np.array_equal(bad.probs.cpu().numpy()[o:o + sz], want_probs[o:o + sz])
