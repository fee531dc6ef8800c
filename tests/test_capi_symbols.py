"""CPU-only: the C-ABI shared library loads without a GPU, exports every symbol that include/pda_b200.h
declares, and fails loudly (no fallback) when asked to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from probabilisticsemslam_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pda_b200.h but not exported"
    # and the Python signature table covers the same set
    assert sorted(_lib.SIGNATURES) == names
    assert lib.pda_version() >= 100


def test_shim_library_exports_the_reference_entry_points():
    path = os.path.join(ROOT, "probabilisticsemslam_b200", "libpda_b200_shims.so")
    assert os.path.exists(path), "libpda_b200_shims.so missing: run __graft_entry__.build()"
    out = os.popen(f"nm -DC --defined-only {path}").read()
    for sym in ["kBest2D(", "kBest2DCutoff(", "assign2D(", "shortestPathCPP(", "assignmentProb(", "permanentProb(",
                "bruteForceProb(", "conditionCosts(", "toProbs(", "conditionedPermanentRaw(", "permanentExactRaw(",
                "permanentApproximationRaw(", "computeQuadricCostMatrixRaw(", "getAssignmentProbsFromMoments(",
                "getAssignmentProbsFromCosts(", "asgnBBRaw(", "assignmentProbBatch(", "permanentProbBatch(",
                "permanentExactShardedRaw(", "permanentExactLongRaw("]:
        assert sym in out, f"{sym} missing from the C++ drop-in layer"


def test_reference_typed_overloads_compile():
    """include/nwPerm.h and include/assignment.h with PDA_HAVE_EIGEN on (a stand-in <Eigen/Core>): the overloads that carry
    the reference's own argument types are at least type-checked on every CPU run (they RUN in tests/test_gpu_dropin.py)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "eigen_overloads_driver.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "cpp", "eigen_stub"),
                        "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]


def test_multi_device_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from probabilisticsemslam_b200 import _lib
    lib = _lib.lib()
    dev = (ctypes.c_int32 * 2)(0, 1)
    A = (ctypes.c_double * 4)(1, 2, 3, 4)
    out = (ctypes.c_double * 1)()
    assert lib.pda_permanent_sharded_host(A, 2, dev, 2, out) == -2 and b"no CUDA device" in lib.pda_last_error()
    assert lib.pda_permanent_sharded_host(A, 2, None, 0, out) == -1
    assert lib.pda_murty_batch_host_multi(None, None, None, None, 5, 10, 1, 42.0, 0, 0, None, None, None, None, None, None,
                                          0, None, None, None, dev, 2) == -1
    assert lib.pda_host_alloc(1024) is None      # page-locked memory needs a device too


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from probabilisticsemslam_b200 import _lib, api, synth
    assert _lib.lib().pda_device_count() == 0
    with pytest.raises(_lib.PdaError):
        api.kBest2DCutoff(5, synth.g1_dense(1, nM=3).matrix(0))
    with pytest.raises(_lib.PdaError):
        api.permanentExact(np.ones((3, 3)))


def test_host_side_knobs_work_without_a_device():
    """Argument checking and the process-wide knobs are host logic: they must behave without a GPU."""
    from probabilisticsemslam_b200 import _lib
    lib = _lib.lib()
    prev = lib.pda_murty_set_path(2)
    assert prev in (0, 1, 2, 3) and lib.pda_murty_set_path(prev) == 2
    assert lib.pda_murty_set_path(3) == prev and lib.pda_murty_set_path(prev) == 3   # the pruning kernel can be pinned
    assert lib.pda_murty_set_path(7) == -1 and b"path" in lib.pda_last_error()
    lib.pda_set_approx_seed(123)
    lib.pda_set_approx_seed(20260217)
    found = np.zeros(1, np.int32)
    assert lib.pda_murty_batch_host(None, None, None, None, 1, 5, 0, 0.0, 0, 0, None, None, None, None, None,
                                    found.ctypes.data, 0, None, None, None, 0) == -1          # NULL inputs
    assert lib.pda_permanent_approx_batch_host(None, None, None, None, 1, 300, 1, None, None, 0) == -1
    assert lib.pda_quadric_cost_batch_host(None, None, None, None, None, None, 1, 10.0, None, 0) == -1
    assert lib.pda_murty_batch_host(None, None, None, None, 0, 5, 0, 0.0, 0, 0, None, None, None, None, None,
                                    None, 0, None, None, None, 0) == 0                          # empty batch: nothing to do


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the package or the public headers may import, include,
    link or load it (or the reference build under oracle/_ref)."""
    bad = re.compile(r"(^|\s)(from|import)\s+oracle\b|liboracle|oracle_capi|load_oracle|load_reference|libpda_ref|#include\s+\"oracle")
    for top in ("probabilisticsemslam_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if os.sep + "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    m = bad.search(text)
                    assert m is None, f"{os.path.join(dirpath, f)} reaches into the oracle: {m.group(0)!r}"


def test_datfile_roundtrip(tmp_path):
    from probabilisticsemslam_b200 import datfile
    C = np.array([[1.5, np.inf], [0.1234567, 2.0], [10.0, np.inf], [np.inf, 10.0]])
    p = datfile.frame_path(str(tmp_path), "o30_p0_k200_perm0_net1", 7)
    assert p.endswith("o30_p0_k200_perm0_net1_frame7.dat")
    datfile.write_dat(p, C)
    assert open(p).read().splitlines()[1] == "0.123457,2.000000"       # std::to_string: 6 decimals
    back = datfile.read_dat(p)
    assert np.isinf(back[0, 1]) and back[1, 0] == 0.123457


def test_host_batch_argument_checks_run_before_any_device_work():
    """pda_murty_batch_host validates the whole batch on the host (dimensions, offsets, nL + numCol == numRow) before it
    touches a device: the same answers with or without a GPU."""
    from probabilisticsemslam_b200 import _lib
    lib = _lib.lib()
    costs = np.zeros(10)
    off = np.zeros(1, np.int64); poff = np.zeros(1, np.int64)
    found = np.zeros(1, np.int32); probs = np.zeros(8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    def call(nr, nc, nl, cost_off=0):
        off[0] = cost_off
        a, b, c = np.array([nr], np.int32), np.array([nc], np.int32), np.array([nl], np.int32)
        return lib.pda_murty_batch_host(p(costs), p(off), p(a), p(b), 1, 5, 1, 42.0, 0, 0, None, None, None, None, None, p(found),
                                        1, p(probs), p(poff), p(c), 0)
    assert call(2, 3, 0) == -1 and b"numRow >= numCol" in lib.pda_last_error()      # more detections than rows
    assert call(5, 2, 2) == -1 and b"nL + numCol" in lib.pda_last_error()           # rows are not landmarks + detections
    assert call(5, 2, 3, cost_off=-4) == -1 and b"negative offset" in lib.pda_last_error()
    assert lib.pda_murty_batch_host(p(costs), p(off), None, None, 1, 5, 1, 42.0, 0, 0, None, None, None, None, None, p(found),
                                    1, p(probs), p(poff), None, 0) == -1           # NULL inputs
