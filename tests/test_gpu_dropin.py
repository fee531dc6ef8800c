"""Drop-in demonstration: ONE compMethods-style C++ caller (tests/cpp/compmethods_driver.cpp, modelled on
comparison.cpp:146-247) built twice -- against the reference's own code (oracle/_ref/compmethods_ref) and
against include/*.h + libpda_b200_shims.so (tests/cpp/build/compmethods_b200) -- must print the same
k-best lists and gains and the same probabilities (1e-9) for the same .dat cost-matrix files."""
import os
import resource
import subprocess

import numpy as np
import pytest

from probabilisticsemslam_b200 import datfile, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "compmethods_ref")
GPU_BIN = os.path.join(ROOT, "tests", "cpp", "build", "compmethods_b200")


def _big_stack():
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))


def _parse(text):
    recs = []
    for line in text.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "h":          # h <i> <gain> : r4c... | c4r...
            recs.append(("h", [float(t[2])], [int(x) for x in t[4:] if x != "|"]))
        elif t[0] in ("frame", "kbest"):
            recs.append((line, [], []))
        else:                     # <tag> <m> p0 p1 ...
            recs.append((t[0] + " " + t[1], [float(x) for x in t[2:]], []))
    return recs


def test_same_caller_two_backends(tmp_path, gpu_api):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/compmethods_ref not built (make -C oracle ref_driver where /root/reference is mounted)")
    assert os.path.exists(GPU_BIN), "tests/cpp/build/compmethods_b200 missing: run __graft_entry__.build()"
    files = []
    g2 = synth.g2_gated(10, first=900)
    for p in range(len(g2)):               # the wire format quantises to 6 decimals: exact ties appear
        path = datfile.frame_path(str(tmp_path), "o30_p0_k200_perm0_net1", p + 1)
        datfile.write_dat(path, g2.matrix(p))
        np.testing.assert_allclose(datfile.read_dat(path), np.round(g2.matrix(p), 6), rtol=0, atol=1e-6)
        files.append(path)
    small = synth.g1_dense(3, nL=9, nM=3, first=77)
    for p in range(3):
        path = datfile.frame_path(str(tmp_path), "small", p + 1)
        datfile.write_dat(path, small.matrix(p))
        files.append(path)
    ref = subprocess.run([REF_BIN, "--k", "150"] + files, capture_output=True, text=True, preexec_fn=_big_stack, timeout=600)
    assert ref.returncode == 0, ref.stderr[-500:]
    gpu = subprocess.run([GPU_BIN, "--k", "150"] + files, capture_output=True, text=True, timeout=600)
    assert gpu.returncode == 0, gpu.stderr[-500:]
    a, b = _parse(ref.stdout), _parse(gpu.stdout)
    assert len(a) == len(b) and len(a) > 100
    for (ta, fa, ia), (tb, fb, ib) in zip(a, b):
        assert ta == tb
        assert ia == ib, f"{ta}: index lists differ"
        if ta == "h":
            assert fa == fb, "gains must be bit-identical"
        elif ta.startswith("permExact"):
            np.testing.assert_allclose(fb, fa, rtol=1e-6, atol=1e-300)   # ill-conditioned NW sums: see test_gpu_permanent
        else:
            np.testing.assert_allclose(fb, fa, rtol=1e-9, atol=0)
