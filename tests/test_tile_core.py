"""Checks the device code's index/arithmetic core (flutas_b200/csrc/tile_fft.cuh) on the CPU: the same
phase functions the CUDA kernels call are run by tests/emulate/emul.cpp and compared with the oracle's
FFTW-definition transforms through the plan's row -> mode map."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
KIND = {"PP": 0, "NN": 1, "DD": 2}
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emulate", "emul.cpp")
    so = os.path.join(HERE, "emulate", "libemul.so")
    deps = [src] + [os.path.join(HERE, "..", "flutas_b200", "csrc", f) for f in ("tile_fft.cuh", "line_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emul_line_transform.argtypes = [C.c_int] * 6 + [_dp, _dp, C.c_double]
    L.emul_mode_index.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def _mode(emul, n, bc):
    mode = (C.c_int * n)()
    assert emul.emul_mode_index(n, KIND[bc], mode) == 0
    return np.array(mode[:])


@pytest.mark.parametrize("bc", ["PP", "NN", "DD"])
@pytest.mark.parametrize("n", [2, 4, 6, 8, 10, 12, 16, 30, 32, 64, 72, 100, 128, 256, 512, 1024])
@pytest.mark.parametrize("tb,rot", [(16, 0), (16, 1), (8, 1), (8, 0), (4, 1)])
def test_tile_forward_matches_fftw_definition(emul, bc, n, tb, rot):
    if n >= 256 and (tb, rot) not in ((16, 0), (8, 1)):
        pytest.skip("large sizes: two layouts are enough")
    rng = np.random.default_rng(n + tb)
    x = rng.uniform(-1, 1, (tb, n))
    out = np.zeros_like(x)
    nworkers = 5
    assert emul.emul_line_transform(n, KIND[bc], tb, rot, nworkers, 1, x.ctypes.data_as(_dp),
                                    out.ctypes.data_as(_dp), 1.0) == 0
    kf, kb, norm = oracle.find_fft(bc)
    ref = oracle.r2r(kf, np.asfortranarray(x.T.reshape(n, tb, 1).copy()), 0)[:, :, 0].T   # (tb, n), FFTW order
    mode = _mode(emul, n, bc)
    assert sorted(mode) == list(range(n))
    scale = max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(out - ref[:, mode])) <= 5e-14 * scale
    # backward: spectral tile layout -> physical, equals FFTW bwd kind applied to FFTW-ordered data
    back = np.zeros_like(x)
    assert emul.emul_line_transform(n, KIND[bc], tb, rot, nworkers, 0, out.ctypes.data_as(_dp),
                                    back.ctypes.data_as(_dp), 1.0) == 0
    assert np.max(np.abs(back - x * norm[0] * (n + norm[1]))) <= 1e-13 * n
    spec = rng.uniform(-1, 1, (tb, n))                       # arbitrary (non-consistent) spectrum
    fftw_order = np.zeros_like(spec)
    fftw_order[:, mode] = spec
    if bc == "PP" and n >= 2:
        pass
    refb = oracle.r2r(kb, np.asfortranarray(fftw_order.T.reshape(n, tb, 1).copy()), 0)[:, :, 0].T
    assert emul.emul_line_transform(n, KIND[bc], tb, rot, nworkers, 0, spec.ctypes.data_as(_dp),
                                    back.ctypes.data_as(_dp), 1.0) == 0
    assert np.max(np.abs(back - refb)) <= 5e-14 * max(1.0, np.max(np.abs(refb)))


def test_unsupported_lengths_are_rejected(emul):
    out = np.zeros(14)
    assert emul.emul_line_transform(14, 0, 16, 0, 1, 1, out.ctypes.data_as(_dp), out.ctypes.data_as(_dp), 1.0) == 1
    assert emul.emul_line_transform(9, 0, 16, 0, 1, 1, out.ctypes.data_as(_dp), out.ctypes.data_as(_dp), 1.0) == 1
