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
KIND = {"PP": 0, "NN": 1, "DD": 2, "ND": 3, "DN": 4}
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emulate", "emul.cpp")
    so = os.path.join(HERE, "emulate", "libemul.so")
    deps = [src] + [os.path.join(HERE, "..", "flutas_b200", "csrc", f)
                    for f in ("tile_fft.cuh", "line_plan.h", "thomas_tile.cuh", "thomas_reg.cuh", "reg_fft.cuh", "thomas_uni.cuh",
                              "thomas_ref.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emul_line_transform.argtypes = [C.c_int] * 6 + [_dp, _dp, C.c_double]
    L.emul_mode_index.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.emul_reg_line_transform.argtypes = [C.c_int] * 3 + [_dp, _dp, C.c_double]
    L.emul_reg_mode_index.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def _mode(emul, n, bc):
    mode = (C.c_int * n)()
    assert emul.emul_mode_index(n, KIND[bc], mode) == 0
    return np.array(mode[:])


@pytest.mark.parametrize("bc", ["PP", "NN", "DD", "ND", "DN"])
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


# ---- register-resident transforms (reg_fft.cuh) --------------------------------------------------
@pytest.mark.parametrize("bc", ["PP", "NN", "DD", "ND", "DN"])
@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, (1024, 8), (1024, "pair"), (2048, "pair")], ids=str)
def test_reg_fft_matches_fftw_definition(emul, bc, n):
    transform = emul.emul_reg_line_transform
    pair = isinstance(n, tuple) and n[1] == "pair"      # small-radix last pass on symmetric pairs, split in registers
    emul.emul_reg_pair_mode(1 if pair else 0)
    if pair:
        n = n[0]
    if isinstance(n, tuple):                               # the 8-values-per-thread schedule of the N = 1024 y kernels
        n = n[0]
        emul.emul_reg_line_transform8.argtypes = emul.emul_reg_line_transform.argtypes
        transform = emul.emul_reg_line_transform8
    rng = np.random.default_rng(7 * n + len(bc))
    nl = 3
    x = rng.uniform(-1, 1, (nl, n))
    kf, kb, norm = oracle.find_fft(bc)
    ref = oracle.r2r(kf, np.asfortranarray(x.T.reshape(n, nl, 1).copy()), 0)[:, :, 0].T      # FFTW order
    mode = (C.c_int * n)()
    assert emul.emul_reg_mode_index(n, KIND[bc], mode) == 0
    mode = np.array(mode[:])
    assert sorted(mode) == list(range(n))
    spec = rng.uniform(-1, 1, (nl, n))                       # arbitrary (non-consistent) spectrum
    fftw_order = np.zeros_like(spec)
    fftw_order[:, mode] = spec
    refb = oracle.r2r(kb, np.asfortranarray(fftw_order.T.reshape(n, nl, 1).copy()), 0)[:, :, 0].T
    for l in range(nl):
        out, back, back2 = np.zeros(n), np.zeros(n), np.zeros(n)
        xin = np.ascontiguousarray(x[l])
        assert transform(n, KIND[bc], 1, xin.ctypes.data_as(_dp), out.ctypes.data_as(_dp), 1.0) == 0
        assert np.max(np.abs(out - ref[l, mode])) <= 5e-14 * max(1.0, np.max(np.abs(ref[l])))
        assert transform(n, KIND[bc], 0, out.ctypes.data_as(_dp), back.ctypes.data_as(_dp), 1.0) == 0
        assert np.max(np.abs(back - xin * norm[0] * (n + norm[1]))) <= 1e-13 * n
        sp = np.ascontiguousarray(spec[l])
        assert transform(n, KIND[bc], 0, sp.ctypes.data_as(_dp), back2.ctypes.data_as(_dp), 0.5) == 0
        assert np.max(np.abs(back2 - 0.5 * refb[l])) <= 5e-14 * max(1.0, np.max(np.abs(refb[l])))


def test_unsupported_lengths_are_rejected(emul):
    out = np.zeros(14)
    assert emul.emul_line_transform(14, 0, 16, 0, 1, 1, out.ctypes.data_as(_dp), out.ctypes.data_as(_dp), 1.0) == 1
    assert emul.emul_line_transform(9, 0, 16, 0, 1, 1, out.ctypes.data_as(_dp), out.ctypes.data_as(_dp), 1.0) == 1


# ---- on-chip Thomas (thomas_tile.cuh) -------------------------------------------------------------
@pytest.fixture(scope="module")
def emul_t(emul):
    emul.emul_thomas_tile.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    emul.emul_thomas_reg.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int]
    emul.emul_thomas_uni.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    return emul


@pytest.mark.parametrize("periodic", [0, 1])
@pytest.mark.parametrize("nz,L", [(8, 2), (16, 4), (32, 8), (64, 8), (72, 8), (40, 8), (64, 16), (128, 16), (512, 16),
                                  (256, 32), (1024, 32), (12, 2), (24, 4), (1024, 16), (48, 16), (4, 2)])
@pytest.mark.parametrize("stretched", [False, True])
@pytest.mark.parametrize("variant", ["tile", "reg", "reg-uniform", "uni"])
def test_thomas_tile_matches_reference_thomas(emul_t, periodic, nz, L, stretched, variant):
    if variant in ("reg-uniform", "uni") and (stretched or nz < 4):
        pytest.skip("scalar-coefficient path: uniform grids only")
    if variant in ("reg", "reg-uniform") and L > 16:
        pytest.skip("the register kernel keeps at most 16 levels per thread")
    if variant == "uni":                                 # the shared-LU kernel picks its own segment length
        L = next((q for q in (4, 8, 16, 32) if nz % q == 0 and 2 <= nz // q <= 32
                  and not (periodic and (nz // q) & (nz // q - 1))), None)
        if L is None:
            pytest.skip("nz not served by the shared-LU kernel (falls back to the register kernel)")
    S = nz // L
    if periodic and (S & (S - 1)):
        pytest.skip("cyclic PCR needs a power-of-two number of separators (falls back to the generic kernel)")
    if periodic and stretched:
        pytest.skip("periodic z implies a uniform grid")
    from flutas_b200 import initsolver
    rng = np.random.default_rng(nz + L)
    dzc, dzf = initsolver.initgrid(nz, 2.0 if stretched else 0.0, 1.0, 1)
    bcz = "PP" if periodic else "NN"
    a, b, c = initsolver.tridmatrix(bcz, nz, 1, 1.0 / dzc, 1.0 / dzf)
    if variant in ("reg-uniform", "uni"):
        # exactly uniform coefficients (what initgrid produces when lz/nz is a binary fraction, e.g. lz = 1, nz = 512);
        # the oracle solves with the same arrays
        a0 = a[1]
        a[:] = a0
        c[:] = a0
        b[:] = -2.0 * a0
        if not periodic:
            b[0] += a0                                   # Neumann walls: b(1) += a(1), b(n) += c(n) (initsolver.f90:228-236)
            b[-1] += a0
    nx, ny = 7, 3                                        # 21 columns: exercises the ragged last tile
    lam = -rng.uniform(0.0, 4.0 * nz * nz, (nx, ny))
    lam[0, 0] = 0.0                                      # singular column
    lam[1, 0] = -1.0e-3                                  # nearly singular column
    rhs = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
    rhs[0, 0, :] -= (rhs[0, 0, :] * dzf[1:-1]).sum() / dzf[1:-1].sum()    # compatible RHS for the singular column
    ref = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), bool(periodic))
    got = rhs.copy(order="F")
    args = (L, nz, nx * ny, periodic, 1, a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), c.ctypes.data_as(_dp),
            np.asfortranarray(lam).ctypes.data_as(_dp), got.ctypes.data_as(_dp))
    if variant == "tile":
        rc = emul_t.emul_thomas_tile(*args)
    elif variant == "uni":
        rc = 0 if emul_t.emul_thomas_uni(*args[1:]) == L else 5
    else:
        rc = emul_t.emul_thomas_reg(*args, 2 if variant == "reg-uniform" else 0)
    assert rc == 0
    for i in range(nx):
        for j in range(ny):
            g, r = got[i, j, :], ref[i, j, :]
            if i == 0 and j == 0:
                # pinned gauge x(nz) = 0.  The reference leaves this column's constant to round-off (and returns
                # NaN when its closure denominator is exactly 0), so check the residual of the singular system.
                assert g[-1] == 0.0
                A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
                if periodic:
                    A[0, nz - 1] += a[0]
                    A[nz - 1, 0] += c[nz - 1]
                assert np.max(np.abs(A @ g - rhs[i, j, :])) <= 1e-11 * np.max(np.abs(a))
                continue
            else:
                # forward error of any stable elimination order ~ cond * eps; cond ~ 4 max(a) / |lambda|
                tol = max(2e-13, 4.0 * a.max() / abs(lam[i, j]) * 2e-15)
            assert np.max(np.abs(g - r)) <= tol * max(np.max(np.abs(r)), 1e-300), (i, j, lam[i, j])
