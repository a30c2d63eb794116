"""Distributed z solve (flutas_b200/csrc/dz.cuh) on the CPU: G ranks each own nz/G levels of every column; two rank-local
sweeps (the shared-LU kernel's own phase functions, tests/emulate) around a 2G x 2G interface system per column must
reproduce the reference's sequential elimination (gaussel / gaussel_periodic, src/solver_cpu.f90:117-223)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from flutas_b200 import initsolver
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emulate", "emul.cpp")
    so = os.path.join(HERE, "emulate", "libemul.so")
    deps = [src] + [os.path.join(HERE, "..", "flutas_b200", "csrc", f)
                    for f in ("tile_fft.cuh", "line_plan.h", "thomas_tile.cuh", "thomas_reg.cuh", "reg_fft.cuh", "thomas_uni.cuh",
                              "thomas_ref.cuh", "dz.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emul_dz.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    L.emul_dz_general.argtypes = L.emul_dz.argtypes
    return L


def _p(x):
    return x.ctypes.data_as(_dp)


@pytest.mark.parametrize("bcz", ["NN", "PP", "DD", "ND"])
@pytest.mark.parametrize("nz,G", [(64, 2), (128, 2), (128, 4), (256, 8), (512, 4), (1024, 8)])
def test_distributed_z_matches_reference_thomas(emul, bcz, nz, G):
    periodic = 1 if bcz == "PP" else 0
    rng = np.random.default_rng(nz + G)
    dzc, dzf = initsolver.initgrid(nz, 0.0, 1.0, 1)          # lz = 1: exactly uniform rows (thomas_detect_uniform)
    a, b, c = initsolver.tridmatrix(bcz, nz, 1, 1.0 / dzc, 1.0 / dzf)
    singular = 1 if bcz in ("NN", "PP") else 0
    nx, ny = 6, 3                                           # 18 columns: a ragged last tile
    lam = -rng.uniform(0.0, 4.0 * nz * nz, (nx, ny))
    lam[0, 0] = 0.0                                          # singular column when the z operator is (pinned gauge)
    lam[1, 0] = -40.0 * nz * nz / 1024.0 ** 2 * 1024.0       # moderately conditioned
    lam = np.asfortranarray(lam)
    rhs = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
    if singular:
        rhs[0, 0, :] -= rhs[0, 0, :].mean()
    ref = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), bool(periodic))
    got = rhs.copy(order="F")
    assert emul.emul_dz(nz, G, nx * ny, periodic, singular, _p(a), _p(b), _p(c), _p(lam), _p(got)) == 0
    for i in range(nx):
        for j in range(ny):
            g, r = got[i, j, :], ref[i, j, :]
            if singular and i == 0 and j == 0:
                assert g[-1] == 0.0                          # the gauge of the single-rank kernels: x(nz) = 0
                A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
                if periodic:
                    A[0, nz - 1] += a[0]
                    A[nz - 1, 0] += c[nz - 1]
                assert np.max(np.abs(A @ g - rhs[i, j, :])) <= 1e-11 * np.max(np.abs(a))
                continue
            tol = max(3e-13, 4.0 * np.abs(a).max() / max(abs(lam[i, j]), 1.0) * 2e-15)   # cond * eps, as for every other elimination order
            assert np.max(np.abs(g - r)) <= tol * np.max(np.abs(r)), (i, j, lam[i, j])


def test_interface_couplings_vanish_at_walls(emul):
    """non-periodic: rank 0 has no lower neighbour and rank G-1 no upper one -- a constant right-hand side with Dirichlet walls
    gives the same parabola whatever G"""
    nz = 256
    dzc, dzf = initsolver.initgrid(nz, 0.0, 1.0, 1)
    a, b, c = initsolver.tridmatrix("DD", nz, 1, 1.0 / dzc, 1.0 / dzf)
    lam = np.asfortranarray(np.zeros((2, 1)))
    sols = []
    for G in (2, 4, 8):
        w = np.asfortranarray(np.ones((2, 1, nz)))
        assert emul.emul_dz(nz, G, 2, 0, 0, _p(a), _p(b), _p(c), _p(lam), _p(w)) == 0
        sols.append(w[0, 0, :].copy())
    z = (np.arange(nz) + 0.5) / nz
    exact = -0.5 * z * (1.0 - z)                             # p'' = 1, p(0) = p(1) = 0 (second-order exact for a parabola up to the wall closure)
    for sol in sols:
        assert np.max(np.abs(sol - sols[0])) <= 1e-13 * np.max(np.abs(sols[0]))
        assert np.max(np.abs(sol - exact)) <= 2.0 / nz ** 2


@pytest.mark.parametrize("bcz,gr", [("NN", 2.0), ("DD", 1.0), ("ND", 3.0), ("NN", 0.0), ("PP", 0.0)])
@pytest.mark.parametrize("nz,G", [(64, 2), (128, 4), (256, 8), (512, 2)])
def test_distributed_z_on_a_general_grid(emul, bcz, gr, nz, G):
    """tanh-stretched (and round-off-non-uniform) z grids: the local blocks go through the general kernel's phase
    functions with the local coefficient rows (thomas_reg_local_run)"""
    periodic = 1 if bcz == "PP" else 0
    rng = np.random.default_rng(nz + G + int(gr))
    dzc, dzf = initsolver.initgrid(nz, gr, 2.0 * np.pi if gr == 0.0 else 1.0, 1)      # lz = 2 pi: rows differ in the last bits
    a, b, c = initsolver.tridmatrix(bcz, nz, 1, 1.0 / dzc, 1.0 / dzf)
    singular = 1 if bcz in ("NN", "PP") else 0
    nx, ny = 4, 4
    lam = -rng.uniform(0.01 * np.abs(a).max(), 4.0 * np.abs(a).max(), (nx, ny))
    lam[0, 0] = 0.0
    lam = np.asfortranarray(lam)
    rhs = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
    if singular:
        rhs[0, 0, :] -= (rhs[0, 0, :] * dzf[1:-1]).sum() / dzf[1:-1].sum()
    ref = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), bool(periodic))
    got = rhs.copy(order="F")
    assert emul.emul_dz_general(nz, G, nx * ny, periodic, singular, _p(a), _p(b), _p(c), _p(lam), _p(got)) == 0
    for i in range(nx):
        for j in range(ny):
            g, r = got[i, j, :], ref[i, j, :]
            if singular and i == 0 and j == 0:
                assert g[-1] == 0.0
                A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
                if periodic:
                    A[0, nz - 1] += a[0]
                    A[nz - 1, 0] += c[nz - 1]
                assert np.max(np.abs(A @ g - rhs[i, j, :])) <= 1e-11 * np.max(np.abs(a))
                continue
            tol = max(3e-13, 4.0 * np.abs(a).max() / max(abs(lam[i, j]), 1.0) * 2e-15)
            assert np.max(np.abs(g - r)) <= tol * np.max(np.abs(r)), (i, j, lam[i, j])
