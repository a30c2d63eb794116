"""Pins the CPU oracle (oracle/flutas_oracle.c): FFTW r2r definitions vs scipy/pocketfft and vs an
O(n^2) long-double evaluation, the Thomas solves vs numpy, the whole solve vs the golden fixtures and
vs the residual of the discrete operator, and the host-side initsolver mirror vs the oracle's."""
import numpy as np
import pytest
import scipy.fft as sf

from flutas_b200 import initsolver
from flutas_b200.cases import Case
from oracle import oracle
from util import gauge_rel_err, golden_files, load_golden

SIZES = [2, 4, 6, 8, 10, 12, 30, 64, 72, 100]


def _halfcomplex(x):
    n = len(x)
    f = sf.rfft(x)
    out = np.zeros(n)
    out[: n // 2 + 1] = f.real
    for k in range(1, (n + 1) // 2):
        out[n - k] = f[k].imag
    return out


def _scipy_kind(kind, x):
    n = len(x)
    if kind == "R2HC":
        return _halfcomplex(x)
    if kind == "HC2R":
        f = np.zeros(n // 2 + 1, dtype=complex)
        f.real = x[: n // 2 + 1]
        for k in range(1, (n + 1) // 2):
            f[k] += 1j * x[n - k]
        return sf.irfft(f, n=n) * n
    fn, t = {"REDFT10": (sf.dct, 2), "REDFT01": (sf.dct, 3), "REDFT11": (sf.dct, 4),
             "RODFT10": (sf.dst, 2), "RODFT01": (sf.dst, 3), "RODFT11": (sf.dst, 4)}[kind]
    return fn(x, type=t, norm=None)


def _longdouble_kind(kind, x):
    """FFTW's published r2r definitions, evaluated directly in long double."""
    n = len(x)
    x = x.astype(np.longdouble)
    j = np.arange(n, dtype=np.longdouble)
    k = j[:, None]
    pi = np.longdouble(np.pi) + np.longdouble(1.2246467991473532e-16)
    if kind == "R2HC":
        re = (x[None, :] * np.cos(2 * pi * j[None, :] * k / n)).sum(1)
        im = -(x[None, :] * np.sin(2 * pi * j[None, :] * k / n)).sum(1)
        out = np.zeros(n, dtype=np.longdouble)
        out[: n // 2 + 1] = re[: n // 2 + 1]
        for q in range(1, (n + 1) // 2):
            out[n - q] = im[q]
        return out
    if kind == "REDFT10":
        return 2 * (x[None, :] * np.cos(pi * (j[None, :] + 0.5) * k / n)).sum(1)
    if kind == "REDFT01":
        return x[0] + 2 * (x[None, 1:] * np.cos(pi * j[None, 1:] * (k + 0.5) / n)).sum(1)
    if kind == "RODFT10":
        return 2 * (x[None, :] * np.sin(pi * (j[None, :] + 0.5) * (k + 1) / n)).sum(1)
    if kind == "RODFT01":
        return (-1.0) ** j * x[n - 1] + 2 * (x[None, :-1] * np.sin(pi * (j[None, :-1] + 1) * (k + 0.5) / n)).sum(1)
    if kind == "REDFT11":
        return 2 * (x[None, :] * np.cos(pi * (j[None, :] + 0.5) * (k + 0.5) / n)).sum(1)
    if kind == "RODFT11":
        return 2 * (x[None, :] * np.sin(pi * (j[None, :] + 0.5) * (k + 0.5) / n)).sum(1)
    raise ValueError(kind)


KINDS = ["R2HC", "HC2R", "REDFT10", "REDFT01", "RODFT10", "RODFT01", "REDFT11", "RODFT11"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", SIZES)
def test_r2r_matches_scipy_both_axes(kind, n):
    rng = np.random.default_rng(n * 31 + len(kind))
    for axis, shape in ((0, (n, 3, 2)), (1, (4, n, 2))):
        x = np.asfortranarray(rng.uniform(-1, 1, shape))
        ref = np.apply_along_axis(lambda v: _scipy_kind(kind, v), axis, x)
        got = oracle.r2r(kind, x.copy(order="F"), axis)
        assert np.max(np.abs(got - ref)) <= 2e-14 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [4, 6, 8, 10, 12, 30, 32, 64, 72, 100, 256, 1024, 2048])
def test_fast_path_matches_definition_path(kind, n):
    """The oracle's default transforms use half-length reductions; the path that transcribes FFTW's definitions through
    one zero-padded complex FFT is kept and must agree with it to round-off (both are also pinned to scipy above)."""
    rng = np.random.default_rng(n + len(kind))
    for axis, shape in ((0, (n, 3, 2)), (1, (3, n, 2))):
        x = np.asfortranarray(rng.uniform(-1, 1, shape))
        try:
            oracle.set_definition_path(True)
            ref = oracle.r2r(kind, x.copy(order="F"), axis)
        finally:
            oracle.set_definition_path(False)
        got = oracle.r2r(kind, x.copy(order="F"), axis)
        assert np.max(np.abs(got - ref)) <= 5e-15 * max(1.0, np.max(np.abs(ref))) * np.log2(n)


@pytest.mark.parametrize("shape,bcz,gr", [((13, 5, 16), "NN", 0.0), ((8, 3, 7), "NN", 2.0), ((17, 4, 12), "PP", 0.0),
                                          ((3, 2, 4), "PP", 0.0), ((16, 16, 64), "DD", 1.0), ((9, 2, 3), "ND", 0.0),
                                          ((24, 3, 2), "NN", 0.0)])
def test_blocked_gaussel_is_bit_identical_to_the_column_sweep(shape, bcz, gr):
    """gaussel / gaussel_periodic (src/solver_cpu.f90:117-223) sweep one column at a time; the oracle's default runs the same
    recurrences on eight adjacent columns at once (one cache line per level).  Same operations in the same order per column
    (-ffp-contract=off), so the two must agree bit for bit -- including the singular column's z == 0 branch."""
    from flutas_b200 import initsolver
    nx, ny, n = shape
    rng = np.random.default_rng(nx * 100 + n)
    dzc, dzf = initsolver.initgrid(n, gr, 1.0, 1)
    a, b, c = initsolver.tridmatrix(bcz, n, 1, 1.0 / dzc, 1.0 / dzf)
    lam = np.asfortranarray(-rng.uniform(0, 4 * n * n, (nx, ny)))
    lam[0, 0] = 0.0
    rhs = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, n)))
    try:
        oracle.set_definition_path(True)
        ref = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), bcz == "PP")
    finally:
        oracle.set_definition_path(False)
    got = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), bcz == "PP")
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])


@pytest.mark.parametrize("kind", [k for k in KINDS if k != "HC2R"])
@pytest.mark.parametrize("n", [2, 6, 8, 12, 30])
def test_r2r_matches_longdouble_definition(kind, n):
    rng = np.random.default_rng(n + 7)
    x = rng.uniform(-1, 1, n)
    ref = _longdouble_kind(kind, x).astype(np.float64)
    got = oracle.r2r(kind, np.asfortranarray(x.reshape(n, 1, 1).copy()), 0).ravel()
    assert np.max(np.abs(got - ref)) <= 1e-14 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("bc", ["PP", "NN", "DD", "ND", "DN"])
def test_forward_backward_is_identity_times_norm(bc):
    kf, kb, norm = oracle.find_fft(bc)
    n = 24
    x = np.asfortranarray(np.random.default_rng(3).uniform(-1, 1, (n, 2, 2)))
    y = oracle.r2r(kb, oracle.r2r(kf, x.copy(order="F"), 0), 0)
    assert np.allclose(y, x * norm[0] * (n + norm[1]), rtol=0, atol=1e-12)


def test_initsolver_mirror_is_bit_exact_with_oracle():
    for bc in ("PP", "NN", "DD", "ND", "DN"):
        for n in (2, 8, 30, 64):
            assert np.array_equal(initsolver.eigenvalues(n, bc), oracle.eigenvalues(n, bc))
    for gr in (0.0, 2.0):
        for nh in (1, 3):
            dzc, dzf = initsolver.initgrid(20, gr, 1.5, nh)
            odzc, odzf = oracle.initgrid(20, gr, 1.5, nh)
            # tanh comes from libm in C and from numpy's SIMD loops in Python: 1-ulp differences (amplified by zf(k)-zf(k-1)) are allowed
            assert np.allclose(dzc, odzc, rtol=1e-13, atol=0) and np.allclose(dzf, odzf, rtol=1e-13, atol=0)
            for bcz in ("PP", "NN", "DD", "ND", "DN"):
                a, b, c = initsolver.tridmatrix(bcz, 20, nh, 1.0 / dzc, 1.0 / dzf)
                oa, ob, oc = oracle.tridmatrix(bcz, 20, nh, 1.0 / dzc, 1.0 / dzf)
                assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    s = initsolver.SolverSetup((16, 12, 10), (1.0, 2.0, 3.0), ("PP", "NN", "DD"))
    assert s.normfft == oracle.normfft(16, 12, "PP", "NN")


def _dense_thomas_check(periodic):
    rng = np.random.default_rng(5)
    n, nx, ny = 12, 3, 2
    a = rng.uniform(0.5, 1.5, n)
    c = rng.uniform(0.5, 1.5, n)
    b = -(a + c)
    lam = -rng.uniform(0.1, 3.0, (nx, ny))
    rhs = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, n)))
    got = oracle.gaussel(a, b, c, lam, rhs.copy(order="F"), periodic)
    for i in range(nx):
        for j in range(ny):
            A = np.diag(b + lam[i, j]) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
            if periodic:
                A[0, n - 1] += a[0]
                A[n - 1, 0] += c[n - 1]
            ref = np.linalg.solve(A, rhs[i, j, :])
            assert np.allclose(got[i, j, :], ref, rtol=1e-12, atol=1e-13)


def test_gaussel_matches_dense_solve():
    _dense_thomas_check(False)


def test_gaussel_periodic_matches_dense_solve():
    _dense_thomas_check(True)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_solver_matches_golden_and_residual(path):
    case, rhs, pgold = load_golden(path)
    s = case.setup
    sol = oracle.Solver(case.ng, case.cbc[0], case.cbc[1])
    p = case.new_p()
    p[1:-1, 1:-1, 1:-1] = rhs
    sol.solve(s.lambdaxy, s.a, s.b, s.c, case.cbc[2], p)
    err = gauge_rel_err(p[1:-1, 1:-1, 1:-1], pgold, case.singular)
    assert err <= 1e-12, err
    case.boundp(p)
    res = np.max(np.abs(case.laplacian(p) - rhs)) / np.max(np.abs(rhs))
    assert res <= 1e-12, res


def test_oracle_pressure_step_projects_to_divergence_free():
    for cbc in (("PP", "PP", "PP"), ("PP", "PP", "NN"), ("NN", "NN", "NN"), ("DD", "NN", "PP")):
        case = Case((16, 12, 10), cbc, (2.0, 1.0, 1.0), rho0=0.5, gr=(1.0 if cbc[2] != "PP" else 0.0), seed=11)
        s = case.setup
        u, v, w = case.velocity()
        p = case.new_p()
        oracle.fillps(case.ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, p)
        oracle.updt_rhs_b(case.ng, s.rhsbx, s.rhsby, s.rhsbz, p)
        sol = oracle.Solver(case.ng, cbc[0], cbc[1])
        sol.solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], p)
        case.boundp(p)
        oracle.correc(case.ng, case.nh_d, case.nh_u, s.dli, s.dzci, case.dt, case.rho0, p, u, v, w)
        case.correct_dirichlet_faces(p, u, v, w)
        case.refresh_velocity_halos(u, v, w)
        divtot, divmax = oracle.chkdiv(case.ng, s.dli, case.nh_d, case.nh_u, s.dzfi, u, v, w)
        assert divmax <= 1e-12, (cbc, divmax)
