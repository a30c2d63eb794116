"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.
Tolerance: max|dp|/max|p| <= 1e-12 in FP64 after removing the mean when the operator is singular
(BASELINE.json north_star); stencils (fillps/correc/chkdiv divmax) bit-exact."""
import numpy as np
import pytest

from flutas_b200 import api, lib
from flutas_b200.cases import Case
from oracle import oracle
from util import gauge_rel_err, golden_files, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _solve_gpu(case, rhs, device=False, z_mode=0, generic_fft=False):
    s = case.setup
    n = case.ng
    lib.load().flutas_b200_debug_generic_fft(int(generic_fft))     # 0 = register kernels, 1 = generic tile, 2 = p2 tile
    pl, nf = api.fftini(n, n, (case.cbc[0], case.cbc[1]))
    assert nf == s.normfft
    if z_mode:                   # 1 = generic scratch-field kernels, 2 = shared-memory tile kernel, 3 = no uniform-grid path
        lib.check(lib.load().flutas_b200_debug_thomas_mode(pl.h, z_mode))
    p = case.new_p()
    p[...] = 7.0                                      # halos must come back untouched
    p[1:-1, 1:-1, 1:-1] = rhs
    if device:
        import torch
        pd = api.device_field(p)
        api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, case.cbc[2], "ccc", pd)
        torch.cuda.synchronize()
        p = api.host_field(pd, p.shape)
    else:
        api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, case.cbc[2], "ccc", p)
    api.fftend(pl)
    lib.load().flutas_b200_debug_generic_fft(0)
    halo = p.copy()
    halo[1:-1, 1:-1, 1:-1] = 7.0
    assert np.all(halo == 7.0)
    return p


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
@pytest.mark.parametrize("z_mode", [0, 3, 2, 1], ids=["zreg", "zreg-tables", "ztile", "zgeneric"])
def test_solver_matches_golden(path, z_mode):
    case, rhs, pgold = load_golden(path)
    p = _solve_gpu(case, rhs, z_mode=z_mode)
    err = gauge_rel_err(p[1:-1, 1:-1, 1:-1], pgold, case.singular)
    assert err <= TOL, err
    case.boundp(p)
    res = np.max(np.abs(case.laplacian(p) - rhs)) / np.max(np.abs(rhs))
    assert res <= 1e-11, res


def test_unsupported_bc_fails_loudly():
    with pytest.raises(lib.FlutasB200Error, match="PP, NN, DD, ND, DN only"):
        api.fftini((8, 8, 8), (8, 8, 8), ("PN", "PP"))          # not a pair of src/fft.f90:233-291
    with pytest.raises(lib.FlutasB200Error, match="factors into 2,3,5"):
        api.fftini((14, 8, 8), (14, 8, 8), ("PP", "PP"))


CASES = [
    ("C1", (64, 64, 64), ("PP", "PP", "PP"), (2 * np.pi,) * 3, 0.0),
    ("chan", (128, 64, 32), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), 2.0),
    ("chan72", (64, 32, 72), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), 1.0),
    ("rb", (96, 80, 40), ("NN", "NN", "NN"), (2.0, 2.0, 1.0), 0.0),
    ("dd", (40, 48, 16), ("DD", "NN", "DD"), (1.0, 2.0, 1.0), 0.0),
    ("pnp", (32, 36, 24), ("PP", "NN", "PP"), (2.0, 1.0, 1.0), 0.0),
    ("ragged", (20, 12, 10), ("NN", "DD", "NN"), (1.0, 1.0, 1.0), 0.5),
    ("wide", (1024, 16, 8), ("PP", "NN", "NN"), (4.0, 1.0, 1.0), 0.0),
    ("tall", (16, 2048, 4), ("DD", "PP", "NN"), (1.0, 4.0, 1.0), 0.0),
    ("deep", (16, 16, 256), ("PP", "PP", "PP"), (1.0, 1.0, 4.0), 0.0),
    ("p2a", (512, 256, 8), ("NN", "DD", "NN"), (2.0, 1.0, 1.0), 0.0),
    ("p2b", (256, 512, 8), ("DD", "NN", "DD"), (2.0, 1.0, 1.0), 0.0),
    ("p2c", (128, 1024, 4), ("PP", "DD", "NN"), (2.0, 1.0, 1.0), 0.0),
    ("p2d", (2048, 64, 4), ("NN", "PP", "NN"), (2.0, 1.0, 1.0), 0.0),
    ("p2e", (72, 2048, 4), ("PP", "NN", "NN"), (2.0, 1.0, 1.0), 0.0),
    ("p2f", (1000, 128, 6), ("DD", "PP", "PP"), (2.0, 1.0, 1.0), 0.0),
    # ND / DN pressure BCs: REDFT11 / RODFT11 (src/fft.f90:256-263), register kernels (32..2048) and tile kernels (others)
    ("nd-x", (64, 48, 16), ("ND", "PP", "NN"), (1.0, 2.0, 1.0), 0.0),
    ("dn-y", (32, 128, 8), ("PP", "DN", "DD"), (2.0, 1.0, 1.0), 0.0),
    ("nd-dn", (40, 72, 12), ("DN", "ND", "NN"), (1.0, 1.0, 1.0), 1.0),
    ("nd1024", (1024, 32, 4), ("ND", "DN", "PP"), (4.0, 1.0, 1.0), 0.0),
    ("dn2048", (16, 2048, 4), ("NN", "DN", "NN"), (1.0, 4.0, 1.0), 0.0),
    # exactly uniform z grids (lz/nz a binary fraction): the shared-LU z kernel (thomas_uni.cuh), L = 32 / 16 / 32 periodic
    ("uni1024", (16, 24, 1024), ("PP", "PP", "NN"), (1.0, 1.0, 1.0), 0.0),
    ("uni512", (40, 16, 512), ("NN", "PP", "DD"), (1.0, 1.0, 2.0), 0.0),
    ("uni1024p", (16, 16, 1024), ("PP", "PP", "PP"), (1.0, 1.0, 4.0), 0.0),
]


@pytest.mark.parametrize("name,ng,cbc,lengths,gr", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("device,generic_fft", [(False, 0), (True, 0), (True, 1), (True, 2)],
                         ids=["hostptr", "devptr", "devptr-genericfft", "devptr-p2fft"])
def test_solver_matches_oracle(name, ng, cbc, lengths, gr, device, generic_fft):
    case = Case(ng, cbc, lengths, gr=gr, seed=4242 + len(name), name=name)
    s = case.setup
    u, v, w = case.velocity()
    rhs_p = case.new_p()
    oracle.fillps(ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, rhs_p)
    rhs = rhs_p[1:-1, 1:-1, 1:-1].copy(order="F")
    pref = rhs_p.copy(order="F")
    oracle.Solver(ng, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], pref)
    p = _solve_gpu(case, rhs, device=device, generic_fft=generic_fft)
    err = gauge_rel_err(p[1:-1, 1:-1, 1:-1], pref[1:-1, 1:-1, 1:-1], case.singular)
    assert err <= TOL, err


@pytest.mark.parametrize("name,ng,cbc,lengths,gr", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ref_tol", [0.0, 1.0e300], ids=["ref-off", "ref-all"])
def test_reference_order_columns(name, ng, cbc, lengths, gr, ref_tol):
    """thomas_ref.cuh: the z stage with the reference-order fix-up switched off and with EVERY column routed through it
    (up to its cap of 8192 columns) must both meet the parity bar on these (well-conditioned) cases."""
    case = Case(ng, cbc, lengths, gr=gr, seed=4242 + len(name), name=name)
    s = case.setup
    u, v, w = case.velocity()
    rhs_p = case.new_p()
    oracle.fillps(ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, rhs_p)
    rhs = rhs_p[1:-1, 1:-1, 1:-1].copy(order="F")
    pref = rhs_p.copy(order="F")
    oracle.Solver(ng, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], pref)
    lib.check(lib.load().flutas_b200_debug_ref_tol(ref_tol))
    try:
        p = _solve_gpu(case, rhs, device=True)
    finally:
        lib.load().flutas_b200_debug_ref_tol(1.0e-5)
    err = gauge_rel_err(p[1:-1, 1:-1, 1:-1], pref[1:-1, 1:-1, 1:-1], case.singular)
    assert err <= TOL, err


@pytest.mark.parametrize("nh_u", [1, 3])
@pytest.mark.parametrize("cbc", [("PP", "PP", "PP"), ("PP", "PP", "NN"), ("NN", "NN", "NN"), ("DD", "NN", "PP")],
                         ids=lambda c: "".join(c))
def test_pressure_step_stencils_bit_exact_and_divergence_free(cbc, nh_u):
    case = Case((48, 40, 24), cbc, (2.0, 1.0, 1.0), rho0=0.1, gr=(1.5 if cbc[2] != "PP" else 0.0), nh_u=nh_u, seed=99)
    s = case.setup
    n = case.ng
    u, v, w = case.velocity()
    uo, vo, wo = u.copy(order="F"), v.copy(order="F"), w.copy(order="F")
    # fillps: bit-exact
    p = case.new_p()
    po = case.new_p()
    api.fillps(*n, case.nh_d, nh_u, *s.dli, s.dzfi, case.dti, case.rho0, u, v, w, p)
    oracle.fillps(n, case.nh_d, nh_u, s.dli, s.dzfi, case.dti, case.rho0, uo, vo, wo, po)
    assert np.array_equal(p, po)
    api.updt_rhs_b(*n, cbc, s.rhsbx, s.rhsby, s.rhsbz, p)
    oracle.updt_rhs_b(n, s.rhsbx, s.rhsby, s.rhsbz, po)
    assert np.array_equal(p, po)
    rhs = p[1:-1, 1:-1, 1:-1].copy()
    # solver
    pl, nf = api.fftini(n, n, (cbc[0], cbc[1]))
    api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, cbc[2], "ccc", p)
    oracle.Solver(n, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], po)
    assert gauge_rel_err(p[1:-1, 1:-1, 1:-1], po[1:-1, 1:-1, 1:-1], case.singular) <= TOL
    case.boundp(p)
    # correc: bit-exact on identical p
    u2, v2, w2 = u.copy(order="F"), v.copy(order="F"), w.copy(order="F")
    api.correc(*n, case.nh_d, nh_u, *s.dli, s.dzci, case.dt, case.rho0, p, u, v, w)
    oracle.correc(n, case.nh_d, nh_u, s.dli, s.dzci, case.dt, case.rho0, p, u2, v2, w2)
    assert np.array_equal(u, u2) and np.array_equal(v, v2) and np.array_equal(w, w2)
    case.correct_dirichlet_faces(p, u, v, w)
    case.refresh_velocity_halos(u, v, w)
    divtot, divmax = api.chkdiv(*n, *s.dli, case.nh_d, nh_u, s.dzfi, u, v, w)
    otot, omax = oracle.chkdiv(n, s.dli, case.nh_d, nh_u, s.dzfi, u, v, w)
    assert divmax == omax
    assert abs(divtot - otot) <= 1e-9 * max(1.0, abs(otot)) + 1e-12
    assert divmax <= 1e-12, divmax
    del rhs
    api.fftend(pl)


def test_device_resident_step_and_linearity():
    """Size-independent properties on a larger grid: linearity of the solve and discrete residual."""
    import torch
    case = Case((256, 128, 64), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), gr=0.0, seed=5)
    s = case.setup
    n = case.ng
    rng = np.random.default_rng(1)
    r1 = rng.uniform(-1, 1, n)
    r2 = rng.uniform(-1, 1, n)
    r1 -= r1.mean()
    r2 -= r2.mean()
    pl, nf = api.fftini(n, n, (case.cbc[0], case.cbc[1]))
    sols = []
    for r in (r1, r2, 2.0 * r1 - 3.0 * r2):
        p = case.new_p()
        p[1:-1, 1:-1, 1:-1] = r
        pd = api.device_field(p)
        api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, case.cbc[2], "ccc", pd)
        torch.cuda.synchronize()
        sols.append(api.host_field(pd, p.shape))
    comb = 2.0 * sols[0] - 3.0 * sols[1]
    assert gauge_rel_err(sols[2][1:-1, 1:-1, 1:-1], comb[1:-1, 1:-1, 1:-1], True) <= 1e-11
    # uniform grid: a zero-mean RHS is compatible, so the discrete residual must vanish
    p = case.boundp(sols[0])
    res = np.max(np.abs(case.laplacian(p) - r1)) / np.max(np.abs(r1))
    assert res <= 1e-10, res
    api.fftend(pl)


@pytest.mark.parametrize("mode", ["nccl", "p2p", "p2p-transpose"])
def test_multi_gpu_slab_solver(mode):
    """N>1: z-slab decomposition, one process per GPU (torchrun), both exchange implementations."""
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 4 if ngpu >= 4 else 2
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "slab_gpu_worker.py"), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0 and "SLAB_OK" in out.stdout
    if mode == "p2p":                                        # the distributed z solve must actually have been exercised
        assert "z=distributed" in out.stdout


FULL_SIZE = [
    ("C2", (512, 512, 512), ("PP", "PP", "PP"), (2 * np.pi,) * 3),
    ("C3", (1024, 512, 512), ("PP", "PP", "NN"), (6.0, 3.0, 1.0)),
    ("C5w1-half", (1024, 512, 512), ("NN", "NN", "NN"), (2.0, 1.0, 1.0)),
]


@pytest.mark.parametrize("name,ng,cbc,lengths", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_residual_and_linearity(name, ng, cbc, lengths):
    """BASELINE-sized grids, where the CPU oracle would take minutes: size-independent properties evaluated on the
    device -- (i) the solution satisfies the discrete 7-point operator the solver inverts (second differences with the
    reference's ghost-cell closures in x,y; the tridiagonal a,b,c of initsolver.f90:188-246 in z), (ii) linearity."""
    import torch
    case = Case(ng, cbc, lengths, gr=0.0, seed=77)
    s = case.setup
    n1, n2, n3 = ng
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    dev = torch.device("cuda")

    def rhs():
        r = torch.rand((n3, n2, n1), dtype=torch.float64, device=dev, generator=g) - 0.5      # [k][j][i] = Fortran (i,j,k)
        return r - r.mean()                                  # compatible with the singular (all-Neumann/periodic) operator

    def solve(r):
        p = torch.zeros((n3 + 2, n2 + 2, n1 + 2), dtype=torch.float64, device=dev)
        p[1:-1, 1:-1, 1:-1] = r
        api.solver(ng, pl, nf, s.lambdaxy, s.a, s.b, s.c, cbc[2], "ccc", p)
        torch.cuda.synchronize()
        return p[1:-1, 1:-1, 1:-1].clone()

    def second_diff(x, dim, bc, dli2):
        # cell-centred ghost closures (bound.f90:247-268): P wrap, N ghost = interior, D ghost = -interior
        lo = torch.roll(x, 1, dim)
        hi = torch.roll(x, -1, dim)
        first = [slice(None)] * 3
        last = [slice(None)] * 3
        first[dim], last[dim] = slice(0, 1), slice(-1, None)
        if bc != "PP":
            sgn = 1.0 if bc == "NN" else -1.0
            lo[tuple(first)] = sgn * x[tuple(first)]
            hi[tuple(last)] = sgn * x[tuple(last)]
        return (lo - 2.0 * x + hi) * dli2

    def apply_operator(x):
        out = second_diff(x, 2, cbc[0], s.dli[0] ** 2)
        out += second_diff(x, 1, cbc[1], s.dli[1] ** 2)
        a = torch.from_numpy(np.ascontiguousarray(s.a)).to(dev)[:, None, None]
        b = torch.from_numpy(np.ascontiguousarray(s.b)).to(dev)[:, None, None]
        c = torch.from_numpy(np.ascontiguousarray(s.c)).to(dev)[:, None, None]
        lo = torch.roll(x, 1, 0)
        hi = torch.roll(x, -1, 0)
        if cbc[2] != "PP":                                   # the wall closures are folded into b (initsolver.f90:228-236)
            lo[0] = 0.0
            hi[-1] = 0.0
        out += a * lo + b * x + c * hi
        return out

    pl, nf = api.fftini(ng, ng, (cbc[0], cbc[1]))
    r1, r2 = rhs(), rhs()
    x1 = solve(r1)
    res = float((apply_operator(x1) - r1).abs().max() / r1.abs().max())
    assert res <= 1e-9, res
    x2 = solve(r2)
    x3 = solve(2.0 * r1 - 3.0 * r2)
    comb = 2.0 * x1 - 3.0 * x2
    d = (x3 - x3.mean()) - (comb - comb.mean())
    lin = float(d.abs().max() / comb.abs().max())
    assert lin <= 1e-11, lin
    print("%s %s: residual %.2e, linearity %.2e" % (name, ng, res, lin))
    api.fftend(pl)
