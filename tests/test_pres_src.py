"""Pressure-gradient source terms and pressure bookkeeping next to the solve:
pres_sp_src (src/source.f90:311-346), pres_tw_src (:247-309, constant-coefficient branch), pold = p / p = pold + p
(main__single_phase.f90:693-699, 734-740).  CPU: the oracle restatement against an independent numpy evaluation of
the Fortran expressions; GPU: the CUDA kernels bit-exact against the oracle."""
import numpy as np
import pytest

from flutas_b200.cases import Case
from oracle import oracle


def _fields(case, seed):
    rng = np.random.default_rng(seed)
    n1, n2, n3 = case.ng
    u, v, w = case.velocity()
    p = np.asfortranarray(rng.uniform(-1, 1, (n1 + 2, n2 + 2, n3 + 2)))
    pold = np.asfortranarray(rng.uniform(-1, 1, (n1 + 2, n2 + 2, n3 + 2)))
    rho = np.asfortranarray(rng.uniform(0.5, 20.0, (n1 + 2, n2 + 2, n3 + 2)))
    return u, v, w, p, pold, rho


def _interior(f, h):
    return f[h:-h, h:-h, h:-h]


@pytest.mark.parametrize("nh_u", [1, 3])
def test_oracle_pres_src_matches_numpy(nh_u):
    case = Case((12, 10, 8), ("PP", "PP", "NN"), (2.0, 1.0, 1.0), gr=1.5, nh_u=nh_u, seed=5)
    s = case.setup
    n = case.ng
    u, v, w, p, pold, rho = _fields(case, 11)
    f_t12, f_t12_o, rho0i = 0.7e-3, 1.1e-3, 1.0 / 0.8
    dzci_int = s.dzci[case.nh_d:case.nh_d + n[2]]            # dzci(1:n3)
    # ---- single phase
    u1, v1, w1 = (x.copy(order="F") for x in (u, v, w))
    oracle.pres_sp_src(n, f_t12, s.dli, case.nh_d, nh_u, s.dzci, rho0i, pold, u1, v1, w1)
    c = pold[1:-1, 1:-1, 1:-1]
    eu = _interior(u, nh_u) + f_t12 * (-((pold[2:, 1:-1, 1:-1] - c) * s.dli[0])) * rho0i
    ev = _interior(v, nh_u) + f_t12 * (-((pold[1:-1, 2:, 1:-1] - c) * s.dli[1])) * rho0i
    ew = _interior(w, nh_u) + f_t12 * (-((pold[1:-1, 1:-1, 2:] - c) * dzci_int[None, None, :])) * rho0i
    assert np.array_equal(_interior(u1, nh_u), eu) and np.array_equal(_interior(v1, nh_u), ev)
    assert np.array_equal(_interior(w1, nh_u), ew)
    # halos untouched
    u1[nh_u:-nh_u, nh_u:-nh_u, nh_u:-nh_u] = _interior(u, nh_u)
    assert np.array_equal(u1, u)
    # ---- two phase, constant-coefficient split
    u2, v2, w2 = (x.copy(order="F") for x in (u, v, w))
    oracle.pres_tw_src(n, s.dli, case.nh_d, nh_u, s.dzci, rho0i, f_t12, f_t12_o, p, pold, rho, u2, v2, w2)
    f1, f2 = 1.0 + (f_t12 / f_t12_o), (f_t12 / f_t12_o)
    ext = f1 * p - f2 * pold
    pc, ec, rc = p[1:-1, 1:-1, 1:-1], ext[1:-1, 1:-1, 1:-1], rho[1:-1, 1:-1, 1:-1]
    for vel, got, sl, dl in ((u, u2, np.s_[2:, 1:-1, 1:-1], s.dli[0]), (v, v2, np.s_[1:-1, 2:, 1:-1], s.dli[1]),
                             (w, w2, np.s_[1:-1, 1:-1, 2:], dzci_int[None, None, :])):
        rhoi = 1.0 / (0.5 * (rho[sl] + rc))
        exp = _interior(vel, nh_u) + f_t12 * ((-((p[sl] - pc) * dl)) * rho0i - (rhoi - rho0i) * (ext[sl] - ec) * dl)
        assert np.array_equal(_interior(got, nh_u), exp)
    # ---- bookkeeping
    p3, o3 = p.copy(order="F"), pold.copy(order="F")
    oracle.pold_update(n, 0, p3, o3)
    assert np.array_equal(o3[1:-1, 1:-1, 1:-1], p[1:-1, 1:-1, 1:-1]) and np.array_equal(o3[0], pold[0]) and np.array_equal(p3, p)
    oracle.pold_update(n, 1, p3, o3)
    assert np.array_equal(p3[1:-1, 1:-1, 1:-1], p[1:-1, 1:-1, 1:-1] + p[1:-1, 1:-1, 1:-1]) and np.array_equal(p3[:, 0], p[:, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("nh_u", [1, 3])
@pytest.mark.parametrize("device", [False, True], ids=["hostptr", "devptr"])
def test_gpu_pres_src_bit_exact(nh_u, device):
    import torch
    from flutas_b200 import api
    api.init(0)
    case = Case((70, 37, 18), ("PP", "NN", "NN"), (2.0, 1.0, 1.0), gr=1.0, nh_u=nh_u, seed=8)
    s = case.setup
    n = case.ng
    u, v, w, p, pold, rho = _fields(case, 3)
    f_t12, f_t12_o, rho0i = 0.7e-3, 1.1e-3, 1.0 / 0.8
    up = (lambda f: api.device_field(f)) if device else (lambda f: f.copy(order="F"))
    down = (lambda d, ref: api.host_field(d, ref.shape)) if device else (lambda d, ref: d)

    uo, vo, wo = (x.copy(order="F") for x in (u, v, w))
    oracle.pres_sp_src(n, f_t12, s.dli, case.nh_d, nh_u, s.dzci, rho0i, pold, uo, vo, wo)
    ud, vd, wd, od = up(u), up(v), up(w), up(pold)
    api.pres_sp_src(*n, f_t12, *s.dli, case.nh_d, nh_u, s.dzci, rho0i, od, ud, vd, wd)
    torch.cuda.synchronize()
    for d, o in ((ud, uo), (vd, vo), (wd, wo)):
        assert np.array_equal(down(d, o), o)

    uo, vo, wo = (x.copy(order="F") for x in (u, v, w))
    oracle.pres_tw_src(n, s.dli, case.nh_d, nh_u, s.dzci, rho0i, f_t12, f_t12_o, p, pold, rho, uo, vo, wo)
    ud, vd, wd, pd, od, rd = up(u), up(v), up(w), up(p), up(pold), up(rho)
    api.pres_tw_src(*n, *s.dli, case.nh_d, nh_u, s.dzci, rho0i, f_t12, f_t12_o, pd, od, rd, ud, vd, wd)
    torch.cuda.synchronize()
    for d, o in ((ud, uo), (vd, vo), (wd, wo)):
        assert np.array_equal(down(d, o), o)

    po, oo = p.copy(order="F"), pold.copy(order="F")
    pd, od = up(p), up(pold)
    for mode in (0, 1):
        oracle.pold_update(n, mode, po, oo)
        api.pold_update(*n, mode, pd, od)
        torch.cuda.synchronize()
        assert np.array_equal(down(pd, po), po) and np.array_equal(down(od, oo), oo)
