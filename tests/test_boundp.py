"""boundp (src/bound.f90:146-225): oracle restatement checked against first principles on the CPU, CUDA path
bit-exact against the oracle on the GPU, including edges/corners and non-zero boundary values."""
import numpy as np
import pytest

from flutas_b200.cases import Case
from oracle import oracle

CBCS = [("PP", "PP", "PP"), ("PP", "PP", "NN"), ("NN", "NN", "NN"), ("DD", "NN", "PP"), ("ND", "PP", "DN"), ("PP", "DD", "ND")]


def _field(case, seed):
    n1, n2, n3 = case.ng
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.uniform(-1, 1, (n1 + 2, n2 + 2, n3 + 2)))      # stale halos on purpose


@pytest.mark.parametrize("cbc", CBCS, ids=["/".join(c) for c in CBCS])
def test_oracle_boundp_ghost_cells(cbc):
    case = Case((12, 10, 8), cbc, (2.0, 1.0, 1.0), gr=(0.0 if cbc[2] == "PP" else 1.5), seed=3)
    s = case.setup
    p = _field(case, 1)
    interior = p[1:-1, 1:-1, 1:-1].copy()
    bc = np.array([[0.3, -0.2], [0.1, 0.4], [-0.5, 0.25]])
    for d in range(3):
        if cbc[d] == "PP":
            bc[d] = 0.0
    oracle.boundp(cbc, case.ng, bc, s.nh_d, s.dl, s.dzc, p)
    assert np.array_equal(p[1:-1, 1:-1, 1:-1], interior)
    # face ghosts from first principles (bound.f90:247-268): D -> 2v - p_in, N -> p_in -/+ dr v, P -> wrap
    dr = [(s.dl[0], s.dl[0]), (s.dl[1], s.dl[1]), (s.dzc[s.nh_d - 1], s.dzc[s.nh_d - 1 + case.ng[2]])]
    for d in range(3):
        n = case.ng[d]
        pm = np.moveaxis(p, d, 0)[:, 1:-1, 1:-1]
        if cbc[d] == "PP":
            assert np.array_equal(pm[0], pm[n]) and np.array_equal(pm[n + 1], pm[1])
            continue
        for side, (g, i) in enumerate(((0, 1), (n + 1, n))):
            v = bc[d][side]
            if cbc[d][side] == "D":
                ref = 2.0 * v + (-1.0) * pm[i]
            else:
                ref = (-dr[d][side] * v if side == 0 else dr[d][side] * v) + 1.0 * pm[i]
            assert np.array_equal(pm[g], ref)
    # homogeneous values: identical to the simple ghost rule used elsewhere in the tests
    q = _field(case, 1)
    oracle.boundp(cbc, case.ng, np.zeros((3, 2)), s.nh_d, s.dl, s.dzc, q)
    if all(c in ("PP", "NN", "DD") for c in cbc):
        r = case.boundp(_field(case, 1))
        assert np.array_equal(q[1:-1, 1:-1, :], r[1:-1, 1:-1, :]) and np.array_equal(q[1:-1, :, 1:-1], r[1:-1, :, 1:-1])


def test_oracle_boundp_slabs_match_single_rank():
    """z-slab decomposition with neighbour planes == single-rank result on every rank's window."""
    for cbc in (("PP", "NN", "PP"), ("NN", "PP", "ND")):
        case = Case((6, 8, 12), cbc, (1.0, 1.0, 1.0), seed=2)
        s = case.setup
        bc = np.array([[0.0, 0.0], [0.2, -0.1], [0.3, 0.7]]) * (np.array([c != "PP" for c in cbc])[:, None])
        full = _field(case, 5)
        ref = oracle.boundp(cbc, case.ng, bc, s.nh_d, s.dl, s.dzc, full.copy(order="F"))
        P, n3l = 3, 4
        slabs = [np.asfortranarray(full[:, :, r * n3l:r * n3l + n3l + 2].copy()) for r in range(P)]
        pz = cbc[2] == "PP"
        # what the neighbours send: their boundary planes after their own y-halo update
        sent = []
        for r in range(P):
            t = slabs[r].copy(order="F")
            if cbc[1] == "PP":
                oracle._set_bc(t, "P", 0, 1, 0.0, 0.0)
            sent.append((t[:, :, 1].copy(), t[:, :, n3l].copy()))
        for r in range(P):
            lo = r - 1 if r > 0 else (P - 1 if pz else None)
            hi = r + 1 if r < P - 1 else (0 if pz else None)
            dzc_loc = s.dzc[r * n3l:r * n3l + n3l + 2 * s.nh_d]
            oracle.boundp(cbc, (6, 8, n3l), bc, s.nh_d, s.dl, dzc_loc, slabs[r],
                          below=None if lo is None else sent[lo][1], above=None if hi is None else sent[hi][0],
                          first_rank=(r == 0), last_rank=(r == P - 1))
            assert np.array_equal(slabs[r], ref[:, :, r * n3l:r * n3l + n3l + 2]), (cbc, r)


@pytest.mark.gpu
@pytest.mark.parametrize("cbc", CBCS, ids=["/".join(c) for c in CBCS])
@pytest.mark.parametrize("device", [False, True], ids=["hostptr", "devptr"])
def test_boundp_bit_exact(cbc, device):
    from flutas_b200 import api
    case = Case((40, 24, 18), cbc, (2.0, 1.0, 1.0), gr=(0.0 if cbc[2] == "PP" else 1.5), seed=3)
    s = case.setup
    bc = np.array([[0.3, -0.2], [0.1, 0.4], [-0.5, 0.25]])
    for d in range(3):
        if cbc[d] == "PP":
            bc[d] = 0.0
    for values in (np.zeros((3, 2)), bc):
        ref = oracle.boundp(cbc, case.ng, values, s.nh_d, s.dl, s.dzc, _field(case, 9))
        p = _field(case, 9)
        if device:
            import torch
            pd = api.device_field(p)
            api.boundp(cbc, case.ng, values, s.nh_d, 1, s.dl, s.dzc, s.dzf, pd)
            torch.cuda.synchronize()
            p = api.host_field(pd, p.shape)
        else:
            api.boundp(cbc, case.ng, values, s.nh_d, 1, s.dl, s.dzc, s.dzf, p)
        assert np.array_equal(p, ref)
