"""bounduvw and the chkdt reduction on the device against the numpy oracle
(oracle.bounduvw / oracle.chkdt_dti: restatement of src/bound.f90:17-144, 227-646, 649-773 and src/chkdt.f90:150-173).
Integer/byte-class work: the bar is BIT-EXACT, halos, edges and corners included."""
import numpy as np
import pytest

from flutas_b200 import api
from oracle import oracle

pytestmark = pytest.mark.gpu


def _cbc(x, y, z):
    """cbc[ibound][idir][field] from per-direction (lo, hi) types applied to all three components"""
    return [[[d[ib]] * 3 for d in (x, y, z)] for ib in (0, 1)]


def _fields(n, nh, seed):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.uniform(-1, 1, tuple(x + 2 * nh for x in n))) for _ in range(3)]


def _grid(nz, nh_d, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.05, 0.15, nz + 2 * nh_d), rng.uniform(0.05, 0.15, nz + 2 * nh_d)


CASES = [
    ("channel", ("PP", "PP", "DD"), {(1, 2, 0): 2.5}, None),
    ("triperiodic", ("PP", "PP", "PP"), {}, None),
    ("cavity", ("DD", "DD", "DD"), {(1, 2, 0): 1.0, (0, 0, 1): -0.3}, None),
    ("neumann-values", ("NN", "DN", "ND"), {(0, 1, 0): 0.7, (1, 2, 1): -0.4, (0, 2, 2): 0.2, (1, 0, 2): 0.9}, None),
    ("outflow-x+", ("DN", "NN", "DD"), {(0, 0, 0): 1.0}, (1, 0)),
    ("outflow-x-", ("ND", "PP", "DD"), {}, (0, 0)),
    ("outflow-y+", ("PP", "DN", "NN"), {(0, 1, 1): 0.5}, (1, 1)),
    ("outflow-y-", ("DD", "ND", "PP"), {}, (0, 1)),
    ("outflow-z+", ("PP", "PP", "DN"), {(0, 2, 2): 0.25}, (1, 2)),
    ("outflow-z-", ("PP", "DD", "ND"), {}, (0, 2)),
]


@pytest.mark.parametrize("name,types,values,outflow", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("nh", [1, 3])
@pytest.mark.parametrize("device", [False, True], ids=["hostptr", "devptr"])
def test_bounduvw_bit_exact(name, types, values, outflow, nh, device):
    n, nh_d = (12, 9, 10), 3
    u, v, w = _fields(n, nh, 7 + nh)
    dzc, dzf = _grid(n[2], nh_d, 3)
    dl = (0.1, 0.125, 0.2)
    cbc = _cbc(*types)
    bc = [[[0.0] * 3 for _ in range(3)] for _ in range(2)]
    for (ib, idir, fld), val in values.items():
        bc[ib][idir][fld] = val
    iso = [[False] * 3, [False] * 3]
    if outflow is not None:
        iso[outflow[0]][outflow[1]] = True
    ro = [f.copy(order="F") for f in (u, v, w)]
    oracle.bounduvw(cbc, n, bc, nh_d, nh, iso, dl, dzc, dzf, *ro)
    if device:
        import torch
        d = [api.device_field(f) for f in (u, v, w)]
        api.bounduvw(cbc, n, bc, nh_d, nh, iso, dl, dzc, dzf, *d)
        torch.cuda.synchronize()
        got = [api.host_field(t, f.shape) for t, f in zip(d, (u, v, w))]
    else:
        got = [f.copy(order="F") for f in (u, v, w)]
        api.bounduvw(cbc, n, bc, nh_d, nh, iso, dl, dzc, dzf, *got)
    for g, r, nm in zip(got, ro, "uvw"):
        assert np.array_equal(g, r), (name, nm, np.argwhere(g != r)[:5])


@pytest.mark.parametrize("nh_u", [1, 3])
@pytest.mark.parametrize("n", [(16, 12, 10), (70, 33, 5)])
def test_chkdt_reduction_bit_exact(nh_u, n):
    nh_d = 3
    u, v, w = _fields(n, nh_u, 11 + nh_u)
    rng = np.random.default_rng(2)
    dzci = rng.uniform(5.0, 9.0, n[2] + 2 * nh_d)
    dzfi = rng.uniform(5.0, 9.0, n[2] + 2 * nh_d)
    dli = (7.0, 6.0, 8.0)
    ref = oracle.chkdt_dti(n, dli, nh_d, nh_u, dzci, dzfi, u, v, w)
    assert api.chkdt(*n, *dli, nh_d, nh_u, dzci, dzfi, u, v, w) == ref
    d = [api.device_field(f) for f in (u, v, w)]
    assert api.chkdt(*n, *dli, nh_d, nh_u, dzci, dzfi, *d) == ref


def test_bounduvw_rejects_bad_arguments():
    from flutas_b200 import lib
    n = (8, 8, 8)
    u, v, w = _fields(n, 1, 0)
    dzc, dzf = _grid(8, 1, 0)
    cbc = _cbc("PP", "PP", "DD")
    bc = [[[0.0] * 3 for _ in range(3)] for _ in range(2)]
    iso = [[False] * 3, [False] * 3]
    bad = _cbc("PP", "PX", "DD")
    with pytest.raises(lib.FlutasB200Error, match="bad boundary type"):
        api.bounduvw(bad, n, bc, 1, 1, iso, (0.1, 0.1, 0.1), dzc, dzf, u, v, w)
    u9, v9, w9 = _fields(n, 9, 0)
    dzc9, dzf9 = _grid(8, 9, 0)
    with pytest.raises(lib.FlutasB200Error, match="halo width"):
        api.bounduvw(cbc, n, bc, 9, 9, iso, (0.1, 0.1, 0.1), dzc9, dzf9, u9, v9, w9)
