"""Oracle groundwork for SURVEY.md 8(f) rank 3: the numpy restatement of bounduvw (src/bound.f90:17-144, set_bc :227-646,
outflow :649-773) checked (i) against a literal, Fortran-indexed loop transcription of set_bc on a tiny grid and
(ii) through the physical meaning of each boundary type.  No CUDA kernel exists for this row yet."""
import numpy as np
import pytest

from oracle import oracle


class FArr:
    """3-D array addressed with Fortran indices (1-nh:) -- for the literal loop transcription"""

    def __init__(self, a, nh):
        self.a, self.nh = a, nh

    def __getitem__(self, ijk):
        return self.a[tuple(x + self.nh - 1 for x in ijk)]

    def __setitem__(self, ijk, v):
        self.a[tuple(x + self.nh - 1 for x in ijk)] = v


def set_bc_loops(n, ctype, ibound, idir, centered, rvalue, dr, nh, arr):
    """bound.f90:227-646 transcribed loop by loop (idir 1-based like the Fortran)"""
    p = FArr(arr, nh)
    nn = n[idir - 1]
    others = [d for d in (1, 2, 3) if d != idir]
    ranges = [range(1 - nh, n[d - 1] + nh + 1) for d in others]

    def idx(i, a, b):
        out = [0, 0, 0]
        out[idir - 1], out[others[0] - 1], out[others[1] - 1] = i, a, b
        return tuple(out)

    factor = [rvalue] * nh
    sgn = 0.0
    if ctype == "D" and centered:
        factor = [2.0 * f for f in factor]
        sgn = -1.0
    if ctype == "N":
        factor = [(-dr[q] * factor[q]) if ibound == 0 else (dr[q] * factor[q]) for q in range(nh)]
        sgn = 1.0
    for q in range(nh):
        for a in ranges[0]:
            for b in ranges[1]:
                if ctype == "P":
                    p[idx(0 - q, a, b)] = p[idx(nn - q, a, b)]
                    p[idx(nn + 1 + q, a, b)] = p[idx(1 + q, a, b)]
                elif centered:
                    if ibound == 0:
                        p[idx(0 - q, a, b)] = factor[q] + sgn * p[idx(1 + q, a, b)]
                    else:
                        p[idx(nn + 1 + q, a, b)] = factor[q] + sgn * p[idx(nn - q, a, b)]
                elif ctype == "D":
                    if ibound == 0:
                        p[idx(0 - q, a, b)] = factor[q]
                    else:
                        p[idx(nn + q, a, b)] = factor[q]
                        p[idx(nn + 1 + q, a, b)] = p[idx(nn - 1 - q, a, b)]
                elif ctype == "N":
                    if ibound == 0:
                        p[idx(0 - q, a, b)] = 1.0 * factor[q] + p[idx(1 + q, a, b)]
                    else:
                        p[idx(nn + q, a, b)] = 1.0 * factor[q] + p[idx(nn - 1 - q, a, b)]
                        p[idx(nn + 1 + q, a, b)] = 2.0 * factor[q] + p[idx(nn - 1 - q, a, b)]


@pytest.mark.parametrize("nh", [1, 3])
@pytest.mark.parametrize("ctype", ["P", "D", "N"])
@pytest.mark.parametrize("centered", [True, False])
@pytest.mark.parametrize("idir", [1, 2, 3])
@pytest.mark.parametrize("ibound", [0, 1])
def test_set_bc_matches_loop_transcription(nh, ctype, centered, idir, ibound):
    n = (5, 4, 6)
    rng = np.random.default_rng(nh + 10 * idir + ibound)
    a = np.asfortranarray(rng.uniform(-1, 1, tuple(x + 2 * nh for x in n)))
    b = a.copy(order="F")
    dr = list(rng.uniform(0.1, 0.3, nh))
    oracle.set_bc_general(a, ctype, ibound, idir - 1, centered, 0.37, dr, nh, n[idir - 1])
    set_bc_loops(n, ctype, ibound, idir, centered, 0.37, dr, nh, b)
    assert np.array_equal(a, b)


def _cbc(x, y, z):
    """cbc[ibound][idir][field] from per-direction (lo, hi) types applied to all three components"""
    return [[[d[ib]] * 3 for d in (x, y, z)] for ib in (0, 1)]


def test_channel_walls_and_periodic_wrap():
    """turbulent-channel velocity BCs: periodic x, y; no-slip walls in z (u, v centred D; w face-centred D)"""
    n, nh, nh_d = (8, 6, 10), 3, 3
    rng = np.random.default_rng(4)
    u, v, w = (np.asfortranarray(rng.uniform(-1, 1, tuple(x + 2 * nh for x in n))) for _ in range(3))
    dzc = np.full(n[2] + 2 * nh_d, 0.1)
    dzf = np.full(n[2] + 2 * nh_d, 0.1)
    cbc = _cbc("PP", "PP", "DD")
    bc = [[[0.0] * 3 for _ in range(3)] for _ in range(2)]
    bc[1][2][0] = 2.5                                          # moving top wall: u = 2.5 at z = lz
    oracle.bounduvw(cbc, n, bc, nh_d, nh, [[False] * 3, [False] * 3], (0.1, 0.1, 0.1), dzc, dzf, u, v, w)
    X = lambda i: i + nh - 1
    for f in (u, v, w):                                        # periodic wrap of all nh layers, full extent of the other dims
        for q in range(nh):
            assert np.array_equal(f[X(0 - q)], f[X(n[0] - q)]) and np.array_equal(f[X(n[0] + 1 + q)], f[X(1 + q)])
    # y was wrapped BEFORE the z walls were set, so compare on the z interior only
    zi = slice(X(1), X(n[2]) + 1)
    for f in (u, v, w):
        for q in range(nh):
            assert np.array_equal(f[:, X(0 - q), zi], f[:, X(n[1] - q), zi])
    for q in range(nh):                                        # centred Dirichlet: the wall value is the mean of ghost and mirror cell
        assert np.allclose(0.5 * (u[:, :, X(0 - q)] + u[:, :, X(1 + q)]), 0.0, atol=1e-15)
        assert np.allclose(0.5 * (u[:, :, X(n[2] + 1 + q)] + u[:, :, X(n[2] - q)]), 2.5, atol=1e-15)
    assert np.all(w[:, :, X(0)] == 0.0) and np.all(w[:, :, X(n[2])] == 0.0)      # no penetration on the wall faces
    assert np.array_equal(w[:, :, X(n[2] + nh)], w[:, :, X(n[2] - nh)])          # after the q loop: p(n+nh) = p(n-nh)


def test_neumann_gradient_and_outflow_is_divergence_free():
    n, nh, nh_d = (6, 5, 4), 1, 1
    rng = np.random.default_rng(9)
    u, v, w = (np.asfortranarray(rng.uniform(-1, 1, tuple(x + 2 * nh for x in n))) for _ in range(3))
    dl = (0.2, 0.25, 0.5)
    dzc = np.full(n[2] + 2 * nh_d, dl[2])
    dzf = np.full(n[2] + 2 * nh_d, dl[2])
    cbc = _cbc("DN", "NN", "DD")                               # inflow (D) at x = 0, zero-gradient outflow at x = lx
    bc = [[[0.0] * 3 for _ in range(3)] for _ in range(2)]
    bc[0][1][0] = 0.7                                          # du/dy = 0.7 on the front wall (centred N for u)
    iso = [[False] * 3, [True, False, False]]                  # outflow on the right x boundary
    oracle.bounduvw(cbc, n, bc, nh_d, nh, iso, dl, dzc, dzf, u, v, w)
    X = lambda i: i + nh - 1
    xi, zi = slice(X(1), X(n[0]) + 1), slice(X(1), X(n[2]) + 1)
    # the y ghost is set after x and before z: check on the interior of the other two directions; the outflow step then
    # rewrites u(nx,1:ny,1:nz) (bound.f90:131-141 run last), so the last x row is excluded
    xi = slice(X(1), X(n[0] - 1) + 1)
    assert np.allclose((u[xi, X(1), zi] - u[xi, X(0), zi]) / dl[1], 0.7, atol=1e-13)
    # outflow: the cell next to the boundary is divergence free with the new face velocity u(nx)
    i = n[0]
    ys, zs = slice(X(1), X(n[1]) + 1), slice(X(1), X(n[2]) + 1)
    div = ((u[X(i), ys, zs] - u[X(i - 1), ys, zs]) / dl[0] +
           (v[X(i), ys, zs] - v[X(i), X(0):X(n[1] - 1) + 1, zs]) / dl[1] +
           (w[X(i), ys, zs] - w[X(i), ys, X(0):X(n[2] - 1) + 1]) / dl[2])
    assert np.max(np.abs(div)) <= 1e-13


@pytest.mark.parametrize("nh_u", [1, 3])
def test_chkdt_reduction_matches_loop_transcription(nh_u):
    """chkdt.f90:62-85: the numpy statement against the Fortran loop written out cell by cell"""
    n, nh_d = (5, 4, 6), 3
    rng = np.random.default_rng(nh_u)
    u, v, w = (np.asfortranarray(rng.uniform(-1, 1, tuple(x + 2 * nh_u for x in n))) for _ in range(3))
    dzci = rng.uniform(5.0, 9.0, n[2] + 2 * nh_d)
    dzfi = rng.uniform(5.0, 9.0, n[2] + 2 * nh_d)
    dli = (7.0, 6.0, 8.0)
    U, V, W = FArr(u, nh_u), FArr(v, nh_u), FArr(w, nh_u)
    zc = lambda k: dzci[k + nh_d - 1]
    zf = lambda k: dzfi[k + nh_d - 1]
    dti = 0.0
    for k in range(1, n[2] + 1):
        for j in range(1, n[1] + 1):
            for i in range(1, n[0] + 1):
                ux = abs(U[i, j, k])
                vx = 0.25 * abs(V[i, j, k] + V[i, j - 1, k] + V[i + 1, j, k] + V[i + 1, j - 1, k])
                wx = 0.25 * abs(W[i, j, k] + W[i, j, k - 1] + W[i + 1, j, k] + W[i + 1, j, k - 1])
                dtix = ux * dli[0] + vx * dli[1] + wx * zf(k)
                uy = 0.25 * abs(U[i, j, k] + U[i, j + 1, k] + U[i - 1, j + 1, k] + U[i - 1, j, k])
                vy = abs(V[i, j, k])
                wy = 0.25 * abs(W[i, j, k] + W[i, j + 1, k] + W[i, j + 1, k - 1] + W[i, j, k - 1])
                dtiy = uy * dli[0] + vy * dli[1] + wy * zf(k)
                uz = 0.25 * abs(U[i, j, k] + U[i - 1, j, k] + U[i - 1, j, k + 1] + U[i, j, k + 1])
                vz = 0.25 * abs(V[i, j, k] + V[i, j - 1, k] + V[i, j - 1, k + 1] + V[i, j, k + 1])
                wz = abs(W[i, j, k])
                dtiz = uz * dli[0] + vz * dli[1] + wz * zc(k)
                dti = max(dti, dtix, dtiy, dtiz)
    assert oracle.chkdt_dti(n, dli, nh_d, nh_u, dzci, dzfi, u, v, w) == dti
