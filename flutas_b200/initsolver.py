"""Host-side set-up of the pressure solver: mirrors FluTAS `initsolver` (src/initsolver.f90:21-120).

The north star keeps `initsolver.f90` unchanged (it stays Fortran, runs once, host only); this module
is its stand-in for a Python host so the C-ABI (`flutas_b200_fftini` / `flutas_b200_solver`) receives
exactly what the Fortran main would hand over: `lambdaxy`, `a,b,c`, `rhsb{x,y,z}`, `normfft`.

Arrays use the reference's shapes and Fortran order; `dzci/dzfi` are 1-D with lower bound 1-nh_d.
"""
import numpy as np


def eigenvalues(n, bc, c_or_f="c"):
    """Modified wavenumbers, src/initsolver.f90:122-186 (CPU branch: FFTW half-complex order for PP)."""
    if c_or_f != "c":
        raise ValueError("only cell-centred ('c') transforms are used by FluTAS (SURVEY.md 8a-3)")
    l = np.arange(1, n + 1, dtype=np.float64)
    pi = np.arccos(-1.0)
    bcs = bc[0] + bc[1]
    if bcs == "PP":
        s = np.sin((1.0 * (l - 1)) * pi / (1.0 * n))
    elif bcs == "NN":
        s = np.sin((1.0 * (l - 1)) * pi / (2.0 * n))
    elif bcs == "DD":
        s = np.sin((1.0 * (l - 0)) * pi / (2.0 * n))
    elif bcs in ("ND", "DN"):
        s = np.sin((1.0 * (2 * l - 1)) * pi / (4.0 * n))
    else:
        raise ValueError("unsupported BC pair %r" % bcs)
    return -4.0 * s * s


def tridmatrix(bcz, n, nh_d, dzci, dzfi):
    """a,b,c of the z operator, src/initsolver.f90:188-246 (cell-centred)."""
    o = nh_d - 1                                  # dzci[k + o] == dzci(k)
    k = np.arange(1, n + 1)
    a = dzfi[k + o] * dzci[k - 1 + o]
    c = dzfi[k + o] * dzci[k + o]
    b = -(a + c)
    fac = {"P": 0.0, "D": -1.0, "N": 1.0}
    b[0] = b[0] + fac[bcz[0]] * a[0]
    b[n - 1] = b[n - 1] + fac[bcz[1]] * c[n - 1]
    return a, b, c


def bc_rhs(cbc, bc, dlc, dlf, shape):
    """Boundary RHS constants, src/initsolver.f90:248-298 (cell-centred). Returns rhs(shape[0],shape[1],0:1)."""
    rhs = np.zeros(shape + (2,), order="F")
    for ib in (0, 1):
        if cbc[ib] == "P":
            factor = 0.0
        elif cbc[ib] == "D":
            factor = -2.0 * bc[ib]
        else:
            sgn = 1.0 if ib == 0 else -1.0
            factor = sgn * dlc[ib] * bc[ib]
        rhs[:, :, ib] = factor / dlc[ib] / dlf[ib]
    return rhs


def initgrid(n, gr, lz, nh_d):
    """dzc, dzf with halos, src/initgrid.f90:17-97 (two-end tanh clustering, :102-118)."""
    o = nh_d - 1
    dzc = np.zeros(n + 2 * nh_d)
    dzf = np.zeros(n + 2 * nh_d)
    zf = np.zeros(n + 2)
    for k in range(1, n + 1):
        z0 = (k - 0.0) / (1.0 * n)
        z = 0.5 * (1.0 + np.tanh((z0 - 0.5) * gr) / np.tanh(gr / 2.0)) if gr != 0.0 else z0
        zf[k] = z * lz
    for k in range(1, n + 1):
        dzf[k + o] = zf[k] - zf[k - 1]
    dzf[0 + o] = dzf[1 + o]
    dzf[n + 1 + o] = dzf[n + o]
    for k in range(0, n + 1):
        dzc[k + o] = 0.5 * (dzf[k + o] + dzf[k + 1 + o])
    dzc[n + 1 + o] = dzc[n + o]
    for k in range(1 - nh_d, 1):
        dzf[k + o] = dzf[-k + 1 + o]
        dzc[k + o] = dzc[-k + o]
    for k in range(n + 1, n + nh_d + 1):
        dzf[k + o] = dzf[2 * n - k - 1 + o]
        dzc[k + o] = dzc[2 * n - k + o]
    return dzc, dzf


def find_fft(bc, c_or_f="c"):
    """BC pair -> (kind_fwd, kind_bwd, norm), src/fft.f90:233-291 (FFTW kind codes of src/fftw.f90:41-61)."""
    if c_or_f != "c":
        raise ValueError("face-centred transforms are dead code in FluTAS")
    table = {"PP": (0, 1, (1.0, 0.0)), "NN": (5, 4, (2.0, 0.0)), "DD": (9, 8, (2.0, 0.0)),
             "ND": (6, 6, (2.0, 0.0)), "DN": (10, 10, (2.0, 0.0))}
    return table[bc[0] + bc[1]]


class SolverSetup:
    """Everything `initsolver` returns for one rank (single-rank window: n_z = ng)."""

    def __init__(self, ng, lengths, cbc, bc=None, gr=0.0, nh_d=1):
        self.ng = tuple(int(x) for x in ng)
        self.cbc = tuple(cbc)                     # ("PP","PP","NN")
        self.bc = bc if bc is not None else ((0.0, 0.0),) * 3
        self.nh_d = nh_d
        self.gr = float(gr)
        n1, n2, n3 = self.ng
        self.dl = (lengths[0] / n1, lengths[1] / n2, lengths[2] / n3)
        self.dli = tuple(1.0 / d for d in self.dl)
        dzc, dzf = initgrid(n3, gr, lengths[2], nh_d)
        self.dzc, self.dzf = dzc, dzf
        self.dzci, self.dzfi = 1.0 / dzc, 1.0 / dzf
        lx = eigenvalues(n1, self.cbc[0]) * self.dli[0] ** 2
        ly = eigenvalues(n2, self.cbc[1]) * self.dli[1] ** 2
        self.lambdaxy = np.asfortranarray(lx[:, None] + ly[None, :])          # initsolver.f90:87-93
        self.a, self.b, self.c = tridmatrix(self.cbc[2], n3, nh_d, self.dzci, self.dzfi)
        o = nh_d - 1
        dl = self.dl
        self.rhsbx = bc_rhs(self.cbc[0], self.bc[0], (dl[0], dl[0]), (dl[0], dl[0]), (n2, n3))
        self.rhsby = bc_rhs(self.cbc[1], self.bc[1], (dl[1], dl[1]), (dl[1], dl[1]), (n1, n3))
        self.rhsbz = bc_rhs(self.cbc[2], self.bc[2], (dzc[0 + o], dzc[n3 + o]), (dzf[1 + o], dzf[n3 + o]), (n1, n2))
        nf = 1.0
        for d, nn in ((0, n1), (1, n2)):
            _, _, norm = find_fft(self.cbc[d])
            nf = nf * norm[0] * (nn + norm[1])
        self.normfft = 1.0 / nf                                               # fft.f90:87,125,150
