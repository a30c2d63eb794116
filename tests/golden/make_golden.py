"""Generates tests/golden/*.npz: small pressure-step problems solved by an implementation that is
independent of both oracle/ and the CUDA path: scipy.fft (pocketfft; rfft/dct/dst share FFTW's r2r
definitions and scaling) + a vectorised numpy restatement of dgtsv_homebrewed / gaussel_periodic
(src/solver_cpu.f90:147-223).  The reference itself (Fortran+MPI+FFTW) cannot run in this image.

    python tests/golden/make_golden.py      # rewrites the fixtures
"""
import os
import sys

import numpy as np
import scipy.fft as sf

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from flutas_b200.cases import Case  # noqa: E402

FWD = {"PP": None, "NN": ("dct", 2), "DD": ("dst", 2), "ND": ("dct", 4), "DN": ("dst", 4)}
BWD = {"PP": None, "NN": ("dct", 3), "DD": ("dst", 3), "ND": ("dct", 4), "DN": ("dst", 4)}


def r2hc(x, axis):
    n = x.shape[axis]
    f = sf.rfft(x, axis=axis)
    f = np.moveaxis(f, axis, 0)
    out = np.empty((n,) + f.shape[1:])
    out[: n // 2 + 1] = f.real
    for k in range(1, (n + 1) // 2):
        out[n - k] = f[k].imag
    return np.moveaxis(out, 0, axis)


def hc2r(x, axis):
    n = x.shape[axis]
    x = np.moveaxis(x, axis, 0)
    f = np.zeros((n // 2 + 1,) + x.shape[1:], dtype=complex)
    f.real = x[: n // 2 + 1]
    for k in range(1, (n + 1) // 2):
        f[k].imag = x[n - k]
    out = sf.irfft(f, n=n, axis=0) * n
    return np.moveaxis(out, 0, axis)


def transform(x, bc, axis, fwd):
    tab = FWD if fwd else BWD
    if tab[bc] is None:
        return r2hc(x, axis) if fwd else hc2r(x, axis)
    fn, t = tab[bc]
    return getattr(sf, fn)(x, type=t, axis=axis, norm=None)


def dgtsv(a, bb, c, p):
    """vectorised over leading axes; p[..., n], bb[..., n]"""
    n = p.shape[-1]
    d = np.zeros_like(p)
    p = p.copy()
    z = 1.0 / bb[..., 0]
    d[..., 0] = c[0] * z
    p[..., 0] = p[..., 0] * z
    for l in range(1, n - 1):
        z = 1.0 / (bb[..., l] - a[l] * d[..., l - 1])
        d[..., l] = c[l] * z
        p[..., l] = (p[..., l] - a[l] * p[..., l - 1]) * z
    z = bb[..., n - 1] - a[n - 1] * d[..., n - 2]
    num = p[..., n - 1] - a[n - 1] * p[..., n - 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        p[..., n - 1] = np.where(z != 0.0, num / z, 0.0)
    for l in range(n - 2, -1, -1):
        p[..., l] = p[..., l] - d[..., l] * p[..., l + 1]
    return p


def solve(case, rhs):
    s = case.setup
    n3 = case.ng[2]
    x = transform(rhs, case.cbc[0], 0, True)
    x = transform(x, case.cbc[1], 1, True)
    bb = s.b[None, None, :] + s.lambdaxy[:, :, None]
    if case.cbc[2] == "PP":
        m = n3 - 1
        p1 = dgtsv(s.a[:m], bb[..., :m], s.c[:m], x[..., :m])
        p2 = np.zeros_like(p1)
        p2[..., 0] = -s.a[0]
        p2[..., m - 1] = -s.c[m - 1]
        p2 = dgtsv(s.a[:m], bb[..., :m], s.c[:m], p2)
        with np.errstate(divide="ignore", invalid="ignore"):
            pn = (x[..., m] - s.c[m] * p1[..., 0] - s.a[m] * p1[..., m - 1]) / \
                 (bb[..., m] + s.c[m] * p2[..., 0] + s.a[m] * p2[..., m - 1])
        x = np.concatenate([p1 + p2 * pn[..., None], pn[..., None]], axis=-1)
    else:
        x = dgtsv(s.a, bb, s.c, x)
    x = transform(x, case.cbc[1], 1, False)
    x = transform(x, case.cbc[0], 0, False)
    return x * s.normfft


SPECS = [
    ("ppp_16x12x10", (16, 12, 10), ("PP", "PP", "PP"), (2 * np.pi,) * 3, 0.0, 1.0),
    ("ppn_16x12x10", (16, 12, 10), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), 0.0, 1.0),
    ("ppn_stretch_12x8x14", (12, 8, 14), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), 2.0, 1.0),
    ("nnn_8x16x12", (8, 16, 12), ("NN", "NN", "NN"), (2.0, 2.0, 1.0), 0.0, 1.0),
    ("nnd_10x12x8", (10, 12, 8), ("NN", "NN", "DD"), (1.0, 1.0, 1.0), 0.0, 1.0),
    ("ddn_12x10x8", (12, 10, 8), ("DD", "NN", "NN"), (1.0, 2.0, 1.0), 1.5, 0.1),
    ("ndp_8x6x10", (8, 6, 10), ("ND", "PP", "NN"), (1.0, 1.0, 1.0), 0.0, 1.0),
    ("pdn_6x8x12", (6, 8, 12), ("PP", "DN", "DD"), (1.0, 1.0, 2.0), 0.0, 1.0),
    ("pnp_16x18x6", (16, 18, 6), ("PP", "NN", "PP"), (2.0, 1.0, 1.0), 0.0, 1.0),   # coarse_two_layer_rb layout
    ("ppn_32x6x72", (32, 6, 72), ("PP", "PP", "NN"), (6.0, 3.0, 1.0), 0.0, 1.0),   # 72 = 2^3 3^2 levels
]


def main():
    for i, (name, ng, cbc, l, gr, rho0) in enumerate(SPECS):
        case = Case(ng, cbc, l, rho0=rho0, gr=gr, seed=777 + i, name=name)
        u, v, w = case.velocity()
        s = case.setup
        o = s.nh_d - 1
        k = np.arange(1, ng[2] + 1)
        h = case.nh_u
        rhs = ((w[h:-h, h:-h, h:-h] - w[h:-h, h:-h, h - 1:-h - 1]) * case.dti * s.dzfi[k + o][None, None, :]
               + (v[h:-h, h:-h, h:-h] - v[h:-h, h - 1:-h - 1, h:-h]) * (case.dti * s.dli[1])
               + (u[h:-h, h:-h, h:-h] - u[h - 1:-h - 1, h:-h, h:-h]) * (case.dti * s.dli[0])) * case.rho0
        psol = solve(case, rhs)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ng=np.array(ng), cbc=np.array(cbc),
                            lengths=np.array(l), gr=gr, rho0=rho0, seed=777 + i, rhs=rhs, p=psol)
        p = case.new_p()
        p[1:-1, 1:-1, 1:-1] = psol
        case.boundp(p)
        res = np.max(np.abs(case.laplacian(p) - rhs)) / np.max(np.abs(rhs))
        print("%-24s residual %.2e" % (name, res))


if __name__ == "__main__":
    main()
