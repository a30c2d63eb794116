"""ctypes front-end of the CPU oracle (oracle/flutas_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under flutas_b200/ imports this module.

Arrays are numpy float64, Fortran-ordered, shaped like the reference's Fortran arrays:
p(0:n1+1,0:n2+1,0:n3+1), u/v/w(1-nh_u:n+nh_u)^3, dzci/dzfi(1-nh_d:n3+nh_d).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = dict(R2HC=0, HC2R=1, REDFT00=3, REDFT01=4, REDFT10=5, REDFT11=6,
             RODFT00=7, RODFT01=8, RODFT10=9, RODFT11=10)   # src/fftw.f90:41-61

_dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "flutas_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_r2r_create.restype = C.c_void_p
        L.oracle_r2r_create.argtypes = [C.c_int, C.c_int]
        L.oracle_r2r_destroy.argtypes = [C.c_void_p]
        L.oracle_r2r_execute.argtypes = [C.c_void_p, _dp, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long]
        L.oracle_normfft.restype = C.c_double
        L.oracle_normfft.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.oracle_find_fft.argtypes = [C.c_char, C.c_char, C.c_char, C.POINTER(C.c_int), C.POINTER(C.c_int), _dp]
        L.oracle_eigenvalues.argtypes = [C.c_int, C.c_char, C.c_char, _dp]
        L.oracle_tridmatrix.argtypes = [C.c_char, C.c_char, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.oracle_initgrid.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp]
        L.oracle_gaussel.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.oracle_gaussel_periodic.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.oracle_solver_cpu.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_double, _dp, _dp, _dp, _dp,
                                        C.c_char_p, _dp, _dp]
        L.oracle_fillps.argtypes = [C.c_int] * 5 + [C.c_double] * 3 + [_dp, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.oracle_updt_rhs_b.argtypes = [C.c_int] * 3 + [_dp] * 4
        L.oracle_correc.argtypes = [C.c_int] * 5 + [C.c_double] * 3 + [_dp, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.oracle_chkdiv.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 2 + [_dp] * 4 + [_dp, _dp]
        L.oracle_set_definition_path.argtypes = [C.c_int]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.ndim == 1), "need float64 Fortran-contiguous"
    return a.ctypes.data_as(_dp)


def set_definition_path(on):
    """True: every r2r transform goes through the path that transcribes FFTW's definitions (one zero-padded complex FFT);
    False (default): even lengths use the half-length reductions (R2HC split, Makhoul, type-IV twiddles), 1.7x faster --
    tests/test_oracle.py holds the two against each other, scipy/pocketfft and the long-double O(n^2) sums."""
    lib().oracle_set_definition_path(1 if on else 0)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def find_fft(bc, c_or_f="c"):
    kf, kb = C.c_int(), C.c_int()
    norm = np.zeros(2)
    rc = lib().oracle_find_fft(bc[0].encode(), bc[1].encode(), c_or_f.encode(), C.byref(kf), C.byref(kb), _p(norm))
    if rc:
        raise ValueError("unsupported BC pair %r / %r" % (bc, c_or_f))
    return kf.value, kb.value, norm


def normfft(ng1, ng2, bcx, bcy):
    return lib().oracle_normfft(ng1, ng2, (bcx + bcy).encode())


def eigenvalues(n, bc):
    lam = np.zeros(n)
    lib().oracle_eigenvalues(n, bc[0].encode(), bc[1].encode(), _p(lam))
    return lam


def tridmatrix(bcz, n, nh_d, dzci, dzfi):
    a, b, c = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().oracle_tridmatrix(bcz[0].encode(), bcz[1].encode(), n, nh_d, _p(dzci), _p(dzfi), _p(a), _p(b), _p(c))
    return a, b, c


def initgrid(n, gr, lz, nh_d):
    dzc, dzf = np.zeros(n + 2 * nh_d), np.zeros(n + 2 * nh_d)
    lib().oracle_initgrid(n, gr, lz, nh_d, _p(dzc), _p(dzf))
    return dzc, dzf


def r2r(kind, arr, axis):
    """In-place FFTW-r2r-equivalent transform of a Fortran-ordered 3-D array along axis 0 or 1."""
    L = lib()
    n1, n2, n3 = arr.shape
    kind = KINDS[kind] if isinstance(kind, str) else kind
    pl = L.oracle_r2r_create(arr.shape[axis], kind)
    if not pl:
        raise ValueError("unsupported kind")
    try:
        if axis == 0:
            L.oracle_r2r_execute(pl, _p(arr), 1, n2, n1, n3, n1 * n2)
        elif axis == 1:
            L.oracle_r2r_execute(pl, _p(arr), n1, n1, 1, n3, n1 * n2)
        else:
            raise ValueError("axis")
    finally:
        L.oracle_r2r_destroy(pl)
    return arr


class Solver:
    """fftini + solver_cpu on one rank (src/fft.f90:24-157, src/solver_cpu.f90:20-115)."""

    def __init__(self, n, bcx, bcy):
        L = lib()
        self.n = tuple(int(x) for x in n)
        kfx, kbx, _ = find_fft(bcx)
        kfy, kby, _ = find_fft(bcy)
        self.plans = (C.c_void_p * 4)(L.oracle_r2r_create(self.n[0], kfx), L.oracle_r2r_create(self.n[0], kbx),
                                      L.oracle_r2r_create(self.n[1], kfy), L.oracle_r2r_create(self.n[1], kby))
        self.normfft = normfft(self.n[0], self.n[1], bcx, bcy)
        self.work = np.zeros(self.n[0] * self.n[1] * self.n[2])

    def solve(self, lambdaxy, a, b, c, bcz, p):
        n = (C.c_int * 3)(*self.n)
        lib().oracle_solver_cpu(n, self.plans, self.normfft, _p(np.asfortranarray(lambdaxy)), _p(a), _p(b), _p(c),
                                bcz.encode(), _p(p), _p(self.work))
        return p

    def close(self):
        for h in self.plans:
            lib().oracle_r2r_destroy(h)
        self.plans = None

    def __del__(self):
        if getattr(self, "plans", None) is not None:
            try:
                self.close()
            except Exception:
                pass


def fillps(n, nh_d, nh_u, dli, dzfi, dti, rho0, u, v, w, p):
    lib().oracle_fillps(n[0], n[1], n[2], nh_d, nh_u, dli[0], dli[1], dli[2], _p(dzfi), dti, rho0,
                        _p(u), _p(v), _p(w), _p(p))
    return p


def updt_rhs_b(n, rhsbx, rhsby, rhsbz, p):
    lib().oracle_updt_rhs_b(n[0], n[1], n[2], _p(rhsbx), _p(rhsby), _p(rhsbz), _p(p))
    return p


def correc(n, nh_d, nh_u, dli, dzci, dt, rho0, p, u, v, w):
    lib().oracle_correc(n[0], n[1], n[2], nh_d, nh_u, dli[0], dli[1], dli[2], _p(dzci), dt, rho0,
                        _p(p), _p(u), _p(v), _p(w))


def pres_sp_src(n, f_t12, dli, nh_d, nh_u, dzci, rho0i, pold, u, v, w):
    L = lib()
    L.oracle_pres_sp_src.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int] * 2 + [_dp, C.c_double, _dp, _dp, _dp, _dp]
    L.oracle_pres_sp_src(n[0], n[1], n[2], f_t12, dli[0], dli[1], dli[2], nh_d, nh_u, _p(dzci), rho0i, _p(pold), _p(u), _p(v), _p(w))


def pres_tw_src(n, dli, nh_d, nh_u, dzci, rho0i, f_t12, f_t12_o, p, pold, rho, u, v, w):
    L = lib()
    L.oracle_pres_tw_src.argtypes = ([C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 2 + [_dp] + [C.c_double] * 3 +
                                     [_dp] * 6)
    L.oracle_pres_tw_src(n[0], n[1], n[2], dli[0], dli[1], dli[2], nh_d, nh_u, _p(dzci), rho0i, f_t12, f_t12_o,
                         _p(p), _p(pold), _p(rho), _p(u), _p(v), _p(w))


def pold_update(n, mode, p, pold):
    L = lib()
    L.oracle_pold_update.argtypes = [C.c_int] * 4 + [_dp, _dp]
    L.oracle_pold_update(n[0], n[1], n[2], mode, _p(p), _p(pold))


def chkdiv(n, dli, nh_d, nh_u, dzfi, u, v, w):
    tot, mx = C.c_double(), C.c_double()
    lib().oracle_chkdiv(n[0], n[1], n[2], dli[0], dli[1], dli[2], nh_d, nh_u, _p(dzfi), _p(u), _p(v), _p(w),
                        C.byref(tot), C.byref(mx))
    return tot.value, mx.value


def gaussel(a, b, c, lambdaxy, pz, periodic):
    nx, ny, n = pz.shape
    f = lib().oracle_gaussel_periodic if periodic else lib().oracle_gaussel
    f(nx, ny, n, _p(a), _p(b), _p(c), _p(np.asfortranarray(lambdaxy)), _p(pz))
    return pz


# ---------------------------------------------------------------------------------------------------
# boundp (numpy restatement; the arithmetic is one multiply-add per ghost cell, so numpy is bit-exact)
def _set_bc(p, ctype, ibound, idir, rvalue, dr):
    """set_bc, src/bound.f90:227-420, centred, nh_p = 1: ghost = factor + sgn * inner."""
    n = p.shape[idir] - 2
    ghost = [slice(None)] * 3
    inner = [slice(None)] * 3
    if ctype == "P":                                      # bound.f90:268-318 (both sides at once)
        lo, hi, first, last = ([slice(None)] * 3 for _ in range(4))
        lo[idir], hi[idir], first[idir], last[idir] = 0, n + 1, 1, n
        p[tuple(lo)] = p[tuple(last)]
        p[tuple(hi)] = p[tuple(first)]
        return
    factor, sgn = np.float64(rvalue), np.float64(0.0)
    if ctype == "D":                                      # :247-252
        factor, sgn = np.float64(2.0) * factor, np.float64(-1.0)
    if ctype == "N":                                      # :253-264
        factor = -np.float64(dr) * factor if ibound == 0 else np.float64(dr) * factor
        sgn = np.float64(1.0)
    ghost[idir], inner[idir] = (0, 1) if ibound == 0 else (n + 1, n)
    p[tuple(ghost)] = factor + sgn * p[tuple(inner)]      # :334,350,...


def boundp(cbc, n, bc, nh_d, dl, dzc, p, below=None, above=None, first_rank=True, last_rank=True):
    """boundp, src/bound.f90:146-225, for the _DECOMP_X layout with the y direction undivided (z-slabs).

    cbc: three 2-character strings, bc: (3,2) values, dzc(1-nh_d:), p(0:n1+1,0:n2+1,0:n3+1) updated in place.
    Single rank: below = above = None.  On a slab decomposition `below` / `above` are the neighbouring
    ranks' planes p(:,:,n3) / p(:,:,1) *after their own y-halo update* (what MPI_SENDRECV delivers,
    bound.f90:1098-1103), or None where the neighbour is MPI_PROC_NULL."""
    n3 = n[2]
    py, pz = cbc[1] == "PP", cbc[2] == "PP"
    if py:                                                # updthalo, idir = 2: the rank is its own neighbour
        _set_bc(p, "P", 0, 1, 0.0, 0.0)
    single = below is None and above is None and first_rank and last_rank
    if single:
        if pz:                                            # updthalo, idir = 3
            _set_bc(p, "P", 0, 2, 0.0, 0.0)
    else:
        if below is not None:
            p[:, :, 0] = below
        if above is not None:
            p[:, :, n3 + 1] = above
    if cbc[0] == "PP":                                    # x: left = right = MPI_PROC_NULL -> set_bc (initmpi.f90:124)
        _set_bc(p, "P", 0, 0, 0.0, 0.0)
    else:
        _set_bc(p, cbc[0][0], 0, 0, bc[0][0], dl[0])
        _set_bc(p, cbc[0][1], 1, 0, bc[0][1], dl[0])
    if not py:
        _set_bc(p, cbc[1][0], 0, 1, bc[1][0], dl[1])
        _set_bc(p, cbc[1][1], 1, 1, bc[1][1], dl[1])
    if not pz:
        if first_rank:
            _set_bc(p, cbc[2][0], 0, 2, bc[2][0], dzc[nh_d - 1])           # dr = dzc(0)
        if last_rank:
            _set_bc(p, cbc[2][1], 1, 2, bc[2][1], dzc[nh_d - 1 + n3])      # dr = dzc(n3)
    return p


# ---------------------------------------------------------------------------------------------------
# bounduvw: numpy restatement of src/bound.f90:17-144 with set_bc (:227-646) for an arbitrary
# halo width, updthalo (:946-1110) and outflow (:649-773), single rank in the _DECOMP_X layout.  Groundwork for the
# device version: this round ships the oracle and its tests only, no CUDA kernel yet.
def _fx(nh, i):
    """numpy index of Fortran index i of an array dimensioned (1-nh:)"""
    return i + nh - 1


def _plane(idir, idx):
    s = [slice(None)] * 3
    s[idir] = idx
    return tuple(s)


def set_bc_general(p, ctype, ibound, idir, centered, rvalue, dr, nh, n):
    """set_bc(nx,ny,nz,ctype,ibound,idir,centered,rvalue,qq_d,nh_p,dr,p), src/bound.f90:227-646.
    idir 0-based; dr[q], q = 0..nh-1; loops over q run in the reference's order (later q may overwrite earlier ones)."""
    P = lambda i: _plane(idir, _fx(nh, i))
    if ctype == "P":                                      # :268-318
        for q in range(nh):
            p[P(0 - q)] = p[P(n - q)]
            p[P(n + 1 + q)] = p[P(1 + q)]
        return
    factor = [np.float64(rvalue)] * nh
    sgn = np.float64(0.0)
    if ctype == "D" and centered:                         # :251-256
        factor = [np.float64(2.0) * f for f in factor]
        sgn = np.float64(-1.0)
    if ctype == "N":                                      # :257-268
        factor = [(-np.float64(dr[q]) * factor[q]) if ibound == 0 else (np.float64(dr[q]) * factor[q]) for q in range(nh)]
        sgn = np.float64(1.0)
    for q in range(nh):
        f = factor[q]
        if centered:                                      # :320-431
            if ibound == 0:
                p[P(0 - q)] = f + sgn * p[P(1 + q)]
            else:
                p[P(n + 1 + q)] = f + sgn * p[P(n - q)]
        elif ctype == "D":                                # :432-533 (face-centred component normal to the wall)
            if ibound == 0:
                p[P(0 - q)] = f
            else:
                p[P(n + q)] = f
                p[P(n + 1 + q)] = p[P(n - 1 - q)]
        elif ctype == "N":                                # :534-646
            if ibound == 0:
                p[P(0 - q)] = np.float64(1.0) * f + p[P(1 + q)]
            else:
                p[P(n + q)] = np.float64(1.0) * f + p[P(n - 1 - q)]
                p[P(n + 1 + q)] = np.float64(2.0) * f + p[P(n - 1 - q)]


def _outflow(n, idir_signed, nh_d, nh_u, dl, dzf, u, v, w):
    """outflow, src/bound.f90:649-773: face velocity from zero divergence on an outflow boundary"""
    nx, ny, nz = n
    qmin = abs(1 - nh_u)
    dx, dy = dl[0], dl[1]
    dxi, dyi = dx ** (-1), dy ** (-1)
    dzfi = dzf ** (-1)                                     # dzf(1-nh_d:) -> numpy index k + nh_d - 1
    h = nh_u
    I = lambda a, b: slice(_fx(h, a), _fx(h, b) + 1)       # Fortran range a:b
    X = lambda i: _fx(h, i)
    zk = lambda k: k + nh_d - 1
    if idir_signed == 1:                                   # x, right
        i = nx
        dzk = dzfi[zk(1):zk(nz) + 1][None, :]
        for q in range(qmin + 2):
            u[X(i + q), I(1, ny), I(1, nz)] = u[X(i - 1 - q), I(1, ny), I(1, nz)] - dx * (
                (v[X(i + q), I(1, ny), I(1, nz)] - v[X(i + q), I(0, ny - 1), I(1, nz)]) * dyi +
                (w[X(i + q), I(1, ny), I(1, nz)] - w[X(i + q), I(1, ny), I(0, nz - 1)]) * dzk)
    elif idir_signed == 2:                                 # y, back
        j = ny
        dzk = dzfi[zk(1):zk(nz) + 1][None, :]
        for q in range(qmin + 2):
            v[I(1, nx), X(j + q), I(1, nz)] = v[I(1, nx), X(j - 1 - q), I(1, nz)] - dy * (
                (u[I(1, nx), X(j + q), I(1, nz)] - u[I(0, nx - 1), X(j + q), I(1, nz)]) * dxi +
                (w[I(1, nx), X(j + q), I(1, nz)] - w[I(1, nx), X(j + q), I(0, nz - 1)]) * dzk)
    elif idir_signed == 3:                                 # z, top
        k = nz
        for q in range(qmin + 2):
            w[I(1, nx), I(1, ny), X(k + q)] = w[I(1, nx), I(1, ny), X(k - 1 - q)] - dzf[zk(k + q)] * (
                (u[I(1, nx), I(1, ny), X(k + q)] - u[I(0, nx - 1), I(1, ny), X(k + q)]) * dxi +
                (v[I(1, nx), I(1, ny), X(k + q)] - v[I(1, nx), I(0, ny - 1), X(k + q)]) * dyi)
    elif idir_signed == -1:                                # x, left
        i = 0
        dzk = dzfi[zk(1):zk(nz) + 1][None, :]
        for q in range(qmin + 1):
            u[X(i - q), I(1, ny), I(1, nz)] = u[X(i + 1 + q), I(1, ny), I(1, nz)] + dx * (
                (v[X(i + 1 + q), I(1, ny), I(1, nz)] - v[X(i + 1 + q), I(0, ny - 1), I(1, nz)]) * dyi +
                (w[X(i + 1 + q), I(1, ny), I(1, nz)] - w[X(i + 1 + q), I(1, ny), I(0, nz - 1)]) * dzk)
    elif idir_signed == -2:                                # y, front
        j = 0
        dzk = dzfi[zk(1):zk(nz) + 1][None, :]
        for q in range(qmin + 1):
            v[I(1, nx), X(j - q), I(1, nz)] = v[I(1, nx), X(j + 1 + q), I(1, nz)] + dy * (
                (u[I(1, nx), X(j + 1 + q), I(1, nz)] - u[I(0, nx - 1), X(j + 1 + q), I(1, nz)]) * dxi +
                (w[I(1, nx), X(j + 1 + q), I(1, nz)] - w[I(1, nx), X(j + 1 + q), I(0, nz - 1)]) * dzk)
    elif idir_signed == -3:                                # z, bottom
        k = 0
        for q in range(qmin + 1):
            w[I(1, nx), I(1, ny), X(k - q)] = w[I(1, nx), I(1, ny), X(k + 1 + q)] + dzf[zk(k - q)] * (
                (u[I(1, nx), I(1, ny), X(k + 1 + q)] - u[I(0, nx - 1), I(1, ny), X(k + 1 + q)]) * dxi +
                (v[I(1, nx), I(1, ny), X(k + 1 + q)] - v[I(1, nx), I(0, ny - 1), X(k + 1 + q)]) * dyi)


def bounduvw(cbc, n, bc, nh_d, nh_u, isoutflow, dl, dzc, dzf, u, v, w):
    """bounduvw(cbc,n,bc,nh_d,nh_u,halo,isoutflow,dl,dzc,dzf,u,v,w), src/bound.f90:17-144, on ONE rank.
    cbc[ibound][idir][field] (characters), bc likewise (values), isoutflow[ibound][idir]; dzc, dzf dimensioned (1-nh_d:).
    Periodic y / z: the rank is its own neighbour, updthalo wraps nh_u layers (:55-62) and set_bc is skipped because the
    neighbour is not MPI_PROC_NULL; x is never divided in _DECOMP_X, so every x boundary goes through set_bc (:68-79)."""
    nx, ny, nz = n
    fields = (u, v, w)
    per_y = all(cbc[b][1][f] == "P" for b in (0, 1) for f in range(3))
    per_z = all(cbc[b][2][f] == "P" for b in (0, 1) for f in range(3))
    for fld in fields:                                     # updthalo ind1 = 2, ind2 = 3 (:55-62)
        if per_y:
            set_bc_general(fld, "P", 0, 1, True, 0.0, None, nh_u, ny)
        if per_z:
            set_bc_general(fld, "P", 0, 2, True, 0.0, None, nh_u, nz)
    qmin = abs(1 - nh_u)
    zk = lambda k: k + nh_d - 1
    for idir, nn in ((0, nx), (1, ny), (2, nz)):
        if (idir == 1 and per_y) or (idir == 2 and per_z):
            continue
        for ib in (0, 1):
            for f, fld in enumerate(fields):
                centered = (f != idir)                     # the wall-normal component is face-centred (.false. in :68-128)
                if idir < 2:
                    dr = [dl[idir]] * (qmin + 1)
                elif ib == 0:
                    dr = [(dzc if centered else dzf)[zk(-q)] for q in range(qmin + 1)]            # :95-106
                else:
                    dr = [(dzc if centered else dzf)[zk(nz + q)] for q in range(qmin + 1)]        # :108-119
                set_bc_general(fld, cbc[ib][idir][f], ib, idir, centered, bc[ib][idir][f], dr, nh_u, nn)
    # NOTE the reference's order inside one direction is: bottom u, v, w, then top u, v, w for x and y; for z it is bottom
    # u, v (dzc), bottom w (dzf), top u, v, top w -- the same sequence as the loops above.
    for q in range(3):                                     # :131-141
        for ib in (0, 1):
            if isoutflow[ib][q]:
                _outflow(n, (q + 1) * (-1 if ib == 0 else 1), nh_d, nh_u, dl, dzf, u, v, w)
    return u, v, w


def chkdt_dti(n, dli, nh_d, nh_u, dzci, dzfi, u, v, w):
    """The field reduction of chkdt_sp / chkdt_tw, src/chkdt.f90:62-85 = :150-173 (identical in both): the convective
    inverse time scale dti = max over cells of (dtix, dtiy, dtiz).  The scalar formulas that follow (:92-110, :180-190) use
    physical parameters of mod_param and stay on the host.  Oracle groundwork (no CUDA kernel yet)."""
    nx, ny, nz = n
    h = nh_u
    S = lambda a, b: slice(a + h - 1, b + h)               # Fortran range a:b of a (1-nh_u:) dimension
    c = (S(1, nx), S(1, ny), S(1, nz))
    sh = lambda f, di, dj, dk: f[S(1 + di, nx + di), S(1 + dj, ny + dj), S(1 + dk, nz + dk)]
    dzf = np.asarray(dzfi)[nh_d:nh_d + nz][None, None, :]  # dzfi(1:nz)
    dzc = np.asarray(dzci)[nh_d:nh_d + nz][None, None, :]
    q = np.float64(0.25)
    ux = np.abs(u[c])
    vx = q * np.abs(sh(v, 0, 0, 0) + sh(v, 0, -1, 0) + sh(v, 1, 0, 0) + sh(v, 1, -1, 0))
    wx = q * np.abs(sh(w, 0, 0, 0) + sh(w, 0, 0, -1) + sh(w, 1, 0, 0) + sh(w, 1, 0, -1))
    dtix = ux * dli[0] + vx * dli[1] + wx * dzf
    uy = q * np.abs(sh(u, 0, 0, 0) + sh(u, 0, 1, 0) + sh(u, -1, 1, 0) + sh(u, -1, 0, 0))
    vy = np.abs(v[c])
    wy = q * np.abs(sh(w, 0, 0, 0) + sh(w, 0, 1, 0) + sh(w, 0, 1, -1) + sh(w, 0, 0, -1))
    dtiy = uy * dli[0] + vy * dli[1] + wy * dzf
    uz = q * np.abs(sh(u, 0, 0, 0) + sh(u, -1, 0, 0) + sh(u, -1, 0, 1) + sh(u, 0, 0, 1))
    vz = q * np.abs(sh(v, 0, 0, 0) + sh(v, 0, -1, 0) + sh(v, 0, -1, 1) + sh(v, 0, 0, 1))
    wz = np.abs(w[c])
    dtiz = uz * dli[0] + vz * dli[1] + wz * dzc
    return float(max(0.0, dtix.max(), dtiy.max(), dtiz.max()))
