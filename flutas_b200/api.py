"""Host-side mirror of the reference's operator interface for the pressure path, on top of the C ABI.

Same names, argument order and meaning as the Fortran module procedures:
    fftini   src/fft.f90:24        fftend  src/fft.f90:159       solver  src/solver_cpu.f90:20
    fillps   src/fillps.f90:16     correc  src/correc.f90:16     chkdiv  src/chkdiv.f90:18
    updt_rhs_b  src/bound.f90:829
Fields may be numpy arrays (host memory, Fortran order; staged through the device inside the call) or
torch CUDA tensors holding the same Fortran-ordered storage (used in place, asynchronously on torch's
current stream).  torch is only used for device memory and streams.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

_dp = C.POINTER(C.c_double)


def _ptr(x):
    """address of the first element of a numpy array or torch tensor"""
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64:
            raise TypeError("float64 required")
        if not (x.flags.f_contiguous or x.ndim <= 1):
            raise ValueError("Fortran-contiguous array required")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            raise TypeError("contiguous float64 tensor required")
        return C.c_void_p(x.data_ptr())
    raise TypeError("unsupported array type %r" % type(x))


def _use_torch_stream():
    try:
        import torch
        if torch.cuda.is_available():
            _lib.load().flutas_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    except ImportError:
        pass


def init(device=0, rank=0, nranks=1):
    _lib.check(_lib.load().flutas_b200_init(device, rank, nranks))


def device_field(arr):
    """Copy a Fortran-ordered numpy array into a torch CUDA tensor with identical storage order."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr.T)).cuda()      # C-order of the transpose == F-order storage
    return t


def host_field(t, shape):
    a = np.asfortranarray(t.cpu().numpy().T)
    assert a.shape == tuple(shape), (a.shape, shape)
    return a


class Plans:
    """arrplan(2,2): four opaque handles, Fortran order fwd-x, bwd-x, fwd-y, bwd-y."""

    def __init__(self):
        self.h = (C.c_void_p * 4)()
        self.normfft = None

    def __del__(self):
        try:
            if self.h[0]:
                fftend(self)
        except Exception:
            pass


def fftini(n_x, n_y, bcxy, c_or_f=("c", "c")):
    """bcxy = ("PP","NN") style pair of BC strings for x and y. Returns (arrplan, normfft)."""
    L = _lib.load()
    pl = Plans()
    nf = C.c_double()
    nx = (C.c_int * 3)(*n_x)
    ny = (C.c_int * 3)(*n_y)
    _lib.check(L.flutas_b200_fftini(nx, ny, (bcxy[0] + bcxy[1]).encode(), "".join(c_or_f).encode(), pl.h, C.byref(nf)))
    pl.normfft = nf.value
    return pl, nf.value


def fftend(arrplan):
    _lib.check(_lib.load().flutas_b200_fftend(arrplan.h))


def fft(plan, arr, n=None):
    """fft(plan,arr), src/fft.f90:181-193: unnormalised in-place r2r transform of a dense pencil array in FFTW's element
    order.  plan = `arrplan.h[q]` (q = 0 fwd-x, 1 bwd-x, 2 fwd-y, 3 bwd-y) or an `R2RPlan`; arr = Fortran-ordered numpy
    array, or a device tensor together with its extents n."""
    _use_torch_stream()
    if isinstance(plan, R2RPlan):
        plan = plan.h
    if n is None:
        n = arr.shape
    nn = (C.c_int * 3)(*n)
    _lib.check(_lib.load().flutas_b200_fft(C.c_void_p(plan) if isinstance(plan, int) else plan, nn, _ptr(arr)))
    return arr


FFTW_KINDS = {"R2HC": 0, "HC2R": 1, "REDFT01": 4, "REDFT10": 5, "REDFT11": 6, "RODFT01": 8, "RODFT10": 9, "RODFT11": 10}


class R2RPlan:
    """Stand-alone plan with the arguments of the reference's fftw_plan_guru_r2r calls (src/fft.f90:75-86,113-124)."""

    def __init__(self, n, stride, howmany_n, howmany_stride, kind):
        self.h = C.c_void_p()
        hn = (C.c_int * 2)(*howmany_n)
        hs = (C.c_int * 2)(*howmany_stride)
        code = FFTW_KINDS[kind] if isinstance(kind, str) else int(kind)
        _lib.check(_lib.load().flutas_b200_plan_r2r(n, stride, hn, hs, code, C.byref(self.h)))

    @property
    def dims(self):
        nn = (C.c_int * 3)()
        _lib.check(_lib.load().flutas_b200_plan_dims(self.h, nn))
        return tuple(nn)

    def destroy(self):
        if self.h:
            _lib.check(_lib.load().flutas_b200_destroy_plan(self.h))
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def solver(n, arrplan, normfft, lambdaxy, a, b, c, bcz, c_or_f, p):
    _use_torch_stream()
    nn = (C.c_int * 3)(*n)
    _lib.check(_lib.load().flutas_b200_solver(nn, arrplan.h, normfft, _ptr(lambdaxy), _ptr(a), _ptr(b), _ptr(c),
                                              bcz.encode(), "".join(c_or_f).encode(), _ptr(p)))
    return p


def fillps(nx, ny, nz, nh_d, nh_u, dxi, dyi, dzi, dzfi, dti, rho0, u, v, w, p):
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_fillps(nx, ny, nz, nh_d, nh_u, dxi, dyi, dzi, _ptr(dzfi), dti, rho0,
                                              _ptr(u), _ptr(v), _ptr(w), _ptr(p)))
    return p


def updt_rhs_b(nx, ny, nz, cbc, rhsbx, rhsby, rhsbz, p):
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_updt_rhs_b(nx, ny, nz, "".join(cbc).encode(), _ptr(rhsbx), _ptr(rhsby),
                                                  _ptr(rhsbz), _ptr(p)))
    return p


def correc(nx, ny, nz, nh_d, nh_u, dxi, dyi, dzi, dzci, dt, rho0, p, u, v, w, rho=None):
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_correc(nx, ny, nz, nh_d, nh_u, dxi, dyi, dzi, _ptr(dzci), dt, rho0,
                                              _ptr(p), _ptr(u), _ptr(v), _ptr(w), None))


def pres_sp_src(nx, ny, nz, f_t12, dxi, dyi, dzi, nh_d, nh_u, dzci, rho0i, pold, u, v, w):
    """pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,pold,u,v,w), src/source.f90:311"""
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_pres_sp_src(nx, ny, nz, f_t12, dxi, dyi, dzi, nh_d, nh_u, _ptr(dzci), rho0i,
                                                   _ptr(pold), _ptr(u), _ptr(v), _ptr(w)))


def pres_tw_src(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, dzci, rho0i, f_t12, f_t12_o, p, pold, rho, u, v, w):
    """pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,f_t12,f_t12_o,p,pold,rho,u,v,w), src/source.f90:247
    (constant-coefficient Poisson branch)"""
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_pres_tw_src(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, _ptr(dzci), rho0i, f_t12,
                                                   f_t12_o, _ptr(p), _ptr(pold), _ptr(rho), _ptr(u), _ptr(v), _ptr(w)))


def pold_update(nx, ny, nz, mode, p, pold):
    """mode 0: pold = p (main__single_phase.f90:693-699); mode 1: p = pold + p (:734-740); interior only"""
    _use_torch_stream()
    _lib.check(_lib.load().flutas_b200_pold_update(nx, ny, nz, mode, _ptr(p), _ptr(pold)))


def load(io, filename, n, fld, ng=None, start=(0, 0, 0), nh=0):
    """load(io,filename,n,fld), src/load.f90:21: read ('r') / write ('w') this rank's block of a raw FP64 restart file in
    global column-major order.  ng defaults to n (single rank); start = 0-based global offset of the block; nh = halo
    width of fld (numpy F-ordered array or CUDA tensor)."""
    ng = tuple(n) if ng is None else tuple(ng)
    _lib.check(_lib.load().flutas_b200_load(io.encode(), str(filename).encode(), (C.c_int * 3)(*ng), (C.c_int * 3)(*n),
                                            (C.c_int * 3)(*start), nh, _ptr(fld)))
    return fld


def boundp(cbc, n, bc, nh_d, nh_p, dl, dzc, dzf, p):
    """boundp(cbc,n,bc,nh_d,nh_p,halo,dl,dzc,dzf,p), src/bound.f90:146 (the MPI `halo` datatypes have no counterpart).
    cbc: three 2-character strings; bc: (3,2) boundary values."""
    _use_torch_stream()
    nn = (C.c_int * 3)(*n)
    bcv = (C.c_double * 6)(*[float(bc[d][s]) for d in range(3) for s in range(2)])
    dlv = (C.c_double * 3)(*[float(x) for x in dl])
    _lib.check(_lib.load().flutas_b200_boundp("".join(cbc).encode(), nn, bcv, nh_d, nh_p, dlv, _ptr(dzc), _ptr(dzf),
                                              _ptr(p)))
    return p


def bounduvw(cbc, n, bc, nh_d, nh_u, isoutflow, dl, dzc, dzf, u, v, w):
    """bounduvw(cbc,n,bc,nh_d,nh_u,halo,isoutflow,dl,dzc,dzf,u,v,w), src/bound.f90:17 (no `halo` MPI datatypes).
    cbc[ibound][idir][field] characters, bc[ibound][idir][field] values, isoutflow[ibound][idir] booleans."""
    _use_torch_stream()
    cc = "".join(cbc[ib][d][f] for f in range(3) for d in range(3) for ib in range(2))       # Fortran order (0:1,3,3)
    bv = (C.c_double * 18)(*[float(bc[ib][d][f]) for f in range(3) for d in range(3) for ib in range(2)])
    io = (C.c_int * 6)(*[1 if isoutflow[ib][d] else 0 for d in range(3) for ib in range(2)])
    nn = (C.c_int * 3)(*n)
    dlv = (C.c_double * 3)(*[float(x) for x in dl])
    _lib.check(_lib.load().flutas_b200_bounduvw(cc.encode(), nn, bv, nh_d, nh_u, io, dlv, _ptr(dzc), _ptr(dzf),
                                                _ptr(u), _ptr(v), _ptr(w)))


def chkdt(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, dzci, dzfi, u, v, w):
    """the field reduction of chkdt_sp / chkdt_tw (src/chkdt.f90:150-173): this rank's max(dtix, dtiy, dtiz)"""
    _use_torch_stream()
    dti = C.c_double()
    _lib.check(_lib.load().flutas_b200_chkdt(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, _ptr(dzci), _ptr(dzfi), _ptr(u), _ptr(v),
                                             _ptr(w), C.byref(dti)))
    return dti.value


def chkdiv(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, dzfi, u, v, w):
    _use_torch_stream()
    tot, mx = C.c_double(), C.c_double()
    _lib.check(_lib.load().flutas_b200_chkdiv(nx, ny, nz, dxi, dyi, dzi, nh_d, nh_u, _ptr(dzfi), _ptr(u), _ptr(v),
                                              _ptr(w), C.byref(tot), C.byref(mx)))
    return tot.value, mx.value


def launch_count():
    return _lib.load().flutas_b200_launch_count()


def profile_enable(on=True):
    _lib.load().flutas_b200_profile_enable(1 if on else 0)


def profile_read():
    """{stage name: (ms_sum, launches)} since the previous read (synchronises the stream)."""
    L = _lib.load()
    n = L.flutas_b200_profile_stage_count()
    ms = (C.c_double * n)()
    cnt = (C.c_long * n)()
    _lib.check(L.flutas_b200_profile_read(ms, cnt))
    return {L.flutas_b200_profile_stage_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}
