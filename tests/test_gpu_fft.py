"""fft(plan,arr) (src/fft.f90:181-193) kind by kind against the oracle's FFTW-definition r2r: the four arrplan handles,
stand-alone guru plans (src/fft.f90:75-86,113-124) and the FFTW-named seam library driven the way fft.f90 drives FFTW."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

TOL = 2e-13          # max|Y_gpu - Y_ref| / max|Y_ref| of one unnormalised transform (n <= 2048: ~ eps * log2 n * few)


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from flutas_b200 import api as a
    a.init(0)
    return a


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


# register kernels: powers of two 32..2048; shared-memory radix-2/3/5 kernels: the rest
@pytest.mark.parametrize("bcx,bcy", [("PP", "NN"), ("NN", "PP"), ("DD", "ND"), ("ND", "DN"), ("DN", "DD")])
@pytest.mark.parametrize("n1,n2,n3", [(64, 32, 5), (12, 72, 3), (256, 1024, 2), (2048, 20, 2), (30, 512, 3), (1024, 16, 3)])
def test_arrplan_handles_match_fftw_definitions(api, bcx, bcy, n1, n2, n3):
    rng = np.random.default_rng(n1 * 7 + n2)
    pl, _ = api.fftini((n1, n2, n3), (n1, n2, n3), (bcx, bcy))
    for axis, bc, q in ((0, bcx, 0), (1, bcy, 2)):
        kf, kb, _norm = oracle.find_fft(bc)
        x = np.asfortranarray(rng.uniform(-1, 1, (n1, n2, n3)))
        ref = oracle.r2r(kf, x.copy(order="F"), axis)
        got = api.fft(pl.h[q], x.copy(order="F"))                       # forward, host array
        assert _rel(got, ref) <= TOL, (bc, axis, "fwd")
        refb = oracle.r2r(kb, ref.copy(order="F"), axis)
        gotb = api.fft(pl.h[q + 1], ref.copy(order="F"))                # backward of the reference spectrum
        assert _rel(gotb, refb) <= TOL, (bc, axis, "bwd")
        # device-resident array: same bytes in, same result as the host call
        xd = api.device_field(x)
        api.fft(pl.h[q], xd, (n1, n2, n3))
        assert np.array_equal(api.host_field(xd, (n1, n2, n3)), got)
    api.fftend(pl)


def test_fft_rejects_a_mismatched_array(api):
    from flutas_b200.lib import FlutasB200Error
    pl, _ = api.fftini((32, 16, 2), (32, 16, 2), ("PP", "PP"))
    with pytest.raises(FlutasB200Error, match="does not match the plan"):
        api.fft(pl.h[0], np.zeros((16, 16, 2), order="F"))
    with pytest.raises(FlutasB200Error, match="does not match the plan"):
        api.fft(pl.h[2], np.zeros((32, 32, 2), order="F"))
    api.fftend(pl)


@pytest.mark.parametrize("kind", ["R2HC", "HC2R", "REDFT10", "REDFT01", "RODFT10", "RODFT01", "REDFT11", "RODFT11"])
@pytest.mark.parametrize("layout", ["x", "y"])
def test_guru_plans(api, kind, layout):
    n1, n2, n3 = 128, 48, 3
    rng = np.random.default_rng(3)
    x = np.asfortranarray(rng.uniform(-1, 1, (n1, n2, n3)))
    if layout == "x":
        pl = api.R2RPlan(n1, 1, (n2, n3), (n1, n1 * n2), kind)          # src/fft.f90:75-86
    else:
        pl = api.R2RPlan(n2, n1, (n1, n3), (1, n1 * n2), kind)          # src/fft.f90:113-124
    assert pl.dims == (n1, n2, n3)
    ref = oracle.r2r(kind, x.copy(order="F"), 0 if layout == "x" else 1)
    got = api.fft(pl, x.copy(order="F"))
    assert _rel(got, ref) <= TOL
    pl.destroy()


def test_guru_plan_rejections(api):
    from flutas_b200.lib import FlutasB200Error
    with pytest.raises(FlutasB200Error, match="r2r kind"):
        api.R2RPlan(32, 1, (4, 2), (32, 128), 3)                        # REDFT00: face-centred, dead code in FluTAS
    with pytest.raises(FlutasB200Error, match="pencil layouts"):
        api.R2RPlan(32, 2, (4, 2), (64, 256), "R2HC")                   # strided x lines
    with pytest.raises(FlutasB200Error, match="transform length"):
        api.R2RPlan(14, 1, (4, 2), (14, 56), "R2HC")                    # 7 is not a supported radix


def test_fftw_seam_library_runs_the_references_call_sequence(api):
    """plan (bind(C) guru call) -> dfftw_execute_r2r_(plan, arr, arr) -> dfftw_destroy_plan_(plan), all arguments of
    the legacy entry points by reference, as src/fft.f90:85-86,123-124,165-175,188-190 issue them."""
    from flutas_b200 import build as b
    if os.environ.get("FLUTAS_B200_LIB"):
        pytest.skip("the seam library is linked against the default build, not the FLUTAS_B200_LIB override")
    S = C.CDLL(b.SEAM_SO)

    class IoDim(C.Structure):
        _fields_ = [("n", C.c_int), ("is_", C.c_int), ("os", C.c_int)]

    S.fftw_plan_guru_r2r.restype = C.c_void_p
    S.fftw_plan_guru_r2r.argtypes = [C.c_int, C.POINTER(IoDim), C.c_int, C.POINTER(IoDim), C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_int), C.c_int]
    n1, n2, n3 = 64, 40, 4
    rng = np.random.default_rng(11)
    arr = np.asfortranarray(rng.uniform(-1, 1, (n1, n2, n3)))
    ierr = C.c_int(0)
    S.dfftw_init_threads_(C.byref(ierr))
    assert ierr.value != 0
    S.dfftw_plan_with_nthreads_(C.byref(C.c_int(4)))
    for bcx, bcy in (("PP", "NN"), ("DD", "DN")):
        plans = {}
        for axis, bc in ((0, bcx), (1, bcy)):
            kf, kb, _ = oracle.find_fft(bc)
            if axis == 0:
                dim = (IoDim * 1)(IoDim(n1, 1, 1))
                hm = (IoDim * 2)(IoDim(n2, n1, n1), IoDim(n3, n1 * n2, n1 * n2))
            else:
                dim = (IoDim * 1)(IoDim(n2, n1, n1))
                hm = (IoDim * 2)(IoDim(n1, 1, 1), IoDim(n3, n1 * n2, n1 * n2))
            for name, k in (("fwd", kf), (("bwd"), kb)):
                code = oracle.KINDS[k] if isinstance(k, str) else int(k)
                h = S.fftw_plan_guru_r2r(1, dim, 2, hm, arr.ctypes.data, arr.ctypes.data, C.byref(C.c_int(code)), 64)
                assert h
                plans[(axis, name)] = (C.c_void_p(h), k)
        work = arr.copy(order="F")
        ref = arr.copy(order="F")
        for key in ((0, "fwd"), (1, "fwd"), (1, "bwd"), (0, "bwd")):          # the transform sandwich of solver_cpu.f90:59-89
            h, k = plans[key]
            S.dfftw_execute_r2r_(C.byref(h), work.ctypes.data, work.ctypes.data)
            oracle.r2r(k, ref, key[0])
            assert _rel(work, ref) <= 4 * TOL, (bcx, bcy, key)
        nf = oracle.normfft(n1, n2, bcx, bcy)
        assert _rel(work * nf, arr) <= 1e-13                                   # round trip = 1 / normfft (src/fft.f90:150)
        # out-of-place execute: `in` is left alone
        h, k = plans[(0, "fwd")]
        out = np.zeros_like(arr)
        keep = arr.copy(order="F")
        S.dfftw_execute_r2r_(C.byref(h), arr.ctypes.data, out.ctypes.data)
        assert np.array_equal(arr, keep)
        assert _rel(out, oracle.r2r(k, arr.copy(order="F"), 0)) <= TOL
        for h, _k in plans.values():
            S.dfftw_destroy_plan_(C.byref(h))
            assert not h.value
    S.dfftw_cleanup_threads_(C.byref(ierr))
