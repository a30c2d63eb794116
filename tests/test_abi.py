"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/flutas_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from flutas_b200 import api, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "flutas_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(flutas_b200_\w+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported():
    L = lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert set(lib.EXPORTS) == set(names)


def test_version_string():
    assert b"sm_100a" in lib.load().flutas_b200_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.FlutasB200Error, match="no CUDA device"):
        api.fftini((8, 8, 8), (8, 8, 8), ("PP", "PP"))
    with pytest.raises(lib.FlutasB200Error, match="no CUDA device"):
        api.init(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "flutas_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("with the oracle", ""), f


def test_header_is_plain_c_and_the_shim_binds_every_compute_entry_point():
    """The boundary is a C ABI: the header must compile as C99 (no C++/torch types), and the Fortran shim
    (flutas_b200/fortran/flutas_b200_shim.f90) must bind every entry point the reference's call sites need."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc:
        subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                               os.path.join(ROOT, "include", "flutas_b200.h")])
    shim = open(os.path.join(ROOT, "flutas_b200", "fortran", "flutas_b200_shim.f90")).read()
    bound = set(re.findall(r"bind\(C,\s*name='(flutas_b200_\w+)'\)", shim))
    needed = {"flutas_b200_fftini", "flutas_b200_fftend", "flutas_b200_solver", "flutas_b200_fillps", "flutas_b200_correc",
              "flutas_b200_chkdiv", "flutas_b200_boundp", "flutas_b200_pres_sp_src", "flutas_b200_pres_tw_src",
              "flutas_b200_pold_update", "flutas_b200_load", "flutas_b200_updt_rhs_b", "flutas_b200_solver_slab",
              "flutas_b200_p2p_export", "flutas_b200_p2p_attach", "flutas_b200_set_halo_exchange", "flutas_b200_bounduvw",
              "flutas_b200_chkdt", "flutas_b200_init", "flutas_b200_alloc_managed"}
    for mod in ("mod_fft", "mod_solver_cpu", "mod_solver_gpu", "mod_fillps", "mod_correc", "mod_chkdiv"):
        assert re.search(r"^module %s\b" % mod, shim, re.M), mod          # the main's `use` lines resolve unchanged
    for proc in ("solver_cpu(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p)",
                 "updt_rhs_b(nx,ny,nz,c_or_f,cbc,nh_p,rhsbx,rhsby,rhsbz,p)",
                 "bounduvw(cbc,n,bc,nh_d,nh_u,halo,isoutflow,dl,dzc,dzf,u,v,w)",
                 "chkdt_sp(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,dzfi,u,v,w,dtmax)"):
        assert "subroutine " + proc in shim, proc                            # the reference's argument lists
    assert needed <= bound, needed - bound
    assert bound <= set(_declared_symbols()), bound - set(_declared_symbols())


def test_fftw_seam_library_exports_the_references_fftw_symbols():
    """src/fftw.f90:15-36 binds fftw_plan_guru_r2r; src/fft.f90:53-57,165-175,188-190 call the legacy dfftw_* entry points
    (compiler-mangled with a trailing underscore).  The seam library must load without a GPU and carry all of them."""
    import ctypes as C
    from flutas_b200 import build as b
    lib.load()
    assert os.path.exists(b.SEAM_SO)
    S = C.CDLL(b.SEAM_SO)
    for name in ("fftw_plan_guru_r2r", "dfftw_execute_r2r_", "dfftw_destroy_plan_", "dfftw_init_threads_",
                 "dfftw_plan_with_nthreads_", "dfftw_cleanup_threads_", "dfftw_execute_r2r", "dfftw_destroy_plan"):
        assert hasattr(S, name), name
    ierr = C.c_int(0)
    S.dfftw_init_threads_(C.byref(ierr))
    assert ierr.value != 0                       # FFTW: non-zero = success
    # the main library must NOT carry FFTW's names (it would shadow a real FFTW in the host program)
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", b.SO], capture_output=True, text=True).stdout
    assert "fftw" not in syms
