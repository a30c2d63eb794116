"""ctypes binding of libflutas_b200.so (include/flutas_b200.h).  Fails loudly when the library is
missing or there is no CUDA device -- there is no CPU fallback on this path."""
import ctypes as C
import os

from . import build as _build

_dp = C.POINTER(C.c_double)
_LIB = None


class FlutasB200Error(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.environ.get("FLUTAS_B200_LIB") or _build.SO      # override: A/B runs of kernel variants
    if not os.path.exists(so):
        so = _build.build()
    L = C.CDLL(so)
    vp, ci, cd, cc = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    ip = C.POINTER(C.c_int)
    L.flutas_b200_version.restype = cc
    L.flutas_b200_last_error.restype = cc
    L.flutas_b200_launch_count.restype = C.c_long
    L.flutas_b200_init.argtypes = [ci, ci, ci]
    L.flutas_b200_set_stream.argtypes = [vp]
    L.flutas_b200_alloc.restype = vp
    L.flutas_b200_alloc.argtypes = [C.c_size_t]
    L.flutas_b200_alloc_managed.restype = vp
    L.flutas_b200_alloc_managed.argtypes = [C.c_size_t]
    L.flutas_b200_free.argtypes = [vp]
    L.flutas_b200_memcpy.argtypes = [vp, vp, C.c_size_t]
    L.flutas_b200_fftini.argtypes = [ip, ip, cc, cc, C.POINTER(vp), _dp]
    L.flutas_b200_fftend.argtypes = [C.POINTER(vp)]
    L.flutas_b200_solver.argtypes = [ip, C.POINTER(vp), cd, vp, vp, vp, vp, cc, cc, vp]
    L.flutas_b200_fft.argtypes = [vp, ip, vp]
    L.flutas_b200_plan_r2r.argtypes = [ci, ci, ip, ip, ci, C.POINTER(vp)]
    L.flutas_b200_plan_dims.argtypes = [vp, ip]
    L.flutas_b200_destroy_plan.argtypes = [vp]
    L.flutas_b200_solver_invalidate.argtypes = [C.POINTER(vp)]
    L.flutas_b200_debug_thomas_mode.argtypes = [C.POINTER(vp), ci]
    L.flutas_b200_debug_generic_fft.argtypes = [ci]
    L.flutas_b200_debug_ref_tol.argtypes = [cd]
    L.flutas_b200_fillps.argtypes = [ci] * 5 + [cd] * 3 + [vp, cd, cd, vp, vp, vp, vp]
    L.flutas_b200_updt_rhs_b.argtypes = [ci] * 3 + [cc, vp, vp, vp, vp]
    L.flutas_b200_correc.argtypes = [ci] * 5 + [cd] * 3 + [vp, cd, cd, vp, vp, vp, vp, vp]
    L.flutas_b200_chkdiv.argtypes = [ci] * 3 + [cd] * 3 + [ci] * 2 + [vp] * 4 + [_dp, _dp]
    L.flutas_b200_boundp.argtypes = [cc, ip, _dp, ci, ci, _dp, vp, vp, vp]
    L.flutas_b200_pres_sp_src.argtypes = [ci] * 3 + [cd] * 4 + [ci] * 2 + [vp, cd, vp, vp, vp, vp]
    L.flutas_b200_pres_tw_src.argtypes = [ci] * 3 + [cd] * 3 + [ci] * 2 + [vp] + [cd] * 3 + [vp] * 6
    L.flutas_b200_pold_update.argtypes = [ci] * 4 + [vp, vp]
    L.flutas_b200_load.argtypes = [C.c_char, cc, ip, ip, ip, ci, vp]
    L.flutas_b200_bounduvw.argtypes = [cc, ip, _dp, ci, ci, ip, _dp, vp, vp, vp, vp, vp]
    L.flutas_b200_chkdt.argtypes = [ci] * 3 + [cd] * 3 + [ci] * 2 + [vp] * 5 + [_dp]
    L.flutas_b200_set_halo_exchange.argtypes = [vp, vp]
    L.flutas_b200_set_alltoall.argtypes = [vp, vp]
    L.flutas_b200_p2p_handle_bytes.restype = C.c_size_t
    L.flutas_b200_p2p_export.argtypes = [C.POINTER(vp), ip, vp]
    L.flutas_b200_p2p_attach.argtypes = [C.POINTER(vp), vp]
    L.flutas_b200_p2p_errors.argtypes = [C.POINTER(vp)]
    L.flutas_b200_slab_config.argtypes = [ci, ci, ci]
    L.flutas_b200_slab_distributed_z.argtypes = [ci]
    L.flutas_b200_slab_last_distributed.argtypes = [C.POINTER(vp)]
    L.flutas_b200_solver_slab.argtypes = [ip, C.POINTER(vp), cd, vp, vp, vp, vp, cc, cc, vp]
    L.flutas_b200_profile_enable.argtypes = [ci]
    L.flutas_b200_profile_stage_name.restype = cc
    L.flutas_b200_profile_stage_name.argtypes = [ci]
    L.flutas_b200_profile_read.argtypes = [_dp, C.POINTER(C.c_long)]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise FlutasB200Error(load().flutas_b200_last_error().decode())


EXPORTS = [
    "flutas_b200_version", "flutas_b200_last_error", "flutas_b200_init", "flutas_b200_set_stream",
    "flutas_b200_alloc", "flutas_b200_alloc_managed", "flutas_b200_free", "flutas_b200_memcpy", "flutas_b200_synchronize",
    "flutas_b200_fftini", "flutas_b200_fftend", "flutas_b200_solver", "flutas_b200_solver_invalidate",
    "flutas_b200_fft", "flutas_b200_plan_r2r", "flutas_b200_plan_dims", "flutas_b200_destroy_plan",
    "flutas_b200_fillps", "flutas_b200_updt_rhs_b", "flutas_b200_correc", "flutas_b200_chkdiv",
    "flutas_b200_launch_count", "flutas_b200_profile_enable", "flutas_b200_profile_stage_count",
    "flutas_b200_profile_stage_name", "flutas_b200_profile_read", "flutas_b200_set_alltoall",
    "flutas_b200_p2p_handle_bytes", "flutas_b200_p2p_export", "flutas_b200_p2p_attach", "flutas_b200_p2p_errors",
    "flutas_b200_solver_slab", "flutas_b200_slab_config", "flutas_b200_slab_distributed_z", "flutas_b200_slab_last_distributed", "flutas_b200_boundp", "flutas_b200_set_halo_exchange",
    "flutas_b200_bounduvw", "flutas_b200_chkdt", "flutas_b200_pres_sp_src", "flutas_b200_pres_tw_src", "flutas_b200_pold_update", "flutas_b200_load",
]
