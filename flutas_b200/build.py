"""Builds flutas_b200/csrc/libflutas_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SO = os.path.join(CSRC, "libflutas_b200.so")
SEAM_SO = os.path.join(CSRC, "libflutas_b200_fftw.so")     # FFTW-named entry points over the C ABI (csrc/fftw_seam.cpp)
SOURCES = ["capi.cu", "fft_p2_x.cu", "fft_p2_y.cu", "fft_reg_x_fwd.cu", "fft_reg_x_bwd.cu", "fft_reg_y_fwd.cu",
           "fft_reg_y_bwd.cu"]
HEADERS = ["kernels.cuh", "tile_fft.cuh", "line_plan.h", "thomas_tile.cuh", "thomas_reg.cuh", "thomas_uni.cuh", "thomas_ref.cuh", "geom.cuh", "fft_p2.cuh", "fft_p2.h", "reg_fft.cuh", "fft_reg.cuh", "fft_reg.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libflutas_b200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(SO) or not os.path.exists(SEAM_SO):
        return True
    t = min(os.path.getmtime(SO), os.path.getmtime(SEAM_SO))
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS + ["fftw_seam.cpp"]]
    deps.append(os.path.join(CSRC, "..", "..", "include", "flutas_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), tag=None):
    """defines/tag: A/B builds of compile-time variants -> csrc/libflutas_b200_<tag>.so (select with FLUTAS_B200_LIB)"""
    so = SO if tag is None else SO.replace(".so", "_%s.so" % tag)
    if tag is None and not force and not is_stale():
        return SO
    nvcc = [_nvcc()]
    if os.path.exists("/usr/bin/g++"):
        nvcc += ["-ccbin", "/usr/bin/g++"]
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines]
    objdir = os.path.join(CSRC, "build" if tag is None else "build_" + tag)
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        subprocess.check_call(nvcc + flags + ["-c", "-o", obj, src], cwd=CSRC)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call(nvcc + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs, cwd=CSRC)
    if tag is None:
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SEAM_SO, "fftw_seam.cpp", "-L.", "-lflutas_b200",
                               "-Wl,-rpath,$ORIGIN"], cwd=CSRC)
    return so


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1:                                   # python -m flutas_b200.build <tag> DEF=1 DEF2=0 ...
        print(build(force=True, defines=sys.argv[2:], tag=sys.argv[1]))
    else:
        print(build(force=True, verbose=True))
