"""Builds flutas_b200/csrc/libflutas_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SO = os.path.join(CSRC, "libflutas_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["kernels.cuh", "tile_fft.cuh", "line_plan.h", "thomas_tile.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libflutas_b200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(CSRC, "..", "..", "include", "flutas_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not is_stale():
        return SO
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    env = dict(os.environ)
    for cc in ("/usr/bin/g++",):
        if os.path.exists(cc):
            cmd[1:1] = ["-ccbin", cc]
            break
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
