"""Optional kernel variants (selected by environment variables that the library reads once per process) must stay
correct: each one re-runs a subset of the parity tests in a subprocess."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SUBSET = "devptr and not genericfft and not p2fft and (uni or C1 or chan or deep or p2c or rb or wide)"
VARIANTS = {
    "z-8col-tiles": {"FLUTAS_B200_THOMAS_UNI": "0", "FLUTAS_B200_THOMAS_CFG": "0"},
    "z-cluster-pair": {"FLUTAS_B200_THOMAS_UNI": "0", "FLUTAS_B200_THOMAS_CFG": "2"},
    "z-general-tma": {"FLUTAS_B200_THOMAS_UNI": "0", "FLUTAS_B200_THOMAS_TMA_GEN": "1"},
    "z-uniform-cpasync": {"FLUTAS_B200_THOMAS_TMA": "0"},
    "z-double-buffer": {"FLUTAS_B200_THOMAS_UNI": "0", "FLUTAS_B200_THOMAS_NBUF": "2"},
    "y-8-values": {"FLUTAS_B200_Y8": "1"},
    "y-wide": {"FLUTAS_B200_YWIDE": "1"},
    "y-narrow": {"FLUTAS_B200_YWIDE": "0"},
    "y-8-values-wide": {"FLUTAS_B200_Y8": "1", "FLUTAS_B200_Y8WIDE": "1", "_file": "test_gpu_fft.py", "_subset": "arrplan"},
    "x-8-values": {"FLUTAS_B200_X8": "1", "_file": "test_gpu_fft.py", "_subset": "arrplan"},
    "x-16-values": {"FLUTAS_B200_X8": "0", "_file": "test_gpu_fft.py", "_subset": "arrplan"},
    "correc-scalar": {"FLUTAS_B200_CORREC_VEC": "0", "_subset": "stencils"},
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_passes_parity_subset(name):
    env = dict(os.environ)
    var = dict(VARIANTS[name])
    subset = var.pop("_subset", SUBSET)
    target = var.pop("_file", "test_gpu_parity.py")
    env.update(var)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, target), "-m", "gpu", "-x", "-q",
                          "-k", subset], env=env, capture_output=True, text=True, timeout=900)
    tail = out.stdout[-1500:] + out.stderr[-500:]
    assert out.returncode == 0 and " passed" in out.stdout, tail
