"""flutas_b200: B200-native (sm_100a) FP64 pressure-Poisson path for FluTAS (fillps -> solver -> correc).

The product is csrc/libflutas_b200.so (C ABI in include/flutas_b200.h); this package is the Python
host side: `api` mirrors the reference's Fortran interface, `initsolver` mirrors the host-only set-up,
`cases` builds the synthetic BASELINE configurations.
"""
from . import api, cases, initsolver  # noqa: F401

__all__ = ["api", "cases", "initsolver"]
