"""Shared helpers for the tests (CPU side)."""
import glob
import os

import numpy as np

from flutas_b200.cases import Case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(path):
    z = np.load(path)
    case = Case(tuple(int(x) for x in z["ng"]), tuple(str(x) for x in z["cbc"]), tuple(float(x) for x in z["lengths"]),
                rho0=float(z["rho0"]), gr=float(z["gr"]), seed=int(z["seed"]),
                name=os.path.basename(path)[:-4])
    return case, np.asfortranarray(z["rhs"]), np.asfortranarray(z["p"])


def gauge_rel_err(a, b, singular):
    """max|a-b|/max|b| on interior arrays, mean removed when the operator is singular."""
    if singular:
        a = a - a.mean()
        b = b - b.mean()
    return np.max(np.abs(a - b)) / np.max(np.abs(b))
