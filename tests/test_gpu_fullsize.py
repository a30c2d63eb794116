"""Parity at the BASELINE.json sizes: the CUDA pressure step (through the C ABI, device resident) against the CPU oracle
(`oracle.Solver.solve`, restatement of src/solver_cpu.f90:20-223) on the SAME bytes.

Bars (BASELINE.json north_star, written here):
  * right-hand side after fillps + updt_rhs_b: bit for bit;
  * pressure: max|p - p_oracle| / max|p_oracle| <= 1e-12 on p - mean(p) (every config is all-periodic/Neumann: the
    additive constant is round-off defined in the reference; the raw figure is printed);
  * chkdiv's divmax after correc: <= 1e-12 of max|div u*| and no larger than twice what the oracle's own
    projection leaves on the same bytes.  (The absolute value has a grid-set floor, eps |p| dt / dz^2 -- 4e-12 at 512^3 --
    that the reference arithmetic itself sits on; it is printed.)
The oracle takes 1-15 s per case on the GPU box's host cores (1024^3: 8.6 GB per field, needs ~45 GB of host memory).
"""
import os

import numpy as np
import pytest

from flutas_b200 import api
from flutas_b200.cases import Case
from oracle import oracle, parity

pytestmark = pytest.mark.gpu
TOL = 1e-12

FULL = [
    ("C2", {}),                      # 512^3 PP/PP/PP
    ("C3", {}),                      # 1024x512x512 PP/PP/NN, uniform z (shared-LU z kernel + reference-order columns)
    ("C3", {"gr": 2.0}),             # ... tanh-stretched z (general z kernel)
    ("C4", {}),                      # 512^3, rho0 = 0.1
    ("C5w1", {}),                    # 1024x1024x512 NN/NN/NN
    ("NS", {}),                      # 1024^3 PP/PP/NN: the north-star grid
]


def _host_mem_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 1e9


@pytest.mark.parametrize("cid,kw", FULL, ids=["%s%s" % (c, "-gr%g" % k["gr"] if k else "") for c, k in FULL])
def test_pressure_step_matches_oracle_at_baseline_size(cid, kw):
    import torch
    case = Case.from_config(cid, **kw)
    n1, n2, n3 = case.ng
    need_gb = 5.2 * 8e-9 * (n1 + 2) * (n2 + 2) * (n3 + 2)            # u, v, w, p, oracle work array + slack
    if _host_mem_gb() < need_gb:
        pytest.skip("host has less than %.0f GB available for the oracle at this size" % need_gb)
    r = parity.pressure_step_parity(case, api, oracle, threads=os.cpu_count())
    print("%s %s %s gr=%s: max|dp|/max|p| = %.2e (gauge-fixed; raw %.2e), divmax after correc = %.2e (oracle's own %.2e; "
          "before %.2e -> relative %.2e), rhs bit-exact %s, oracle %.1f s on %d threads"
          % (cid, case.ng, "/".join(case.cbc), kw.get("gr", 0.0), r["err"], r["raw"], r["divmax"], r["divmax_oracle"],
             r["divmax_before"], r["divmax_rel"], r["rhs_bit_exact"], r["oracle_solve_s"], r["oracle_threads"]))
    assert r["rhs_bit_exact"]
    assert r["err"] <= TOL, r
    assert r["divmax_rel"] <= TOL, r
    assert r["divmax"] <= 2.0 * r["divmax_oracle"] + 1e-13, r
    torch.cuda.empty_cache()
