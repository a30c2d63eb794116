"""Full-size parity of the CUDA pressure step against the CPU oracle, on the same bytes.

TEST INFRASTRUCTURE (lives under oracle/): used by tests/test_gpu_fullsize.py and by bench.py's `parity` key.  The
product (flutas_b200/) never imports it.  What it does for one configuration:

    u,v,w  <- Case.velocity()                       seeded, BC-consistent, host
    oracle : fillps -> updt_rhs_b -> solver_cpu     (oracle.Solver.solve, restatement of solver_cpu.f90:20-223)
    CUDA   : fillps -> updt_rhs_b -> solver -> boundp -> correc -> (halo refresh) -> chkdiv     through the C ABI
    compare: right-hand side bit for bit; max|dp|/max|p| after removing the mean of each field when the operator is
             singular (all BASELINE configs), the raw figure too; chkdiv's divmax after the correction.

The comparison runs on the device in k-blocks (the fields are up to 8.7 GB each).
"""
import time

import numpy as np


def refresh_velocity_halos_device(case, ud, vd, wd):
    """torch version of Case.refresh_velocity_halos (stand-in for bounduvw, src/bound.f90:17-144) for tensors stored
    [k][j][i]: periodic wrap of the halos, zero wall-normal face velocity where the pressure BC is Neumann."""
    h = case.nh_u
    for d, n in enumerate(case.ng):
        bc = case.cbc[d]
        td = 2 - d                                       # tensor dimension of Fortran dimension d
        fld = (ud, vd, wd)[d]
        if bc == "PP":
            for f in (ud, vd, wd):
                f.narrow(td, 0, h).copy_(f.narrow(td, n, h))
                f.narrow(td, n + h, h).copy_(f.narrow(td, h, h))
        for ib in (0, 1):
            if bc[ib] == "N":
                fld.select(td, (h - 1) if ib == 0 else (n + h - 1)).zero_()
            elif bc[ib] == "D":
                raise ValueError("Dirichlet pressure walls need correct_dirichlet_faces (host path of the small tests)")


def _interior(t):
    return t[1:-1, 1:-1, 1:-1]


def compare_on_device(pd, pref_d, singular, nblk=16):
    """max|dp|/max|p| between two device fields with halo 1 ([k][j][i] storage), gauge-fixed (mean removed) and raw."""
    import torch
    a, b = _interior(pd), _interior(pref_d)
    n3 = a.shape[0]
    npts = float(a.numel())
    sa = sb = 0.0
    for k0 in range(0, n3, max(1, n3 // nblk)):
        k1 = min(n3, k0 + max(1, n3 // nblk))
        sa += float(a[k0:k1].sum(dtype=torch.float64))
        sb += float(b[k0:k1].sum(dtype=torch.float64))
    ma, mb = sa / npts, sb / npts
    raw = gauge = ref_raw = ref_gauge = 0.0
    for k0 in range(0, n3, max(1, n3 // nblk)):
        k1 = min(n3, k0 + max(1, n3 // nblk))
        d = a[k0:k1] - b[k0:k1]
        raw = max(raw, float(d.abs().max()))
        gauge = max(gauge, float((d - (ma - mb)).abs().max()))
        ref_raw = max(ref_raw, float(b[k0:k1].abs().max()))
        ref_gauge = max(ref_gauge, float((b[k0:k1] - mb).abs().max()))
        del d
    if singular:
        return gauge / ref_gauge, raw / ref_raw
    return raw / ref_raw, raw / ref_raw


def pressure_step_parity(case, api, oracle, threads=None, host_fields=None, dev_fields=None):
    """One pressure step of `case` on cuda (through `api`) and on the CPU oracle; returns a dict of parity figures.
    `threads`: oracle OpenMP threads (None = leave as is).  host_fields = (u, v, w) as returned by case.velocity() and
    dev_fields = their device copies may be passed in (bench.py reuses them); the device copies are not modified."""
    import torch
    if threads:
        oracle.set_num_threads(int(threads))
    s, n, cbc = case.setup, case.ng, case.cbc
    t0 = time.perf_counter()
    u, v, w = host_fields if host_fields is not None else case.velocity()
    t_gen = time.perf_counter() - t0
    if dev_fields is not None:
        ud, vd, wd = (f.clone() for f in dev_fields)      # correc below updates them in place
    else:
        ud, vd, wd = (api.device_field(f) for f in (u, v, w))
    pd = api.device_field(case.new_p())
    # --- right-hand side on both sides, bit for bit
    po = case.new_p()
    oracle.fillps(n, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, po)
    oracle.updt_rhs_b(n, s.rhsbx, s.rhsby, s.rhsbz, po)
    api.fillps(*n, case.nh_d, case.nh_u, *s.dli, s.dzfi, case.dti, case.rho0, ud, vd, wd, pd)
    api.updt_rhs_b(*n, cbc, s.rhsbx, s.rhsby, s.rhsbz, pd)
    _, div_before = api.chkdiv(*n, *s.dli, case.nh_d, case.nh_u, s.dzfi, ud, vd, wd)
    ref_d = api.device_field(po)
    rhs_equal = bool(torch.equal(pd, ref_d))
    del ref_d
    # --- solve on both sides
    t0 = time.perf_counter()
    oracle.Solver(n, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], po)
    t_cpu = time.perf_counter() - t0
    pl, nf = api.fftini(n, n, (cbc[0], cbc[1]))
    api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, cbc[2], "ccc", pd)
    torch.cuda.synchronize()
    ref_d = api.device_field(po)
    err, raw = compare_on_device(pd, ref_d, case.singular)
    del ref_d
    # --- the oracle's own projection: what divergence the reference arithmetic itself leaves (u, v, w are consumed)
    case.boundp(po)
    oracle.correc(n, case.nh_d, case.nh_u, s.dli, s.dzci, case.dt, case.rho0, po, u, v, w)
    case.refresh_velocity_halos(u, v, w)
    _, divmax_oracle = oracle.chkdiv(n, s.dli, case.nh_d, case.nh_u, s.dzfi, u, v, w)
    del u, v, w, po
    # --- projection with the CUDA pressure, divergence after it
    bc0 = np.zeros((3, 2))
    api.boundp(cbc, n, bc0, case.nh_d, 1, s.dl, s.dzc, s.dzf, pd)
    api.correc(*n, case.nh_d, case.nh_u, *s.dli, s.dzci, case.dt, case.rho0, pd, ud, vd, wd)
    refresh_velocity_halos_device(case, ud, vd, wd)
    divtot, divmax = api.chkdiv(*n, *s.dli, case.nh_d, case.nh_u, s.dzfi, ud, vd, wd)
    api.fftend(pl)
    out = {"err": float(err), "raw": float(raw), "divmax": float(divmax), "divtot": float(divtot),
           "divmax_before": float(div_before), "divmax_rel": float(divmax / div_before) if div_before > 0 else float(divmax),
           "divmax_oracle": float(divmax_oracle), "rhs_bit_exact": rhs_equal, "oracle_solve_s": round(t_cpu, 2),
           "oracle_threads": oracle.num_threads(), "input_gen_s": round(t_gen, 1),
           "metric": "max|p - p_oracle|/max|p_oracle| on p - mean(p); raw = without the gauge fix; divmax = chkdiv after correc"}
    del pd, ud, vd, wd
    torch.cuda.empty_cache()
    return out


def slab_solver_parity(case, api, comm, oracle, plan, normfft, threads=None, seed_offset=0):
    """N > 1: the z-slab solver (flutas_b200_solver_slab through `comm.solver`) against the single-rank oracle on the
    same bytes.  Every rank draws its slab of a random right-hand side on the device, made compatible with the singular
    operator (weighted mean removed); rank 0 gathers the slabs, runs oracle.Solver.solve on the whole grid and scatters
    the reference slabs back; the error is all-reduced.  Returns the same keys as pressure_step_parity where they apply."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    s, ng, cbc = case.setup, case.ng, case.cbc
    n1, n2, n3 = ng
    n3l = n3 // world
    k0 = rank * n3l
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(case.seed + 7919 * rank + seed_offset)
    rhs = torch.rand((n3l, n2, n1), dtype=torch.float64, device=dev, generator=g) - 0.5
    o = case.nh_d - 1
    wk = torch.from_numpy(np.ascontiguousarray(s.dzf[k0 + 1 + o:k0 + 1 + o + n3l])).to(dev)        # dzf(k), k = k0+1..k0+n3l
    sums = torch.stack([(rhs.sum(dim=(1, 2)) * wk).sum(), wk.sum() * float(n1 * n2)])
    dist.all_reduce(sums)
    if case.singular:
        rhs -= sums[0] / sums[1]
    # rank 0: whole right-hand side -> oracle
    parts = [torch.empty_like(rhs) for _ in range(world)] if rank == 0 else None
    dist.gather(rhs, parts, dst=0)
    t_cpu = 0.0
    refs = None
    if rank == 0:
        if threads:
            oracle.set_num_threads(int(threads))
        pg = np.zeros((n1 + 2, n2 + 2, n3 + 2), order="F")
        pgt = torch.from_numpy(pg.T)                                        # [k][j][i] view of the Fortran array
        for q in range(world):
            pgt[1 + q * n3l:1 + (q + 1) * n3l, 1:-1, 1:-1] = parts[q].cpu()
        del parts
        t0 = time.perf_counter()
        oracle.Solver(ng, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], pg)
        t_cpu = time.perf_counter() - t0
        refs = [pgt[1 + q * n3l:1 + (q + 1) * n3l, 1:-1, 1:-1].contiguous().to(dev) for q in range(world)]
    ref = torch.empty_like(rhs)
    dist.scatter(ref, refs, src=0)
    del refs
    # every rank: its slab through the slab solver
    pd = torch.zeros((n3l + 2, n2 + 2, n1 + 2), dtype=torch.float64, device=dev)
    pd[1:-1, 1:-1, 1:-1] = rhs
    del rhs
    j0, j1 = rank * (n2 // world), (rank + 1) * (n2 // world)
    lam_win = np.asfortranarray(s.lambdaxy[:, j0:j1])
    comm.solver((n1, n2, n3l), plan, normfft, lam_win, s.a, s.b, s.c, cbc[2], "ccc", pd)
    torch.cuda.synchronize()
    got = pd[1:-1, 1:-1, 1:-1]
    sums = torch.stack([got.sum(), ref.sum()])
    dist.all_reduce(sums)
    npts = float(n1) * n2 * n3
    mg, mr = sums[0] / npts, sums[1] / npts
    d = got - ref
    e = torch.stack([d.abs().max(), (d - (mg - mr)).abs().max(), ref.abs().max(), (ref - mr).abs().max()])
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    raw = float(e[0] / e[2])
    err = float(e[1] / e[3]) if case.singular else raw
    nerr = torch.tensor([comm.p2p_errors(plan)], device=dev)
    dist.all_reduce(nerr, op=dist.ReduceOp.MAX)
    t = torch.tensor([t_cpu], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"err": err, "raw": raw, "p2p_barrier_timeouts": int(nerr.item()), "oracle_solve_s": round(float(t.item()), 2),
            "oracle_threads": oracle.num_threads() if rank == 0 else None, "ranks": world,
            "metric": "max|p - p_oracle|/max|p_oracle| on p - mean(p) over all slabs; raw = without the gauge fix"}
