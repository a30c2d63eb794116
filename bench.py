#!/usr/bin/env python
"""bench.py -- FP64 Poisson-solve throughput of the FluTAS pressure path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl reference]

A "step" is one `solver` call (x/y transforms + z tridiagonal solve + inverse transforms) on one
synthetic right-hand side of the named grid, device resident, in place.  Default workload: NS, the north-star
1024^3 PP/PP/NN channel grid, for every N.  `value` = grid points / step time (Gpts/s, whole job).  `e2e` is the
same call through the C ABI with HOST (pinned) buffers, i.e. including the host->device and device->host copies of
p.  `roofline` describes the slowest kernel of the solve, timed live with CUDA events on the launching stream.
`parity` is measured in the same run: the CUDA pressure step against `oracle.Solver.solve` on the same bytes
(oracle/parity.py; N > 1: the slab solver against the single-rank oracle).  `cpu_baseline` is that oracle solve (a
restatement of solver_cpu.f90 -- NOT FluTAS+FFTW, which cannot be built here), all host cores.
`--impl reference` times real oracle pressure steps (fillps + solver_cpu + correc) on the whole workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_PT = {"xfft_fwd": 16, "yfft_fwd": 16, "thomas_z": 16, "yfft_bwd": 16, "xfft_bwd": 16,
                    "fillps": 32, "correc": 56,
                    "thomas_z_corr": 16}                 # second local sweep of the distributed z solve (N > 1 only)
SOLVER_BYTES_PER_PT = 80


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy peak)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + ",".join(self.FIELDS),
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.time()])

    def stop(self, t0=None, t1=None):
        """t0, t1: host-clock window of the load; samples outside it (the sampler starts before the warm-up so that
        nvidia-smi is already streaming when the timed region begins) are dropped."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if t0 is not None and not (t0 <= r[-1] <= t1 + 0.12):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# Reference arm: the CPU oracle (restatement of fillps.f90 / solver_cpu.f90 / correc.f90) on the WHOLE workload, all host
# threads, real calls inside the timed region.  Under torchrun the launcher exports OMP_NUM_THREADS=1: the thread count
# is set explicitly.
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    from flutas_b200.cases import Case
    from oracle import oracle
    oracle.set_num_threads(host_threads())
    case = Case.from_config(args.workload, gr=args.gr)
    s, n = case.setup, case.ng
    n1, n2, n3 = n
    npts = n1 * n2 * n3
    u, v, w = case.velocity()
    p = case.new_p()
    solver = oracle.Solver(n, case.cbc[0], case.cbc[1])

    def fillps():
        oracle.fillps(n, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, p)
        oracle.updt_rhs_b(n, s.rhsbx, s.rhsby, s.rhsbz, p)

    def solve():
        solver.solve(s.lambdaxy, s.a, s.b, s.c, case.cbc[2], p)

    fillps()
    rhs = p.copy(order="F")
    for _ in range(args.warmup):
        p[...] = rhs
        solve()
    t_solve = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        p[...] = rhs                                   # same right-hand side every step (not timed: the GPU arm solves in place too)
        t0 = time.perf_counter()
        solve()
        t_solve.append(time.perf_counter() - t0)
    wall = time.perf_counter() - t_wall0
    # the rest of the pressure step, a few repetitions: fillps + updt_rhs_b and correc (on the last solution)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.correc(n, case.nh_d, case.nh_u, s.dli, s.dzci, case.dt, case.rho0, case.boundp(p), u, v, w)
    t_correc = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        fillps()
    t_fillps = (time.perf_counter() - t0) / reps
    ms = 1e3 * float(np.mean(t_solve))
    gpts = npts / (ms * 1e-3) / 1e9
    sample = "%d full oracle.Solver.solve calls on the %dx%dx%d grid (no sampling)" % (args.steps, n1, n2, n3)
    line = {"impl": "reference", "metric": "poisson_solve_throughput", "value": gpts, "unit": "Gpts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(case, args.workload),
            "ms_per_pressure_step": ms + 1e3 * (t_fillps + t_correc),
            "pressure_step": "fillps + updt_rhs_b %.1f ms, solver %.1f ms, boundp + correc %.1f ms" % (1e3 * t_fillps, ms, 1e3 * t_correc),
            "cpu_baseline": {"value": gpts, "unit": "Gpts/s", "cores": oracle.num_threads(), "kind": "port",
                             "sample": sample,
                             "note": "CPU restatement of solver_cpu.f90 with its own FFT (oracle/), not FluTAS+FFTW: "
                                     "no Fortran/MPI/FFTW toolchain in this image"},
            "e2e": {"value": gpts, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line))


NOMINAL_HBM_GBS = 8000.0                                 # the north star's nominal figure, reported next to the measured peak


def nvlink_figures(stage_tbl, stages, npts_loc, world, exchange, ncol=None):
    """NVLink side of the roofline.  Transpose path: bytes each GPU sends per exchange and the rate the stage
    that carries them achieves (direct stores: the y-transform / z-solve kernels themselves; NCCL: the all-to-all).
    Distributed z solve: only boundary planes cross the link -- 32 bytes per column per solve."""
    nv = {"exchange": exchange, "peak_GBs_per_direction": 900.0, "measured_peer_copy_GBs": 770.0}
    if "z_interface" in stages and stages["z_interface"][1]:
        nv["z_path"] = "distributed z solve (dz.cuh): no all-to-all transposes"
        nv["z_interface_ms"] = round(stages["z_interface"][0] / stages["z_interface"][1], 4)
        if ncol:
            nv["bytes_sent_per_gpu_per_solve"] = 32.0 * ncol * (world - 1) / world
            nv["bytes_a_transpose_pair_would_send"] = 16.0 * npts_loc * (world - 1) / world
        return nv
    sent = 8.0 * npts_loc * (world - 1) / world
    nv["z_path"] = "two fused all-to-all transposes"
    nv["bytes_sent_per_gpu_per_exchange"] = sent
    if exchange == "p2p":
        for nm in ("yfft_fwd", "thomas_z"):
            if nm in stage_tbl:
                nv[nm + "_GBs"] = round(sent / (stage_tbl[nm]["ms"] * 1e-3) / 1e9, 1)
    for nm in ("exchange_fwd", "exchange_bwd"):
        if nm in stages and stages[nm][1]:
            ms = stages[nm][0] / stages[nm][1]
            nv[nm + "_ms"] = round(ms, 4)
            if exchange == "nccl":
                nv[nm + "_GBs"] = round(sent / (ms * 1e-3) / 1e9, 1)
    return nv


def workload_config(case, wid):
    from flutas_b200.cases import CONFIGS
    n1, n2, n3 = case.ng
    return {"workload": "%s: %s" % (wid, CONFIGS[wid]["desc"]), "grid": [n1, n2, n3], "pressure_bc": "/".join(case.cbc),
            "z_grid": "uniform" if case.setup.gr == 0.0 else "tanh-stretched, gr = %g" % case.setup.gr,
            "l2_policy": "inputs larger than L2: one FP64 field is %.2f GB vs 126 MB of L2" % (8e-9 * n1 * n2 * n3),
            "step": "one solver call (5 kernels), in place, device resident"}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="NS")
    ap.add_argument("--gr", type=float, default=0.0, help="tanh stretching of the z grid (initgrid.f90:17-97); 0 = uniform, as in every BASELINE config")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the oracle run (parity + cpu_baseline keys become null): kernel A/B timing only")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--solver-only", action="store_true",
                    help="time only the solver on a device-generated RHS (large shapes: no host-side velocity fields, "
                         "no pressure-step / e2e / CPU legs); prints a reduced JSON line")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: direct NVLink stores from the kernels (CUDA IPC) or NCCL all-to-all")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from flutas_b200 import api
    from flutas_b200.cases import Case

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the pressure path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.init(local_rank, rank, world)

    case = Case.from_config(args.workload, gr=args.gr)
    s = case.setup
    ng = case.ng
    n1, n2, n3g = ng
    if n3g % world or n1 % world or n2 % world:
        raise SystemExit("grid %s is not divisible by %d ranks" % (ng, world))
    n3 = n3g // world                                   # z-slab: the reference's _DECOMP_X layout with dims_in = (1, N)
    n = (n1, n2, n3)
    npts = n1 * n2 * n3g                                # whole job
    npts_loc = n1 * n2 * n3
    exchange = None
    parity = None
    if args.solver_only:
        ud = vd = wd = None
        dzfi, dzci = s.dzfi, s.dzci
    elif world == 1:
        u, v, w = case.velocity()
        dzfi, dzci = s.dzfi, s.dzci
    else:
        # per-rank slab of synthetic velocities (timing is data independent; parity is covered by tests/)
        rng = np.random.Generator(np.random.PCG64(case.seed + 1000 * rank))
        h = case.nh_u
        shape = (n1 + 2 * h, n2 + 2 * h, n3 + 2 * h)
        u, v, w = (np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape[::-1]).T) for _ in range(3))
        o = case.nh_d - 1
        k0 = rank * n3
        dzfi = np.ascontiguousarray(s.dzfi[k0:k0 + n3 + 2 * case.nh_d])
        dzci = np.ascontiguousarray(s.dzci[k0:k0 + n3 + 2 * case.nh_d])
    if args.solver_only:
        g = torch.Generator(device="cuda")
        g.manual_seed(case.seed + rank)
        pd = torch.rand((n3 + 2, n2 + 2, n1 + 2), dtype=torch.float64, device="cuda", generator=g) - 0.5
        rhs0 = pd.clone()
    else:
        ud, vd, wd = (api.device_field(f) for f in (u, v, w))
        if world == 1 and not args.no_parity:
            # parity of the whole pressure step against the CPU oracle on the same bytes (also the cpu_baseline figure)
            from oracle import oracle, parity as oparity
            parity = oparity.pressure_step_parity(case, api, oracle, threads=host_threads(), host_fields=(u, v, w),
                                                  dev_fields=(ud, vd, wd))
            parity["ok"] = bool(parity["rhs_bit_exact"] and parity["err"] <= 1e-12 and parity["divmax_rel"] <= 1e-12 and
                                parity["divmax"] <= 2.0 * parity["divmax_oracle"] + 1e-13)
            parity["bar"] = ("err <= 1e-12; divmax <= 1e-12 max|div u*| and <= 2 x the oracle's own post-correc divmax "
                             "(absolute floor eps |p| dt / dz^2, see oracle/parity.py); right-hand side bit-exact")
        del u, v, w
        pd = api.device_field(np.zeros((n1 + 2, n2 + 2, n3 + 2), order="F"))
    pl, nf = api.fftini(n, n, (case.cbc[0], case.cbc[1]))
    lam_win = s.lambdaxy
    comm = None
    if world > 1:
        from flutas_b200 import slab
        comm = slab.SlabComm()
        j0, j1 = rank * (n2 // world), (rank + 1) * (n2 // world)
        lam_win = np.asfortranarray(s.lambdaxy[:, j0:j1])
        exchange = args.exchange
        if exchange == "p2p":
            try:
                comm.use_p2p(pl, n)
            except Exception as e:                       # e.g. CUDA IPC not permitted: fall back to the NCCL all-to-all
                if rank == 0:
                    print("p2p exchange unavailable (%s): using NCCL all-to-all" % e, file=sys.stderr)
                exchange = "nccl"
        if exchange == "nccl":
            comm.use_nccl_alltoall()
        comm.use_halo_exchange()                         # z-halo planes of boundp (updthalo, bound.f90:946-1110)
        if not args.no_parity and not args.solver_only:
            # the slab solver against the single-rank oracle on the same bytes, every run (+ barrier time-outs)
            from oracle import oracle, parity as oparity
            parity = oparity.slab_solver_parity(case, api, comm, oracle, pl, nf, threads=host_threads())
            parity["ok"] = bool(parity["err"] <= 1e-12 and parity["p2p_barrier_timeouts"] == 0)
            parity["bar"] = "err <= 1e-12, no cross-GPU barrier time-out"
    slab_schedule = None
    if world > 1 and exchange == "p2p" and not os.environ.get("FLUTAS_B200_PIPE"):
        # plan-time measurement of the forward-half pipelining (x transform of k-chunk c+1 under the y transform + NVLink
        # stores of chunk c), collectively on all ranks; FLUTAS_B200_PIPE=<chunks> pins it instead
        tune_p = torch.zeros((n3 + 2, n2 + 2, n1 + 2), dtype=torch.float64, device="cuda")
        best, times = comm.autotune(lambda: comm.solver(n, pl, nf, lam_win, s.a, s.b, s.c, case.cbc[2], "ccc", tune_p))
        del tune_p
        slab_schedule = {"distributed_z": bool(best[0]), "pipe_chunks": best[1], "pipe_xsm_pct": best[2],
                         "candidates_ms": {"dz=%d pipe=%d/%d" % k: round(v, 4) for k, v in times.items()},
                         "note": "distributed_z: z stage = two local sweeps around a 2N x 2N interface system per column, no "
                                 "all-to-all transposes (flutas_b200/csrc/dz.cuh); else the fused-transpose path"}
    k0 = rank * n3
    rhsbx = np.asfortranarray(s.rhsbx[:, k0:k0 + n3, :])     # boundary constants of this rank's slab (bound.f90:829-944)
    rhsby = np.asfortranarray(s.rhsby[:, k0:k0 + n3, :])
    dzc_loc = np.ascontiguousarray(s.dzc[k0:k0 + n3 + 2 * case.nh_d])
    # the boundary-constant arrays are uploaded once, like the fields (a device-resident time loop does the same)
    rhsb_dev = tuple(api.device_field(np.asfortranarray(a)) for a in (rhsbx, rhsby, s.rhsbz))

    def fill():
        if args.solver_only:
            pd.copy_(rhs0)
            return
        api.fillps(*n, case.nh_d, case.nh_u, *s.dli, dzfi, case.dti, case.rho0, ud, vd, wd, pd)
        api.updt_rhs_b(*n, case.cbc, *rhsb_dev, pd)                  # z faces: ranks 0 and N-1 only (library gates them)

    def solve(p=None):
        p = pd if p is None else p
        if world == 1:
            api.solver(n, pl, nf, s.lambdaxy, s.a, s.b, s.c, case.cbc[2], "ccc", p)
        else:
            comm.solver(n, pl, nf, lam_win, s.a, s.b, s.c, case.cbc[2], "ccc", p)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: K solver calls ----------------------------------------------------------
    fill()
    sampler = ClockSampler(local_rank)                  # started before the warm-up: nvidia-smi needs ~0.2 s to its first sample
    sampler.start()
    for _ in range(args.warmup):
        solve()
    barrier()
    api.profile_enable(True)
    api.profile_read()
    l0 = api.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_load0 = time.time()
    e0.record()
    for _ in range(args.steps):
        solve()
    e1.record()
    barrier()
    t_load1 = time.time()
    launches = api.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    if world > 1:                                       # max over ranks, on the device clock
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    stages = api.profile_read()
    api.profile_enable(False)
    # A timed region shorter than a few sampling periods (8 GPUs: 20 solves = 60 ms) would leave no clock sample: keep the
    # same load running, untimed, for ~0.6 s more.  The number of extra solves follows from ms_total, which is identical on
    # every rank after the all-reduce (the slab solve is collective).
    extra_steps = 0
    if ms_total < 400.0:
        extra_steps = int(min(2000, max(1, 600.0 / max(ms_total / args.steps, 1e-3))))
        for _ in range(extra_steps):
            solve()
        barrier()
        t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1)
    if extra_steps:
        clocks["window"] = "timed region (%.0f ms) + %d untimed solves of the same load" % (ms_total, extra_steps)
    ms_step = ms_total / args.steps
    value = npts / (ms_step * 1e-3) / 1e9

    if args.solver_only:
        peak, peak_src = measured_peak()
        tbl = {}
        for name, (ms, cnt) in stages.items():
            if name in ALG_BYTES_PER_PT and cnt:
                gbs = ALG_BYTES_PER_PT[name] * npts_loc / (ms / cnt * 1e-3) / 1e9
                tbl[name] = {"ms": round(ms / cnt, 4), "GB/s": round(gbs, 1), "frac": round(gbs / peak, 4)}
            elif cnt:
                tbl[name] = {"ms": round(ms / cnt, 4)}
        line = {"metric": "poisson_solve_throughput", "value": round(value, 3), "unit": "Gpts/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
                "scaling": "strong", "dtype": "f64", "data": "synthetic (device-generated uniform RHS, solver only)",
                "config": workload_config(case, args.workload),
                "decomposition": ("z-slabs over %d GPUs, exchange=%s" % (world, exchange)) if world > 1 else "single GPU",
                "slab_schedule": slab_schedule,
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                             "solver": {"bytes_per_pt": SOLVER_BYTES_PER_PT,
                                        "achieved_per_gpu": round(SOLVER_BYTES_PER_PT * value / world, 1),
                                        "frac": round(SOLVER_BYTES_PER_PT * value / world / peak, 4),
                                        "frac_of_nominal_8TBs": round(SOLVER_BYTES_PER_PT * value / world / NOMINAL_HBM_GBS, 4)},
                             "stages": tbl}}
        if world > 1:
            line["roofline"]["nvlink"] = nvlink_figures(tbl, stages, npts_loc, world, exchange, n1 * n2)
        if rank == 0:
            print(json.dumps(line))
        api.fftend(pl)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- ms per pressure step (fillps + updt_rhs_b + solver + correc), device resident -----------
    api.profile_enable(True)
    api.profile_read()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nps = max(3, min(args.steps, 10))
    bc0 = np.zeros((3, 2))

    def pressure_step():
        fill()
        solve()
        api.boundp(case.cbc, n, bc0, case.nh_d, 1, s.dl, dzc_loc, dzc_loc, pd)     # ghost cells of p (bound.f90:146); N > 1: z halo over NCCL
        api.correc(*n, case.nh_d, case.nh_u, *s.dli, dzci, case.dt, case.rho0, pd, ud, vd, wd)

    pressure_step()                                      # untimed: first use of the halo exchange creates its NCCL channels
    barrier()
    api.profile_read()
    p0.record()
    for _ in range(nps):
        pressure_step()
    p1.record()
    barrier()
    ms_pressure_step = p0.elapsed_time(p1) / nps
    if world > 1:
        t = torch.tensor([ms_pressure_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_pressure_step = float(t.item())
    stages_ps = api.profile_read()
    api.profile_enable(False)
    for k in ("fillps", "correc"):
        if k in stages_ps:
            stages[k] = stages_ps[k]

    # ---- e2e: the same solver call with host (pinned) p ----------------------------------------
    pcount = (n1 + 2) * (n2 + 2) * (n3 + 2)
    ph_t = torch.empty(pcount, dtype=torch.float64).pin_memory()
    ph = ph_t.numpy().reshape((n1 + 2, n2 + 2, n3 + 2), order="F")
    fill()
    torch.cuda.synchronize()
    rhs_host = api.host_field(pd, ph.shape)
    e2e_ms = []
    for it in range(args.e2e_steps + 1):
        ph[...] = rhs_host
        barrier()
        t0 = time.perf_counter()
        solve(ph)                                        # host p: H2D + kernels (+ exchanges) + D2H, returns when p is back
        barrier()
        dt = time.perf_counter() - t0
        if it > 0:
            e2e_ms.append(dt * 1e3)
    e2e_mean = float(np.mean(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_mean = float(t.item())
    e2e_val = npts / (e2e_mean * 1e-3) / 1e9

    # ---- roofline of the slowest solver kernel ---------------------------------------------------
    peak, peak_src = measured_peak()
    stage_tbl = {}
    for name, (ms, cnt) in stages.items():
        avg = ms / cnt
        if name not in ALG_BYTES_PER_PT:
            continue
        gbs = ALG_BYTES_PER_PT[name] * npts_loc / (avg * 1e-3) / 1e9
        stage_tbl[name] = {"ms": round(avg, 4), "GB/s": round(gbs, 1), "frac": round(gbs / peak, 4)}
    solver_stages = [k for k in ("xfft_fwd", "yfft_fwd", "thomas_z", "thomas_z_corr", "yfft_bwd", "xfft_bwd") if k in stage_tbl]
    dom = max(solver_stages, key=lambda k: stage_tbl[k]["ms"])
    traffic = None                                       # per-launch DRAM bytes from the committed 1-GPU ncu capture
    try:
        if world > 1:
            raise LookupError("ncu captures are single-GPU")
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stage_tbl[dom]["GB/s"], "peak": peak, "unit": "GB/s",
                "frac": stage_tbl[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_PT[dom] * npts_loc,
                "solver": {"bytes_per_pt": SOLVER_BYTES_PER_PT, "achieved_per_gpu": round(SOLVER_BYTES_PER_PT * value / world, 1),
                           "frac": round(SOLVER_BYTES_PER_PT * value / world / peak, 4),
                           "frac_of_nominal_8TBs": round(SOLVER_BYTES_PER_PT * value / world / NOMINAL_HBM_GBS, 4)},
                "stages": stage_tbl}
    if world > 1:
        roofline["nvlink"] = nvlink_figures(stage_tbl, stages, npts_loc, world, exchange, n1 * n2)

    cpu = None
    if parity is not None and parity.get("oracle_solve_s"):
        # the oracle solve of the parity run: ONE full solver_cpu restatement call on the whole workload, all host threads
        cpu = {"value": round(npts / parity["oracle_solve_s"] / 1e9, 5), "unit": "Gpts/s", "cores": parity["oracle_threads"],
               "kind": "port", "sample": "1 full oracle.Solver.solve call on the %dx%dx%d grid (the parity run; no sampling)" % ng,
               "seconds": parity["oracle_solve_s"],
               "note": "CPU restatement of solver_cpu.f90 with its own FFT (oracle/), not FluTAS+FFTW"}
    if parity is not None:
        parity = {k: v for k, v in parity.items() if not k.startswith("_")}

    line = {"metric": "poisson_solve_throughput", "value": round(value, 3), "unit": "Gpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(case, args.workload),
            "decomposition": ("z-slabs over %d GPUs, exchange=%s" % (world, exchange)) if world > 1 else "single GPU",
            "slab_schedule": slab_schedule,
            "parity": parity,
            "ms_per_pressure_step": round(ms_pressure_step, 4),
            "pressure_step": "fillps + updt_rhs_b + solver + boundp + correc, device resident" +
                             ("" if world == 1 else " (boundp's z-halo planes through NCCL send/recv)"),
            "e2e": {"value": round(e2e_val, 4), "unit": "Gpts/s", "h2d_bytes_per_step": pcount * 8,
                    "d2h_bytes_per_step": pcount * 8, "ms_per_step": round(e2e_mean, 3),
                    "path": "flutas_b200_solver with a pinned host p (H2D + 5 kernels + D2H)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    if rank == 0:
        print(json.dumps(line))
    api.fftend(pl)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
