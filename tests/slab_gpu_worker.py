"""Launched by torchrun (one process per GPU): multi-GPU slab solve vs the single-rank oracle.
usage: torchrun --nproc-per-node N tests/slab_gpu_worker.py [nccl|p2p] """
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flutas_b200 import api, slab  # noqa: E402
from flutas_b200.cases import Case  # noqa: E402
from oracle import oracle  # noqa: E402
from util import gauge_rel_err  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl")
    api.init(lrank, rank, world)
    comm = slab.SlabComm()
    ok = True
    from flutas_b200 import lib
    transposes = mode.endswith("-transpose")                # p2p-transpose: the all-to-all path even where the distributed z solve applies
    if mode.startswith("p2p"):
        lib.check(lib.load().flutas_b200_slab_distributed_z(0 if transposes else 1))
        mode = "p2p"
    # the uni-* grids have exactly uniform z rows (lz = 1) and n3l a multiple of 16: with p2p they run the distributed z solve;
    # uni-fix widens the reference-order selection so that many whole columns travel through the gather / override path
    cases = (("chan", (128, 64, 32), ("PP", "PP", "NN"), 1.0, None), ("hit", (64, 64, 64), ("PP", "PP", "PP"), 0.0, None),
             ("rb", (64, 48, 40), ("NN", "NN", "NN"), 0.0, None), ("dd72", (32, 40, 72), ("DD", "NN", "DD"), 0.0, None),
             ("uni-chan", (64, 32, 128), ("PP", "PP", "NN"), 0.0, None), ("uni-per", (32, 32, 128), ("PP", "PP", "PP"), 0.0, None),
             ("uni-rb", (64, 64, 64), ("NN", "NN", "NN"), 0.0, None), ("uni-dn", (32, 64, 256), ("ND", "PP", "DD"), 0.0, None),
             ("uni-fix", (64, 32, 128), ("PP", "PP", "NN"), 0.0, 1.0e-2),
             # tanh-stretched z: the distributed z solve through the general kernel (coefficient tables)
             ("str-chan", (64, 32, 128), ("PP", "PP", "NN"), 2.0, None), ("str-dd", (32, 32, 256), ("NN", "PP", "DD"), 1.0, 1.0e-3))
    for name, ng, cbc, gr, ref_tol in cases:
        lib.check(lib.load().flutas_b200_debug_ref_tol(1.0e-5 if ref_tol is None else ref_tol))
        if ng[0] % world or ng[2] % world or ng[1] % world:
            continue
        case = Case(ng, cbc, (2.0, 1.0, 1.0), gr=gr, seed=77)
        s = case.setup
        n1, n2, n3 = ng
        n3l = n3 // world
        k0, k1 = slab.local_levels(n3, rank, world)
        u, v, w = case.velocity()
        pg = case.new_p()
        oracle.fillps(ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, pg)
        pref = pg.copy(order="F")
        oracle.Solver(ng, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], pref)
        pl = np.zeros((n1 + 2, n2 + 2, n3l + 2), order="F")
        pl[1:-1, 1:-1, 1:-1] = pg[1:-1, 1:-1, 1 + k0:1 + k1]
        nl = (n1, n2, n3l)
        plan, nf = api.fftini(nl, nl, (cbc[0], cbc[1]))
        if mode == "p2p":
            comm.use_p2p(plan, nl)
        else:
            comm.use_nccl_alltoall()
        j0, j1 = rank * (n2 // world), (rank + 1) * (n2 // world)
        lam_win = np.asfortranarray(s.lambdaxy[:, j0:j1])
        pd = api.device_field(pl)
        for _ in range(3):                                   # repeated solves exercise the barrier epochs
            pd.copy_(api.device_field(pl))
            comm.solver(nl, plan, nf, lam_win, s.a, s.b, s.c, cbc[2], "ccc", pd)
        torch.cuda.synchronize()
        got = api.host_field(pd, pl.shape)[1:-1, 1:-1, 1:-1]
        # gauge: compare after removing the GLOBAL mean (all-reduce of the local sums)
        ref = pref[1:-1, 1:-1, 1 + k0:1 + k1]
        if case.singular:
            sums = torch.tensor([got.sum(), ref.sum()], dtype=torch.float64, device="cuda")
            dist.all_reduce(sums)
            got = got - sums[0].item() / (n1 * n2 * n3)
            ref = ref - sums[1].item() / (n1 * n2 * n3)
        errs = torch.tensor([np.max(np.abs(got - ref)), np.max(np.abs(ref))], dtype=torch.float64, device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        err = errs[0].item() / errs[1].item()
        nerr = comm.p2p_errors(plan) if mode == "p2p" else 0
        if rank == 0:
            zpath = "distributed" if lib.load().flutas_b200_slab_last_distributed(plan.h) else "transposed"
            print("slab %s x%d %-8s %s %s: max|dp|/max|p| = %.2e barrier_timeouts=%d z=%s" % (mode, world, name, ng, "/".join(cbc), err, nerr, zpath),
                  flush=True)
        ok = ok and err <= 1e-12 and nerr == 0
        api.fftend(plan)
        # boundp on the slabs (z halo planes through NCCL send/recv): bit-exact against the single-rank oracle
        comm.use_halo_exchange()
        bc = np.array([[0.0, 0.0], [0.2, -0.1], [0.3, 0.7]]) * (np.array([c != "PP" for c in cbc])[:, None])
        rng = np.random.default_rng(5)
        full = np.asfortranarray(rng.uniform(-1, 1, (n1 + 2, n2 + 2, n3 + 2)))
        bref = oracle.boundp(cbc, ng, bc, s.nh_d, s.dl, s.dzc, full.copy(order="F"))
        mine = np.asfortranarray(full[:, :, k0:k0 + n3l + 2].copy())
        md = api.device_field(mine)
        dzc_loc = np.ascontiguousarray(s.dzc[k0:k0 + n3l + 2 * s.nh_d])
        api.boundp(cbc, nl, bc, s.nh_d, 1, s.dl, dzc_loc, dzc_loc, md)
        torch.cuda.synchronize()
        same = np.array_equal(api.host_field(md, mine.shape), bref[:, :, k0:k0 + n3l + 2])
        flag = torch.tensor([0 if same else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print("slab boundp x%d %-5s %s: %s" % (world, name, "/".join(cbc), "bit-exact" if flag.item() == 0 else "MISMATCH"), flush=True)
        ok = ok and flag.item() == 0
        # bounduvw (nh_u = 1 and 3) and the chkdt reduction on the slabs: bit-exact against the single-rank numpy oracle
        for nh in (1, 3):
            if n3l < nh + 1:
                continue
            nh_d = 3
            sh = (n1 + 2 * nh, n2 + 2 * nh, n3 + 2 * nh)
            rngv = np.random.default_rng(11 + nh)
            full = [np.asfortranarray(rngv.uniform(-1, 1, sh)) for _ in range(3)]
            dzc3 = rngv.uniform(0.05, 0.15, n3 + 2 * nh_d)
            dzf3 = rngv.uniform(0.05, 0.15, n3 + 2 * nh_d)
            vt = ["PP" if c == "PP" else "DD" for c in cbc]                   # no-slip walls where the pressure is not periodic
            vcbc = [[[d[ib]] * 3 for d in vt] for ib in (0, 1)]
            vbc = [[[0.0] * 3 for _ in range(3)] for _ in range(2)]
            if vt[2] == "DD":
                vbc[1][2][0] = 1.5                                             # moving top wall
            iso = [[False] * 3, [False] * 3]
            ref = [f.copy(order="F") for f in full]
            oracle.bounduvw(vcbc, ng, vbc, nh_d, nh, iso, s.dl, dzc3, dzf3, *ref)
            mine = [np.asfortranarray(f[:, :, k0:k0 + n3l + 2 * nh].copy()) for f in full]
            md3 = [api.device_field(f) for f in mine]
            dzc_l = np.ascontiguousarray(dzc3[k0:k0 + n3l + 2 * nh_d])
            dzf_l = np.ascontiguousarray(dzf3[k0:k0 + n3l + 2 * nh_d])
            api.bounduvw(vcbc, nl, vbc, nh_d, nh, iso, s.dl, dzc_l, dzf_l, *md3)
            torch.cuda.synchronize()
            same = all(np.array_equal(api.host_field(t, m.shape), r[:, :, k0:k0 + n3l + 2 * nh]) for t, m, r in zip(md3, mine, ref))
            dti_ref = oracle.chkdt_dti(ng, s.dli, nh_d, nh, 1.0 / dzc3, 1.0 / dzf3, *ref)
            dti = comm.chkdt(n1, n2, n3l, *s.dli, nh_d, nh, np.ascontiguousarray(1.0 / dzc_l), np.ascontiguousarray(1.0 / dzf_l), *md3)
            flag = torch.tensor([0 if (same and dti == dti_ref) else 1], device="cuda")
            dist.all_reduce(flag)
            if rank == 0:
                print("slab bounduvw nh=%d + chkdt x%d %-5s: %s" % (nh, world, name, "bit-exact" if flag.item() == 0 else "MISMATCH"), flush=True)
            ok = ok and flag.item() == 0
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if rank == 0:
        print("SLAB_OK", mode, world, flush=True)


if __name__ == "__main__":
    main()
