"""world_size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: slab partition, the one-off eigenvalue
all-gather, and the two exchange layouts (numpy mirrors of the kernels' SpecGeom/ColGeom addressing).
Each rank runs the ORACLE's transforms/Thomas on its slab, exchanges through gloo with exactly the buffer
layout the CUDA path uses, and the assembled result must equal the single-rank oracle solve bit for bit
(decomposition independence)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flutas_b200 import slab
from flutas_b200.cases import Case
from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cbc, ng, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle.set_num_threads(1)
        case = Case(ng, cbc, (2.0, 1.0, 1.0), gr=(1.0 if cbc[2] != "PP" else 0.0), seed=31)
        s = case.setup
        n1, n2, n3 = ng
        n1l, n3l = n1 // world, n3 // world
        k0, k1 = slab.local_levels(n3, rank, world)
        u, v, w = case.velocity()
        pg = case.new_p()
        oracle.fillps(ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, pg)
        rhs_slab = np.asfortranarray(pg[1:-1, 1:-1, 1 + k0:1 + k1])
        comm = slab.SlabComm()
        # the window initsolver hands this rank (all x, its share of y) -> global lambda
        j0, j1 = rank * (n2 // world), (rank + 1) * (n2 // world)
        lam_full = comm.gather_lambda(np.asfortranarray(s.lambdaxy[:, j0:j1]))
        assert np.array_equal(lam_full, s.lambdaxy)
        kfx, kbx, _ = oracle.find_fft(cbc[0])
        kfy, kby, _ = oracle.find_fft(cbc[1])
        w1 = rhs_slab.copy(order="F")
        oracle.r2r(kfx, w1, 0)
        oracle.r2r(kfy, w1, 1)
        send = torch.from_numpy(slab.pack_spec(w1, world))
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        pencil = slab.pencil_from_recv(recv.numpy(), n1l, n2, n3l, world).copy(order="F")
        lam_win = np.asfortranarray(lam_full[rank * n1l:(rank + 1) * n1l, :])
        oracle.gaussel(s.a, s.b, s.c, lam_win, pencil, cbc[2] == "PP")
        send2 = torch.from_numpy(pencil.ravel(order="F").copy())        # chunk q = levels of rank q: contiguous
        recv2 = torch.empty_like(send2)
        dist.all_to_all_single(recv2, send2)
        w1 = slab.unpack_spec(recv2.numpy(), n1, n2, n3l, world)
        oracle.r2r(kby, w1, 1)
        oracle.r2r(kbx, w1, 0)
        np.save(os.path.join(out_dir, "slab_%d.npy" % rank), w1 * s.normfft)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("cbc", [("PP", "PP", "NN"), ("NN", "DD", "PP")], ids=lambda c: "".join(c))
def test_decomposed_oracle_equals_single_rank(tmp_path, world, cbc):
    ng = (16, 12, 8)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, cbc, ng, str(tmp_path)), nprocs=world, join=True)
    case = Case(ng, cbc, (2.0, 1.0, 1.0), gr=(1.0 if cbc[2] != "PP" else 0.0), seed=31)
    s = case.setup
    u, v, w = case.velocity()
    p = case.new_p()
    oracle.set_num_threads(1)
    oracle.fillps(ng, case.nh_d, case.nh_u, s.dli, s.dzfi, case.dti, case.rho0, u, v, w, p)
    oracle.Solver(ng, cbc[0], cbc[1]).solve(s.lambdaxy, s.a, s.b, s.c, cbc[2], p)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "slab_%d.npy" % r)) for r in range(world)], axis=2)
    assert np.array_equal(got, p[1:-1, 1:-1, 1:-1])


def test_pack_unpack_roundtrip_layout():
    rng = np.random.default_rng(0)
    n1, n2, n3l, P = 8, 3, 2, 4
    slabs = [np.asfortranarray(rng.uniform(size=(n1, n2, n3l))) for _ in range(P)]
    sends = [slab.pack_spec(w, P) for w in slabs]
    chunk = (n1 // P) * n2 * n3l
    # emulate the all-to-all: rank q receives chunk q of every rank r at position r
    recvs = [np.concatenate([sends[r][q * chunk:(q + 1) * chunk] for r in range(P)]) for q in range(P)]
    glob = np.concatenate(slabs, axis=2)
    for q in range(P):
        pen = slab.pencil_from_recv(recvs[q], n1 // P, n2, n3l, P)
        assert np.array_equal(pen, glob[q * (n1 // P):(q + 1) * (n1 // P)])
    # backward: pencil q sends its level range r to rank r
    back = [np.concatenate([slab.pencil_from_recv(recvs[r], n1 // P, n2, n3l, P)[:, :, q * n3l:(q + 1) * n3l].ravel(order="F")
                            for r in range(P)]) for q in range(P)]
    for q in range(P):
        assert np.array_equal(slab.unpack_spec(back[q], n1, n2, n3l, P), slabs[q])


# ---- boundp on z-slabs: the halo exchange of flutas_b200.slab (same op ordering as the NCCL callback) --------
def _boundp_worker(rank, world, port, cbc, ng, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(ng, cbc, (2.0, 1.0, 1.0), gr=(1.0 if cbc[2] != "PP" else 0.0), seed=8)
        s = case.setup
        n1, n2, n3 = ng
        n3l = n3 // world
        k0, _ = slab.local_levels(n3, rank, world)
        rng = np.random.default_rng(12)
        full = np.asfortranarray(rng.uniform(-1, 1, (n1 + 2, n2 + 2, n3 + 2)))
        mine = np.asfortranarray(full[:, :, k0:k0 + n3l + 2].copy())
        bc = np.array([[0.0, 0.0], [0.2, -0.1], [0.3, 0.7]]) * (np.array([c != "PP" for c in cbc])[:, None])
        # what flutas_b200_boundp does on a slab: y halo, then the z planes through the exchange, then set_bc
        if cbc[1] == "PP":
            oracle._set_bc(mine, "P", 0, 1, 0.0, 0.0)
        lo, hi = slab.z_neighbours(rank, world, cbc[2] == "PP")
        send_lo = torch.from_numpy(np.ascontiguousarray(mine[:, :, 1].T))
        send_hi = torch.from_numpy(np.ascontiguousarray(mine[:, :, n3l].T))
        recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
        ops = slab.halo_ops(dist, send_lo, send_hi, recv_lo, recv_hi, lo, hi)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        mine2 = np.asfortranarray(full[:, :, k0:k0 + n3l + 2].copy())
        dzc_loc = s.dzc[k0:k0 + n3l + 2 * s.nh_d]
        oracle.boundp(cbc, (n1, n2, n3l), bc, s.nh_d, s.dl, dzc_loc, mine2,
                      below=recv_lo.numpy().T if lo >= 0 else None, above=recv_hi.numpy().T if hi >= 0 else None,
                      first_rank=(rank == 0), last_rank=(rank == world - 1))
        np.save(os.path.join(out_dir, "bp_%d.npy" % rank), mine2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("cbc", [("PP", "PP", "PP"), ("NN", "PP", "ND"), ("PP", "DD", "NN")], ids=lambda c: "".join(c))
def test_boundp_slabs_gloo(tmp_path, world, cbc):
    ng = (6, 8, 12)
    mp.spawn(_boundp_worker, args=(world, _free_port(), cbc, ng, str(tmp_path)), nprocs=world, join=True)
    case = Case(ng, cbc, (2.0, 1.0, 1.0), gr=(1.0 if cbc[2] != "PP" else 0.0), seed=8)
    s = case.setup
    rng = np.random.default_rng(12)
    full = np.asfortranarray(rng.uniform(-1, 1, (ng[0] + 2, ng[1] + 2, ng[2] + 2)))
    bc = np.array([[0.0, 0.0], [0.2, -0.1], [0.3, 0.7]]) * (np.array([c != "PP" for c in cbc])[:, None])
    ref = oracle.boundp(cbc, ng, bc, s.nh_d, s.dl, s.dzc, full)
    n3l = ng[2] // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "bp_%d.npy" % r))
        assert np.array_equal(got, ref[:, :, r * n3l:r * n3l + n3l + 2]), r
