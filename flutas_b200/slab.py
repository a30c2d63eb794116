"""Multi-GPU host side of the pressure solver: z-slab decomposition, one process per GPU.

Mirrors what the reference's MPI layer does around `solver` for `dims_in = (1, nranks)` (src/initmpi.f90:87-104):
each rank owns p(0:ng1+1, 0:ng2+1, 0:ng3/nranks+1).  torch.distributed is the plumbing (rendezvous, the
one-off eigenvalue all-gather, the NCCL all-to-all when the direct NVLink path is not attached, scalar
all-reduces for chkdiv).  The exchange layout is documented -- and exercised on CPU by tests/test_slab_gloo.py --
through the numpy mirrors `pack_spec`, `unpack_spec`.
"""
import ctypes as C

import numpy as np

from . import api
from . import lib as _lib


# ---- layout of the two exchanges (numpy mirrors of SpecGeom / ColGeom in csrc/geom.cuh) -----------------
def pack_spec(w1, nranks):
    """slab (ng1, ng2, n3l) -> send buffer of nranks chunks (n1l, ng2, n3l): chunk q holds x rows of rank q.
    This is what the forward y-transform kernel writes (the pack of transpose_xc_to_z fused into its store)."""
    n1, n2, n3l = w1.shape
    n1l = n1 // nranks
    return np.concatenate([w1[q * n1l:(q + 1) * n1l].ravel(order="F") for q in range(nranks)])


def pencil_from_recv(recv, n1l, n2, n3l, nranks):
    """receive buffer (chunk r = levels of rank r) viewed as the pencil (n1l, ng2, ng3): no unpack needed."""
    return recv.reshape((n1l, n2, n3l * nranks), order="F")


def unpack_spec(recv, n1, n2, n3l, nranks):
    """receive buffer of the backward exchange (chunk r = x rows of rank r, this rank's levels) -> slab.
    This is what the inverse y-transform kernel reads (the unpack of transpose_z_to_xc fused into its load)."""
    n1l = n1 // nranks
    chunk = n1l * n2 * n3l
    w1 = np.empty((n1, n2, n3l), order="F")
    for r in range(nranks):
        w1[r * n1l:(r + 1) * n1l] = recv[r * chunk:(r + 1) * chunk].reshape((n1l, n2, n3l), order="F")
    return w1


def local_levels(ng3, rank, nranks):
    n3l = ng3 // nranks
    return rank * n3l, (rank + 1) * n3l


# ---- torch.distributed plumbing --------------------------------------------------------------------------
class _DevPtr:
    """a raw device pointer seen as a 1-D float64 CUDA array (zero copy)"""

    def __init__(self, ptr, ndoubles):
        self.__cuda_array_interface__ = {"shape": (ndoubles,), "typestr": "<f8", "data": (ptr, False), "version": 3}


_A2A_PROTO = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
_HALO_PROTO = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                          C.c_void_p)


def z_neighbours(rank, nranks, periodic):
    """bottom / top ranks of a z-slab (MPI_CART_SHIFT along the decomposed direction, src/initmpi.f90:127);
    -1 = MPI_PROC_NULL"""
    lo = rank - 1 if rank > 0 else (nranks - 1 if periodic else -1)
    hi = rank + 1 if rank < nranks - 1 else (0 if periodic else -1)
    return lo, hi


def halo_ops(dist, send_lo, send_hi, recv_lo, recv_hi, lo, hi, group=None):
    """The MPI_SENDRECV pair of updthalo (src/bound.f90:1098-1103) as one batch of point-to-point operations.
    Messages between the same two ranks match in posting order, so when lo == hi (two ranks, periodic) the
    receives are posted top-halo first: the peer's first send is its bottom plane, which is this rank's top halo."""
    ops = []
    if lo >= 0:
        ops.append(dist.P2POp(dist.isend, send_lo, lo, group))
    if hi >= 0:
        ops.append(dist.P2POp(dist.isend, send_hi, hi, group))
    if hi >= 0:
        ops.append(dist.P2POp(dist.irecv, recv_hi, hi, group))
    if lo >= 0:
        ops.append(dist.P2POp(dist.irecv, recv_lo, lo, group))
    return ops


class SlabComm:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        self._cb = None
        self._lam_cache = {}

    # eigenvalues: every rank holds the window lambdaxy(ng1, ng2/nranks) (src/initsolver.f90:87-93); the x-split
    # z stage needs all of y for a range of x -> all-gather once along y.
    def gather_lambda(self, lam_window):
        import torch
        key = (lam_window.__array_interface__["data"][0], lam_window.shape)
        if key in self._lam_cache:                       # the cache holds a reference to the window: its address cannot be recycled
            return self._lam_cache[key][1]
        backend = self.dist.get_backend(self.group)
        t = torch.from_numpy(np.ascontiguousarray(lam_window.T))            # (ng2/P, ng1), C order
        if backend == "nccl":
            t = t.cuda()
        parts = [torch.empty_like(t) for _ in range(self.nranks)]
        self.dist.all_gather(parts, t, group=self.group)
        full = np.asfortranarray(torch.cat(parts, dim=0).cpu().numpy().T)    # (ng1, ng2)
        self._lam_cache[key] = (lam_window, full)
        return full

    def use_nccl_alltoall(self):
        """Exchange (a): torch.distributed all_to_all_single on the library's device buffers."""
        import torch
        dist, group, nranks = self.dist, self.group, self.nranks

        def cb(ctx, send, recv, nbytes, stream):
            try:
                n = nbytes // 8 * nranks
                src = torch.as_tensor(_DevPtr(send, n), device="cuda")
                dst = torch.as_tensor(_DevPtr(recv, n), device="cuda")
                dist.all_to_all_single(dst, src, group=group)
                return 0
            except Exception as e:                                       # never let an exception cross the C boundary
                print("flutas_b200 all-to-all callback failed:", e)
                return 1

        self._cb = _A2A_PROTO(cb)
        _lib.check(_lib.load().flutas_b200_set_alltoall(C.cast(self._cb, C.c_void_p), None))

    def use_p2p(self, arrplan, n_local):
        """Exchange (b): map every peer's exchange buffers (CUDA IPC) so the kernels store over NVLink directly."""
        L = _lib.load()
        nb = L.flutas_b200_p2p_handle_bytes()
        blob = (C.c_char * nb)()
        nl = (C.c_int * 3)(*n_local)
        _lib.check(L.flutas_b200_p2p_export(arrplan.h, nl, blob))
        blobs = [None] * self.nranks
        self.dist.all_gather_object(blobs, bytes(blob), group=self.group)
        allb = (C.c_char * (nb * self.nranks)).from_buffer_copy(b"".join(blobs))
        _lib.check(L.flutas_b200_p2p_attach(arrplan.h, allb))

    def use_halo_exchange(self):
        """z-halo planes of boundp through torch.distributed point-to-point (NCCL send/recv over NVLink)."""
        import torch
        dist, group = self.dist, self.group

        def cb(ctx, send_lo, send_hi, recv_lo, recv_hi, count, lo, hi, stream):
            try:
                t = [torch.as_tensor(_DevPtr(ptr, count), device="cuda") for ptr in (send_lo, send_hi, recv_lo, recv_hi)]
                ops = halo_ops(dist, t[0], t[1], t[2], t[3], lo, hi, group)
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                return 0
            except Exception as e:
                print("flutas_b200 halo callback failed:", e)
                return 1

        self._halo_cb = _HALO_PROTO(cb)
        _lib.check(_lib.load().flutas_b200_set_halo_exchange(C.cast(self._halo_cb, C.c_void_p), None))

    def p2p_errors(self, arrplan):
        return _lib.load().flutas_b200_p2p_errors(arrplan.h)

    def configure(self, pipe_chunks=-1, pipe_xsm_pct=-1, zcopy=-1):
        """schedule knobs of the slab solver (flutas_b200_slab_config); the same values on every rank"""
        _lib.check(_lib.load().flutas_b200_slab_config(int(pipe_chunks), int(pipe_xsm_pct), int(zcopy)))

    def distributed_z(self, on=True):
        """z stage as a distributed tridiagonal solve (no transposes) where the plan allows it; the same value on every rank"""
        _lib.check(_lib.load().flutas_b200_slab_distributed_z(1 if on else 0))

    def autotune(self, solve, candidates=((1, 0, 50), (0, 0, 50), (0, 4, 50)), reps=3):
        """Plan-time measurement (the counterpart of FFTW_MEASURE): runs `solve()` under each (distributed_z, pipe_chunks,
        pipe_xsm_pct) candidate, timed on the device, max over ranks, and keeps the fastest on ALL ranks.
        Returns (choice, {cand: ms})."""
        import torch
        times = {}
        for cand in candidates:
            self.distributed_z(bool(cand[0]))
            self.configure(cand[1], cand[2])
            solve()                                          # first call of a configuration: streams / events are created
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                solve()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            times[cand] = float(t.item())
        best = min(times, key=times.get)
        self.distributed_z(bool(best[0]))
        self.configure(best[1], best[2])
        return best, times

    def solver(self, n_local, arrplan, normfft, lam_window, a, b, c, bcz, c_or_f, p):
        """Collective `solver` on this rank's slab; same argument meaning as the reference's solver_gpu call."""
        lam_full = self.gather_lambda(lam_window)
        api._use_torch_stream()
        nn = (C.c_int * 3)(*n_local)
        _lib.check(_lib.load().flutas_b200_solver_slab(nn, arrplan.h, normfft, api._ptr(lam_full), api._ptr(a), api._ptr(b),
                                                       api._ptr(c), bcz.encode(), "".join(c_or_f).encode(), api._ptr(p)))
        # a cross-GPU barrier that timed out in an EARLIER solve makes this call fail (the library mirrors its error word to
        # pinned host memory behind every barrier); p2p_errors() reports it right away at the price of a stream sync
        return p

    def chkdt(self, *args):
        """the chkdt field reduction + MPI_ALLREDUCE(MAX) of src/chkdt.f90:92,183 (NCCL all-reduce)"""
        import torch
        dti = api.chkdt(*args)
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([dti], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def chkdiv(self, *args):
        """chkdiv + the two MPI_ALLREDUCE of src/chkdiv.f90:64-65"""
        import torch
        tot, mx = api.chkdiv(*args)
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        m = torch.tensor([mx], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item()), float(m.item())
