"""Restart-file format next to the path: load(io,filename,n,fld), src/load.f90:21-89 --
headerless raw FP64 in global column-major order, each rank's block through a subarray view
(src/2decomp/io_write_var.f90:28-60).  The numpy statement of that format is the oracle here."""
import numpy as np
import pytest

from flutas_b200 import api, lib


def _global(ng, seed=0):
    return np.asfortranarray(np.random.default_rng(seed).uniform(-1, 1, ng))


def test_write_blocks_in_any_order_then_read_back(tmp_path):
    ng = (12, 10, 8)
    g = _global(ng)
    f = tmp_path / "fldp.bin"
    # z-slabs (the path's decomposition), written out of order; then an x-pencil style 2x2 decomposition on another file
    for r in (2, 0, 3, 1):
        blk = np.asfortranarray(g[:, :, 2 * r:2 * r + 2])
        api.load("w", f, blk.shape, blk, ng=ng, start=(0, 0, 2 * r))
    raw = np.fromfile(f, dtype=np.float64)
    assert raw.size == np.prod(ng) and np.array_equal(raw.reshape(ng, order="F"), g)      # the reference's on-disk layout
    f2 = tmp_path / "fldu.bin"
    for (j0, k0) in ((5, 4), (0, 0), (0, 4), (5, 0)):
        blk = np.asfortranarray(g[:, j0:j0 + 5, k0:k0 + 4])
        api.load("w", f2, blk.shape, blk, ng=ng, start=(0, j0, k0))
    assert np.array_equal(np.fromfile(f2, dtype=np.float64).reshape(ng, order="F"), g)
    # read a block into the interior of a halo'd array: halos untouched
    p = np.full((12 + 2, 5 + 2, 4 + 2), 7.0, order="F")
    api.load("r", f2, (12, 5, 4), p, ng=ng, start=(0, 5, 4), nh=1)
    assert np.array_equal(p[1:-1, 1:-1, 1:-1], g[:, 5:10, 4:8])
    p[1:-1, 1:-1, 1:-1] = 7.0
    assert np.all(p == 7.0)


def test_read_errors_like_the_reference(tmp_path):
    ng = (4, 4, 4)
    buf = np.zeros(ng, order="F")
    with pytest.raises(lib.FlutasB200Error, match="does not exist"):
        api.load("r", tmp_path / "missing.bin", ng, buf)
    bad = tmp_path / "bad.bin"
    np.zeros(63).tofile(bad)
    with pytest.raises(lib.FlutasB200Error, match="incorrect size"):
        api.load("r", bad, ng, buf)
    with pytest.raises(lib.FlutasB200Error, match="outside the global grid"):
        api.load("r", bad, (4, 4, 4), buf, ng=ng, start=(0, 0, 1))


@pytest.mark.gpu
def test_device_resident_field_round_trip(tmp_path):
    import torch
    api.init(0)
    ng = (40, 24, 16)
    g = _global(ng, 3)
    f = tmp_path / "fldp.bin"
    for r in range(2):                                                    # two z-slabs from device-resident halo'd p
        p = np.zeros((ng[0] + 2, ng[1] + 2, 8 + 2), order="F")
        p[1:-1, 1:-1, 1:-1] = g[:, :, 8 * r:8 * r + 8]
        pd = api.device_field(p)
        api.load("w", f, (ng[0], ng[1], 8), pd, ng=ng, start=(0, 0, 8 * r), nh=1)
    assert np.array_equal(np.fromfile(f, dtype=np.float64).reshape(ng, order="F"), g)
    pd = torch.full((8 + 2, ng[1] + 2, ng[0] + 2), -3.0, dtype=torch.float64, device="cuda")
    api.load("r", f, (ng[0], ng[1], 8), pd, ng=ng, start=(0, 0, 8), nh=1)
    back = api.host_field(pd, (ng[0] + 2, ng[1] + 2, 10))
    assert np.array_equal(back[1:-1, 1:-1, 1:-1], g[:, :, 8:16]) and back[0, 0, 0] == -3.0


def _load_worker(rank, world, port, path, ng):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = _global(ng, 9)
        n3l = ng[2] // world
        # every rank writes its z-slab of the checkpoint concurrently (no ordering, no truncation race) ...
        blk = np.asfortranarray(g[:, :, rank * n3l:(rank + 1) * n3l])
        api.load("w", path, blk.shape, blk, ng=ng, start=(0, 0, rank * n3l))
        dist.barrier()
        # ... and reads back ANOTHER rank's slab into a halo'd array: the file is decomposition independent
        other = (rank + 1) % world
        p = np.zeros((ng[0] + 2, ng[1] + 2, n3l + 2), order="F")
        api.load("r", path, (ng[0], ng[1], n3l), p, ng=ng, start=(0, 0, other * n3l), nh=1)
        assert np.array_equal(p[1:-1, 1:-1, 1:-1], g[:, :, other * n3l:(other + 1) * n3l])
        if rank == 0:
            assert np.array_equal(np.fromfile(path, dtype=np.float64).reshape(ng, order="F"), g)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_written_by_concurrent_ranks(tmp_path, world):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_load_worker, args=(world, port, str(tmp_path / "fldp.bin"), (16, 12, 8)), nprocs=world, join=True)
