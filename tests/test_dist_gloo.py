"""World-size-2 (gloo, CPU) check of the multi-rank plumbing bench.py uses: tiles dealt round-robin
to ranks, no data-path collective, max-over-ranks timing, and the union of the ranks' tiles equal to
the single-rank frame bit for bit.  The per-rank compute is the host shim (test scaffolding)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import shim_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = pkg.scenes.example2(ni=70, nj=45)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    dist.barrier()
    out = shim_lib.render_tiles(p, objs, nobj, cam, tile_offset=rank, tile_stride=world)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)       # stand-in for this rank's time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rays = torch.tensor([float(out["counters"]["rays"])], dtype=torch.float64)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    np.save(os.path.join(tmpdir, "part%d.npy" % rank), out["rgb8"])
    np.save(os.path.join(tmpdir, "ids%d.npy" % rank), out["obj_id"])
    if rank == 0:
        np.save(os.path.join(tmpdir, "agg.npy"), np.array([t.item(), rays.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_the_frame(tmp_path, pkg, shim):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    agg = np.load(tmp_path / "agg.npy")
    assert agg[0] == 2.0            # max over ranks
    assert agg[1] == 70 * 45        # every ray traced exactly once across the ranks
    sc = pkg.scenes.example2(ni=70, nj=45)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    full = shim.render_tiles(p, objs, nobj, cam)
    parts = [np.load(tmp_path / ("part%d.npy" % r)) for r in range(2)]
    ids = [np.load(tmp_path / ("ids%d.npy" % r)) for r in range(2)]
    # the ranks touch disjoint pixels (untouched pixels stay 0; every traced pixel has id > 0)
    assert not np.any((ids[0] > 0) & (ids[1] > 0))
    assert np.array_equal(ids[0] + ids[1], full["obj_id"])
    assert np.array_equal(parts[0] + parts[1], full["rgb8"])


# ---- the cross-GPU dynamic tile queue (rtgr_render_frame), modelled on the host --------------------------
def _frame_worker(rank, world, port, tmpdir, shm_name, ni, nj):
    import ctypes
    from multiprocessing import shared_memory
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import shim_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        base = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        img = np.ndarray((nj, ni, 3), dtype=np.uint8, buffer=shm.buf, offset=256)
        sc = pkg.scenes.example2(ni=ni, nj=nj)
        p, objs, nobj, cam = pkg.scenes.to_abi(sc)
        shares = []
        for frame in range(2):                        # two frames: the heads alternate (byte 0 / byte 128)
            dist.barrier()                            # the caller's barrier between frames
            head = base + 128 * (frame & 1)
            if rank == 0:                             # the owner re-arms the head of the NEXT frame
                ctypes.c_uint64.from_address(base + 128 * ((frame + 1) & 1)).value = 0
            cnt = shim_lib.render_frame(p, objs, nobj, cam, head, img)
            shares.append(cnt["rays"])
        rays = torch.tensor(shares, dtype=torch.float64)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        np.save(os.path.join(tmpdir, "share%d.npy" % rank), np.array(shares))
        if rank == 0:
            np.save(os.path.join(tmpdir, "rays.npy"), rays.numpy())
        dist.barrier()
        del img
    finally:
        shm.close()
    dist.destroy_process_group()


def test_two_ranks_share_one_tile_queue(tmp_path, pkg, shim):
    """Two processes draw 8x4-pixel patches from ONE queue head and store into ONE image (shared memory
    stands in for the owner GPU's peer memory): every ray is traced exactly once whoever gets it, and the
    image equals the single-process render bit for bit."""
    from multiprocessing import shared_memory
    ni, nj = 70, 45
    shm = shared_memory.SharedMemory(create=True, size=256 + ni * nj * 3)
    try:
        shm.buf[:] = bytes(len(shm.buf))
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        mp.spawn(_frame_worker, args=(2, port, str(tmp_path), shm.name, ni, nj), nprocs=2, join=True)
        img = np.ndarray((nj, ni, 3), dtype=np.uint8, buffer=shm.buf, offset=256).copy()
        heads = np.ndarray((2,), dtype=np.uint64, buffer=shm.buf, strides=(128,)).copy()
    finally:
        shm.close()
        shm.unlink()
    rays = np.load(tmp_path / "rays.npy")
    assert list(rays) == [ni * nj, ni * nj]          # each frame: every ray exactly once across the ranks
    shares = [np.load(tmp_path / ("share%d.npy" % r)) for r in range(2)]
    assert all(int(sh.sum()) > 0 for sh in shares)
    sc = pkg.scenes.example2(ni=ni, nj=nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    full = shim.render_tiles(p, objs, nobj, cam)
    assert np.array_equal(img, full["rgb8"])
    ntiles = ((ni + 31) // 32) * ((nj + 31) // 32)
    assert heads[1] >= ntiles * 1024                 # frame 1 drained head[1]; head[0] was re-armed during it
    assert heads[0] == 0


# ---- trace_rays on ONE canvas by two processes (rtgr_trace_canvas_frame), modelled on the host ------------
def _canvas_frame_worker(rank, world, port, tmpdir, shm_name, ni, nj):
    import ctypes
    from multiprocessing import shared_memory
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import shim_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        base = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        canvas = np.ndarray((nj, ni, 11), dtype=np.float64, buffer=shm.buf, offset=256)
        sc = pkg.scenes.example2(ni=ni, nj=nj)
        p, objs, nobj, _cam = pkg.scenes.to_abi(sc)
        dist.barrier()
        cnt = shim_lib.trace_canvas_frame(p, objs, nobj, canvas, base)       # queue head at byte 0 of the segment
        rays = torch.tensor([float(cnt["rays"])], dtype=torch.float64)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        np.save(os.path.join(tmpdir, "cshare%d.npy" % rank), np.array([cnt["rays"]]))
        if rank == 0:
            np.save(os.path.join(tmpdir, "crays.npy"), rays.numpy())
        dist.barrier()
        del canvas
    finally:
        shm.close()
    dist.destroy_process_group()


def test_two_ranks_trace_one_canvas(tmp_path, pkg, shim):
    """The protocol of rtgr_trace_canvas_frame on the CPU: two processes map ONE Pixel canvas (POSIX shared memory),
    derive the same queue order from its contents, draw their rays from one head and write rgb in place -- the
    canvas ends up complete, with no gather step, and equal to the single-process trace bit for bit."""
    from multiprocessing import shared_memory
    ni, nj = 70, 45
    sc = pkg.scenes.example2(ni=ni, nj=nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    canvas0 = shim.make_canvas(p, cam).reshape(nj, ni, 11)
    shm = shared_memory.SharedMemory(create=True, size=256 + ni * nj * 88)
    try:
        shm.buf[:256] = bytes(256)
        shared = np.ndarray((nj, ni, 11), dtype=np.float64, buffer=shm.buf, offset=256)
        shared[...] = canvas0
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        mp.spawn(_canvas_frame_worker, args=(2, port, str(tmp_path), shm.name, ni, nj), nprocs=2, join=True)
        result = shared.copy()
        del shared
    finally:
        shm.close()
        shm.unlink()
    assert list(np.load(tmp_path / "crays.npy")) == [ni * nj]
    shares = [int(np.load(tmp_path / ("cshare%d.npy" % r))[0]) for r in range(2)]
    assert all(sh > 0 for sh in shares) and sum(shares) == ni * nj
    single = shim.trace_pixels(p, objs, nobj, canvas0.reshape(-1, 11))
    assert np.array_equal(result[:, :, :8], canvas0[:, :, :8])              # pos / normal untouched
    assert np.array_equal(result.reshape(-1, 11)[:, 8:], single["rgb"])
    assert np.any(result[:, :, 8:] != 0.0)
