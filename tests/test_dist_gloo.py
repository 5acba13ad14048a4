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
