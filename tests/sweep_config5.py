#!/usr/bin/env python3
"""BASELINE configs[4]: Kerr-Schild a = 0.9 at 7680x4320, tolerance sweep 1e-6 .. 1e-10, the frame shared by all
ranks through the dynamic tile queue of rtgr_frame_* (run under torchrun, one rank per GPU; also works with one
process).  One JSON line per tolerance: frame time (max over ranks, barrier on both sides, device make_canvas, RGB8
image assembled in rank 0's memory), rays/s, RHS evaluations/s, per-rank kernel times, sha256 of the image.
Developer measurement (tests/gpu_session.sh); the parity of this configuration is tests/test_gpu_parity.py's."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


pkg = entry.load_package()
ctx = pkg.Context([local])
sc0 = pkg.scenes.config5()
hb = torch.zeros(pkg._abi.RTGR_IPC_HANDLE_BYTES, dtype=torch.uint8, device="cuda")
frame = None
if rank == 0:
    frame = pkg.Frame(ctx, sc0.ni, sc0.nj)
    hb.copy_(torch.tensor(list(frame.handle), dtype=torch.uint8))
if world > 1:
    dist.broadcast(hb, 0)
if rank != 0:
    frame = pkg.Frame(ctx, sc0.ni, sc0.nj, handle=bytes(hb.cpu().tolist()))
frame.set_participants(world)
steps = 3
for tol in (1e-6, 1e-7, 1e-8, 1e-9, 1e-10):
    sc = pkg.scenes.config5(tol=tol)
    barrier()
    frame.render(sc)            # warm-up
    t_all, k_ms, st = 0.0, 0.0, None
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        st = frame.render(sc)
        barrier()
        t_all += time.perf_counter() - t0
        k_ms += st["kernel_ms"]
    vals = torch.tensor([t_all / steps, k_ms / steps, float(st["rays"]), float(st["rhs_evals"]),
                         float(st["steps_accepted"] + st["steps_rejected"]), float(st["steps_rejected"])], dtype=torch.float64, device="cuda")
    gl = [torch.zeros_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(gl, vals)
    else:
        gl = [vals]
    if rank == 0:
        sec = max(float(g[0]) for g in gl)
        rays = sum(float(g[2]) for g in gl)
        rhs = sum(float(g[3]) for g in gl)
        att = sum(float(g[4]) for g in gl)
        img = frame.read()
        print(json.dumps({"workload": sc.name, "ni": sc.ni, "nj": sc.nj, "tol": tol, "n_gpus": world, "ms_per_frame": 1e3 * sec,
                          "rays_per_s": rays / sec, "rhs_evals_per_s": rhs / sec, "rhs_per_ray": rhs / rays,
                          "steps_rejected": sum(float(g[5]) for g in gl), "model_tflops": (383 * rhs + 516 * att) / sec / 1e12,
                          "kernel_ms_per_rank": [round(float(g[1]), 2) for g in gl],
                          "rgb8_sha256_16": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()[:16]}), flush=True)
    barrier()
frame.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
