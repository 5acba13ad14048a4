#!/usr/bin/env python3
"""Benchmark of the geodesic ray-tracing hot path (BASELINE.json metric: Kerr-Schild rays/s and
RHS evaluations/s on 1/2/4/8 B200, fraction of the FP64 peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full render of the workload frame (default: BASELINE.json configs[3], Kerr-Schild
a=0.99, 3840x2160, wide field of view).  With N > 1 (launched under torchrun, one rank per GPU) the
ranks share ONE frame: a dynamic tile queue and the RGB8 image live in rank 0's HBM and every rank's
kernel draws 8x4-pixel patches from that queue and stores its pixels into that image over NVLink peer
memory (rtgr_frame_*, the handle travels by a broadcast; `--queue static` deals fixed cost-balanced tile
sets to the ranks instead).  There is no data-path collective, and `value` = rays of the whole frame /
max-over-ranks time, i.e. strong scaling of one frame.

  value     kernel path: canvas generated on the device, results left in HBM (rtgr_render_resident)
  e2e       the drop-in call for the reference's trace_rays, rtgr_trace_canvas, on the caller's
            Array{Pixel} in page-locked HOST memory: the kernel reads the rays from and writes rgb into
            host memory in place over PCIe (zero copy), all inside the timed region
  roofline  FP64 CUDA-core roofline: (383*N_rhs + 516*N_attempts) / kernel time from CUDA events,
            against this repo's own register-resident DFMA microbenchmark on the same GPU
  cpu_baseline / --impl reference
            the C++ restatement of the reference path (oracle/, as-written operation order) with
            OpenMP over all host cores, on a bounded sample of the same workload.  The reference
            itself is Julia and cannot run in this image.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402

W_RHS = 383     # algorithmic flops per Kerr-Schild RHS evaluation
W_STEP = 516    # algorithmic flops per step attempt outside the RHS

def load_ncu_capture():
    """The newest committed `ncu --set full` summary of trace_kernel on the default workload
    (profiles/*_trace_kernel_4k_ncu.json, written by tools/ncu_summary.py from the raw ncu CSV beside it).
    Static facts quoted beside the live numbers; the line says whether the capture is of the kernel build just timed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_trace_kernel_4k_ncu.json")))
    if not files:
        return None
    cap = json.load(open(files[-1]))
    cap["file"] = os.path.relpath(files[-1], ROOT)
    return cap


def lattice_sample(scene, target_rays):
    """Pixels of a regular sub-lattice of the frame (same camera, same rays as the full frame)."""
    n = scene.ni * scene.nj
    stride = max(1, int(round((n / max(1, target_rays)) ** 0.5)))
    ii = np.arange(stride // 2, scene.ni, stride)
    jj = np.arange(stride // 2, scene.nj, stride)
    J, I = np.meshgrid(jj, ii, indexing="ij")
    return (I + J * scene.ni).ravel(), stride


def cpu_reference_run(pkg, scene, target_rays, steps=1, warmup=0, budget_s=None):
    """Time the oracle (C++ restatement of the reference path, all host cores) on a lattice sample."""
    import oracle_lib
    p, objs, nobj, cam = pkg.scenes.to_abi(scene)
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or oracle_lib.num_threads()
    px_all = None
    if budget_s is not None:
        # calibrate: a small probe tells how many rays fit the per-step budget
        idx, _ = lattice_sample(scene, 64 * cores)
        px_all = oracle_lib.make_canvas(p, cam)
        t0 = time.perf_counter()
        oracle_lib.trace_pixels(p, objs, nobj, px_all[idx], nthreads=cores)
        rate = len(idx) / (time.perf_counter() - t0)
        target_rays = int(max(64 * cores, min(scene.ni * scene.nj, rate * budget_s)))
    if px_all is None:
        px_all = oracle_lib.make_canvas(p, cam)
    idx, stride = lattice_sample(scene, target_rays)
    px = np.ascontiguousarray(px_all[idx])
    times, stats = [], None
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        r = oracle_lib.trace_pixels(p, objs, nobj, px, nthreads=cores)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
            stats = r["stats"]
    sec = float(np.mean(times))
    return dict(rays=len(idx), stride=stride, sec_per_step=sec, rays_per_s=len(idx) / sec,
                rhs_per_s=stats["rhs_evals"] / sec, cores=cores, stats=stats)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def run_reference(args, pkg, scene, rank):
    """--impl reference: the reference's CPU path (C++ restatement; Julia is not installable here) on
    all host cores.  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    budget = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    r = cpu_reference_run(pkg, scene, 0, steps=args.steps, warmup=args.warmup, budget_s=budget)
    sample = ("every %d-th pixel in i and j of the %dx%d frame = %d rays per step"
              % (r["stride"], scene.ni, scene.nj, r["rays"]))
    line = {
        "impl": "reference", "metric": "rays_per_s", "value": r["rays_per_s"], "unit": "rays/s",
        "rhs_evals_per_s": r["rhs_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["sec_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        # the GPU arm's config, literally (the driver compares the two arms' `config`)
        "config": workload_config(scene, L2_NOTE, "shared" if args.gpus > 1 else "static"),
        "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": "port",
                         "sample": sample, "cpu": cpu_model(),
                         "note": "C++ restatement of the reference path (Julia unavailable in image), OpenMP static chunks"},
        "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


SHARDING = {
    "static": "32x32-pixel tiles, sorted by estimated cost and dealt round-robin to the ranks (rtgr_render_tiles), no collective",
    "shared": "ONE dynamic tile queue + RGB8 image in rank 0's HBM; every rank's kernel draws 8x4-pixel patches from it with "
              "system-scope atomics and stores its pixels into it over NVLink peer memory (rtgr_frame_*, CUDA IPC); no collective",
}


L2_NOTE = "256 MB device memset between steps (L2 is 126 MB); inputs are <1 KB of scene constants"


def workload_config(scene, l2_note, queue="static"):
    return {"workload": scene.name, "note": scene.note, "metric": "kerr_schild" if scene.metric == 1 else "minkowski",
            "M": scene.M, "a": scene.a, "r_formula": "as_written(src:284)" if scene.r_formula == 0 else "corrected",
            "ni": scene.ni, "nj": scene.nj, "reltol": scene.tol, "abstol": scene.tol,
            "sharding": SHARDING[queue], "l2": l2_note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=["example1", "example2", "config3", "config4", "config5"])
    ap.add_argument("--ni", type=int, default=0)
    ap.add_argument("--nj", type=int, default=0)
    ap.add_argument("--tol", type=float, default=0.0, help="reltol = abstol override (config5 tolerance sweep)")
    ap.add_argument("--queue", default="auto", choices=["auto", "static", "shared"],
                    help="how the ranks share the frame: static = cost-balanced tile sets per rank (rtgr_render_tiles); "
                         "shared = ONE dynamic tile queue + image in rank 0's GPU memory that every rank draws from and "
                         "stores into over NVLink peer memory (rtgr_frame_*, CUDA IPC); auto = shared when there is more "
                         "than one rank (static if the frame cannot be shared on this box), static for one rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true", help="profiling runs: no 256 MB memset before a step (its dirty lines would be written back during the captured kernel)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    pkg = entry.load_package()
    scene = pkg.scenes.BY_NAME[args.workload]()
    if args.ni and args.nj:
        scene = scene.with_size(args.ni, args.nj)
    if args.tol > 0:
        from dataclasses import replace
        scene = replace(scene, tol=args.tol, name=scene.name.split("_tol")[0] + "_tol%g" % args.tol)

    if args.impl == "reference":
        run_reference(args, pkg, scene, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU restatement of the reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    ctx = pkg.Context([local_rank])
    n_rays = scene.ni * scene.nj
    class _Flush:      # L2 flush between timed iterations: a write of a buffer larger than the 126 MB L2
        def __init__(self, on):
            self.buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if on else None

        def zero_(self):
            if self.buf is not None:
                self.buf.zero_()
    flush = _Flush(not args.no_l2_flush)

    # FP64 roofline denominator: register-resident DFMA chains on this GPU (best of several launches)
    peak_tf = max(ctx.fp64_peak(0)[0] for _ in range(3))

    # ---------------- kernel path: inputs resident, nothing copied ----------------
    import hashlib

    def sha(img):
        return hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()[:16]

    frame, queue_note = None, None
    if args.queue == "auto":
        args.queue = "shared" if world > 1 else "static"
    if args.queue == "shared":
        # rank 0 owns the frame (queue heads + RGB8 image in its HBM); the others map it through CUDA IPC
        hb = torch.zeros(pkg._abi.RTGR_IPC_HANDLE_BYTES, dtype=torch.uint8, device="cuda")
        ok, why = 1.0, ""
        if rank == 0:
            try:
                frame = pkg.Frame(ctx, scene.ni, scene.nj)
                hb.copy_(torch.tensor(list(frame.handle), dtype=torch.uint8))
            except Exception as e:      # noqa: BLE001 -- reported, then every rank falls back together
                ok, why = 0.0, str(e)
        if world > 1:
            dist.broadcast(hb, 0)
        if rank != 0:
            try:
                frame = pkg.Frame(ctx, scene.ni, scene.nj, handle=bytes(hb.cpu().tolist()))
            except Exception as e:      # noqa: BLE001
                ok, why = 0.0, str(e)
        if frame is not None:
            frame.set_participants(world)
        if allreduce(ok, dist.ReduceOp.MIN if world > 1 else None) < 1.0:
            # e.g. a sandbox without CUDA IPC: all ranks use the static deal (still the GPU path)
            if frame is not None:
                frame.close()
            frame, args.queue = None, "static"
            queue_note = "shared frame unavailable on this box (%s): static deal" % (why or "another rank failed to map it")
            if rank == 0:
                print("bench.py: " + queue_note, file=sys.stderr)

    # The picture one GPU renders alone (rank 0, once, untimed): what every multi-GPU result must equal bit for bit.
    single_sha = None
    if rank == 0:
        single_sha = sha(ctx.render(scene, want=("rgb8",))["rgb8"])

    def step_kernel():
        flush.zero_()
        torch.cuda.synchronize()
        if frame is not None:
            barrier()       # the frame protocol's barrier between consecutive frames (inside the timed region)
            return frame.render(scene)
        return ctx.render_resident(scene, tile_offset=rank, tile_stride=world)

    frame_matches_single = None
    for w in range(args.warmup):
        step_kernel()
        if w == 0 and frame is not None:
            # the first shared frame, checked against the single-GPU picture before anything is timed
            barrier()
            if rank == 0:
                frame_matches_single = (sha(frame.read()) == single_sha)
                if not frame_matches_single:
                    raise SystemExit("bench.py: the frame rendered by %d GPUs differs from the single-GPU frame" % world)
            barrier()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    kernel_ms, stats_sum = 0.0, None
    for _ in range(args.steps):
        st = step_kernel()
        kernel_ms += st["kernel_ms"]
        stats_sum = st if stats_sum is None else {k: stats_sum[k] + st[k] for k in st}
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.result()
    # L2 flush time is inside the bracket; subtract nothing -- it is < 0.1 ms per step
    wall_max = allreduce(wall, dist.ReduceOp.MAX if world > 1 else None)
    kernel_ms_max = allreduce(kernel_ms, dist.ReduceOp.MAX if world > 1 else None)
    rays_total = allreduce(float(stats_sum["rays"]), dist.ReduceOp.SUM if world > 1 else None)
    rhs_total = allreduce(float(stats_sum["rhs_evals"]), dist.ReduceOp.SUM if world > 1 else None)
    acc_total = allreduce(float(stats_sum["steps_accepted"]), dist.ReduceOp.SUM if world > 1 else None)
    rej_total = allreduce(float(stats_sum["steps_rejected"]), dist.ReduceOp.SUM if world > 1 else None)
    attempts_total = acc_total + rej_total
    per_rank_kernel_ms = [kernel_ms / args.steps]
    per_rank_rays = [stats_sum["rays"] / args.steps]
    if world > 1:
        gl = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gl, torch.tensor([kernel_ms / args.steps, stats_sum["rays"] / args.steps], dtype=torch.float64, device="cuda"))
        per_rank_kernel_ms = [float(g[0].item()) for g in gl]
        per_rank_rays = [float(g[1].item()) for g in gl]
    # sha256 of the RGB8 frame the timed path produced: the same at every N (N = 1: this rank's resident image)
    frame_sha = None
    if frame is not None:
        barrier()
        if rank == 0:
            frame_sha = sha(frame.read())
        barrier()
    elif world == 1:
        frame_sha = sha(ctx.render(scene, want=("rgb8",))["rgb8"])
    value = rays_total / wall_max
    # roofline of the trace kernel on this rank (rank 0 reports its own kernel)
    my_attempts = stats_sum["steps_accepted"] + stats_sum["steps_rejected"]
    my_flops = W_RHS * stats_sum["rhs_evals"] + W_STEP * my_attempts
    achieved_tf = my_flops / (kernel_ms * 1e-3) / 1e12

    # ---------------- end to end: the trace_rays drop-in with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        p, objs, nobj, cam = pkg.scenes.to_abi(scene)
        api_note = ("the kernel reads pos/normal (64 B/ray) from and writes rgb (24 B/ray) into HOST memory in place over "
                    "PCIe while it computes (no copy brackets it); timed from call to return, barrier on both sides; "
                    "bytes are whole-job totals per step")

        def canvas_rgb8_sha(canvas):
            q = np.rint(255.0 * np.clip(canvas[:, :, 8:], 0.0, 1.0)).astype(np.uint8)
            return sha(q)

        if world == 1:
            # The caller's input: the whole canvas (Array{Pixel{Float64},2}, 88 B per pixel) in page-locked host
            # memory, built once outside the timed region; rtgr_trace_canvas traces it in place.
            buf = pkg.PinnedArray((scene.nj, scene.ni, 11))
            buf.array[...] = ctx.make_canvas(p, cam).reshape(scene.nj, scene.ni, 11)
            canvas = buf.array

            def call():
                return ctx.trace_canvas(p, objs, nobj, canvas)["stats"]
            api = "rtgr_trace_canvas (drop-in for trace_rays, src:483) on a page-locked host Array{Pixel}: " + api_note
        else:
            # ONE canvas for all ranks: POSIX shared memory that every rank maps and page-locks.  All ranks call
            # rtgr_trace_canvas_frame: rays from the frame's shared queue (dynamic balance), rgb written into the one
            # array in place -- the assembled host canvas is there when the call returns, no gather step.
            name_t = torch.zeros(64, dtype=torch.uint8, device="cuda")
            shared = None
            if rank == 0 and frame is not None and not os.environ.get("RTGR_BENCH_PRIVATE_CANVASES"):   # (the variable: to test the fallback)
                try:
                    shared = pkg.SharedCanvas(scene.nj, scene.ni)
                    shared.array[...] = ctx.make_canvas(p, cam).reshape(scene.nj, scene.ni, 11)
                    nb = shared.name.encode()
                    name_t[:len(nb)] = torch.tensor(list(nb), dtype=torch.uint8)
                except Exception as e:      # noqa: BLE001 -- e.g. /dev/shm smaller than the canvas: reported below
                    print("bench.py: " + str(e), file=sys.stderr)
            dist.broadcast(name_t, 0)
            shm_name = bytes(name_t.cpu().tolist()).rstrip(b"\0").decode()
            host_images = 1
            if shm_name:
                if rank != 0:
                    shared = pkg.SharedCanvas(scene.nj, scene.ni, name=shm_name)
                canvas = shared.array
                barrier()

                def call():
                    return frame.trace_canvas(p, objs, nobj, canvas)
                api = ("rtgr_trace_canvas_frame (trace_rays, src:483, on ONE page-locked host Array{Pixel} in POSIX shared memory "
                       "mapped by all %d ranks; rays drawn from the shared queue in rank 0's HBM): " % world) + api_note
            else:
                # No shared frame (CUDA IPC unavailable) or no room for the canvas in /dev/shm on this box: every rank traces
                # its cost-balanced tiles of a PRIVATE page-locked canvas -- N partial host images, said so in the line.
                host_images = world
                buf = pkg.PinnedArray((scene.nj, scene.ni, 11))
                buf.array[...] = ctx.make_canvas(p, cam).reshape(scene.nj, scene.ni, 11)
                canvas = buf.array

                def call():
                    return ctx.trace_canvas(p, objs, nobj, canvas, tile_offset=rank, tile_stride=world)["stats"]
                api = ("rtgr_trace_canvas with tile_offset/tile_stride on a PRIVATE page-locked canvas per rank (one shared host "
                       "canvas was not possible on this box: %s): " % ("no shared frame" if frame is None else "no POSIX shared memory for it")) + api_note

        one_canvas = (world == 1) or (host_images == 1)

        def reset():
            if rank == 0 or not one_canvas:
                canvas[:, :, 8:] = 0.0       # results of the previous step cannot be reused
            flush.zero_()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            reset()
            barrier()
            call()
            barrier()
        e2e_call_s, e2e_rays, e2e_kernel_ms, e2e_own_call_s = 0.0, 0.0, 0.0, 0.0
        for _ in range(args.steps):
            reset()
            barrier()
            t0 = time.perf_counter()
            st = call()
            t1 = time.perf_counter()
            barrier()
            e2e_call_s += time.perf_counter() - t0
            e2e_own_call_s += t1 - t0                 # this rank's call alone (the rest is waiting for the others)
            e2e_kernel_ms += st["kernel_ms"]          # CUDA-event time of this rank's kernel inside the call
            e2e_rays = st["rays"]
        e2e_wall = allreduce(e2e_call_s, dist.ReduceOp.MAX if world > 1 else None)
        e2e_rays_total = allreduce(float(e2e_rays), dist.ReduceOp.SUM if world > 1 else None)
        e2e_sha = canvas_rgb8_sha(canvas) if (rank == 0 and one_canvas) else None
        e2e = {"value": n_rays * args.steps / e2e_wall, "unit": "rays/s", "ms_per_step": 1e3 * e2e_wall / args.steps,
               "h2d_bytes_per_step": int(e2e_rays_total * 64), "d2h_bytes_per_step": int(e2e_rays_total * 24),
               "kernel_ms_per_rank": None, "call_ms_per_rank": None,
               "api": api, "host_images": 1 if one_canvas else world, "rgb8_sha256_16": e2e_sha,
               "matches_frame": (e2e_sha == frame_sha) if (rank == 0 and frame_sha is not None and e2e_sha is not None) else None}
        # where an e2e step goes: per rank the kernel (device clock) and the whole call (host clock), averaged over the steps
        pr = torch.tensor([e2e_kernel_ms / args.steps, 1e3 * e2e_own_call_s / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            gl2 = [torch.zeros_like(pr) for _ in range(world)]
            dist.all_gather(gl2, pr)
        else:
            gl2 = [pr]
        e2e["kernel_ms_per_rank"] = [round(float(g[0].item()), 3) for g in gl2]
        e2e["call_ms_per_rank"] = [round(float(g[1].item()), 3) for g in gl2]
        if rank == 0 and frame_sha is not None and e2e_sha is not None and e2e_sha != frame_sha:
            raise SystemExit("bench.py: the e2e canvas differs from the frame of the kernel-only path")
        if world > 1 and frame is not None:
            # second form of "one image on the host": every rank renders its share of the frame (device make_canvas)
            # into the image in rank 0's HBM, then rank 0 copies the assembled RGB8 image to the host -- all timed
            g_s = 0.0
            for it in range(args.warmup + args.steps):
                flush.zero_()
                barrier()
                t0 = time.perf_counter()
                frame.render(scene)
                barrier()
                img = frame.read() if rank == 0 else None
                barrier()
                if it >= args.warmup:
                    g_s += time.perf_counter() - t0
            g_wall = allreduce(g_s, dist.ReduceOp.MAX)
            e2e["render_frame_gather"] = {
                "value": n_rays * args.steps / g_wall, "unit": "rays/s", "ms_per_step": 1e3 * g_wall / args.steps,
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(n_rays * 3),
                "api": "rtgr_render_frame on every rank + rtgr_frame_read on rank 0: the RGB8 image assembled in rank 0's "
                       "HBM by the kernels themselves, then ONE device-to-host copy, inside the timed region",
                "matches_frame": (sha(img) == frame_sha) if rank == 0 else None}
            barrier()
        if world > 1:
            if shared is not None:
                shared.close()
            elif not one_canvas:
                buf.free()
        else:
            # the same call on PAGEABLE host memory (staged: whole-canvas H2D, trace, D2H)
            pageable = np.array(canvas, copy=True)
            ctx.trace_canvas(p, objs, nobj, pageable)
            t0 = time.perf_counter()
            for _ in range(max(1, args.steps // 2)):
                ctx.trace_canvas(p, objs, nobj, pageable)
            pg = (time.perf_counter() - t0) / max(1, args.steps // 2)
            e2e["pageable_host_buffer"] = {"value": n_rays / pg, "unit": "rays/s", "ms_per_step": 1e3 * pg,
                                           "h2d_bytes_per_step": int(n_rays * 88), "d2h_bytes_per_step": int(n_rays * 88)}
            del pageable
            # production path with the canvas generated on the device, RGB8 image copied back to the host
            img = np.zeros((scene.nj, scene.ni, 3), dtype=np.uint8)
            out = {"rgb8": img}
            for _ in range(max(1, args.warmup - 1)):
                ctx.render(scene, want=("rgb8",), out=out)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.render(scene, want=("rgb8",), out=out)
            r_wall = time.perf_counter() - t0
            e2e["render_rgb8"] = {"value": n_rays * args.steps / r_wall, "unit": "rays/s",
                                  "api": "rtgr_render: device-side make_canvas, RGB8 frame copied to the host",
                                  "d2h_bytes_per_step": int(n_rays * 3)}
            buf.free()
    if frame is not None:
        barrier()
        frame.close()

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(pkg, scene, 0, steps=1, warmup=0, budget_s=15.0)
        cpu = {"value": r["rays_per_s"], "unit": "rays/s", "rhs_evals_per_s": r["rhs_per_s"], "cores": r["cores"],
               "kind": "port", "cpu": cpu_model(),
               "sample": "every %d-th pixel in i and j of the %dx%d frame = %d rays, %.1f s (a SAMPLED sub-lattice, "
                         "extrapolated as an intensive rate)" % (r["stride"], scene.ni, scene.nj, r["rays"], r["sec_per_step"]),
               "note": "C++ restatement of the reference path in its as-written operation order "
                       "(Julia unavailable in image), OpenMP schedule(static) over the pixel index"}

    # ---------------- the reference's own two scenes beside the headline workload (rank 0, N = 1) ----------------
    # BASELINE.json configs[0] and [1]: example1 / example2 at the reference's 200x200, the GPU against the CPU
    # restatement on the WHOLE frame (40 000 rays, no sampling) -- a few seconds of CPU time.
    small = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "config4" and not (args.ni or args.nj):
        import oracle_lib
        small = {}
        for nm in ("example1", "example2"):
            sc2 = pkg.scenes.BY_NAME[nm]()
            p2, objs2, nobj2, cam2 = pkg.scenes.to_abi(sc2)
            ctx.render_resident(sc2)
            ks = [ctx.render_resident(sc2)["kernel_ms"] for _ in range(5)]
            buf2 = pkg.PinnedArray((sc2.nj, sc2.ni, 11))
            buf2.array[...] = ctx.make_canvas(p2, cam2).reshape(sc2.nj, sc2.ni, 11)
            ctx.trace_canvas(p2, objs2, nobj2, buf2.array)
            t0 = time.perf_counter()
            for _ in range(5):
                ctx.trace_canvas(p2, objs2, nobj2, buf2.array)
            e2e_ms = 1e3 * (time.perf_counter() - t0) / 5
            buf2.free()
            cores = len(os.sched_getaffinity(0))
            px2 = oracle_lib.make_canvas(p2, cam2)
            t0 = time.perf_counter()
            oracle_lib.trace_pixels(p2, objs2, nobj2, px2, nthreads=cores)
            cpu_s = time.perf_counter() - t0
            n2 = sc2.ni * sc2.nj
            small[nm] = {"ni": sc2.ni, "nj": sc2.nj, "kernel_ms": float(np.median(ks)), "rays_per_s": n2 / (float(np.median(ks)) * 1e-3),
                         "e2e_ms": e2e_ms, "e2e_rays_per_s": n2 / (e2e_ms * 1e-3),
                         "cpu_port_rays_per_s": n2 / cpu_s, "cpu_cores": cores, "cpu_sample": "the whole frame"}

    if rank == 0:
        # What the kernel EXECUTES, from the committed ncu capture of this workload (per-attempt instruction counts are
        # a property of the kernel binary; the attempts and the time are this run's): the hardware-side reading next to
        # the model-flop roofline that the benchmark contract defines.
        cap = load_ncu_capture() if (scene.name == "ks_a0.99_4k_wide" and scene.ni == 3840) else None
        executed = None
        if cap is not None:
            same_build = (cap.get("kernel_source_sha16") == pkg._lib.kernel_source_sha16())
            ex_tf = cap["fp64_flops_per_attempt"] * my_attempts / (kernel_ms * 1e-3) / 1e12
            # FP64 pipe: 16 lanes per scheduler -> a warp instruction occupies it for 2 cycles
            pipe = 2.0 * cap["fp64_thread_inst_per_attempt"] * my_attempts / 32.0 / cap["warp_execution_efficiency"] / \
                (kernel_ms * 1e-3 * (clocks["sm_mhz"] or 1965.0) * 1e6 * 148 * 4)
            executed = {"tflops": ex_tf, "frac": ex_tf / peak_tf, "fp64_pipe_busy_frac_live": pipe,
                        "fp64_thread_inst_per_attempt": cap["fp64_thread_inst_per_attempt"],
                        "fp64_pipe_active_pct_under_ncu": cap["fp64_pipe_active_pct"],
                        "issue_active_pct_under_ncu": cap["issue_active_pct"],
                        "warp_execution_efficiency": cap["warp_execution_efficiency"],
                        "registers_per_thread": cap["registers_per_thread"],
                        "capture": cap["file"], "capture_kernel_ms": cap["kernel_ms_bench"],
                        "capture_is_of_this_build": same_build,
                        "note": "tflops = (2 DFMA + DMUL + DADD thread instructions per attempt, from the capture) x this "
                                "run's attempts / this run's kernel time; the model-flop `frac` above credits 383 flops per "
                                "RHS and 516 per attempt (SURVEY 8d) although the kernel executes far fewer, so it can exceed 1"}
        line = {
            "metric": "rays_per_s", "value": value, "unit": "rays/s",
            "rhs_evals_per_s": rhs_total / wall_max,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall_max / args.steps, "kernel_ms_per_step": kernel_ms_max / args.steps,
            "kernel_ms_per_rank": per_rank_kernel_ms, "rays_per_rank": per_rank_rays,
            "queue": args.queue, "queue_note": queue_note,
            "frame_rgb8_sha256_16": frame_sha, "single_gpu_rgb8_sha256_16": single_sha,
            "frame_matches_single_gpu": (frame_sha == single_sha) if frame_sha is not None else None,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(scene, L2_NOTE, args.queue),
            "work": {"rays": rays_total / args.steps, "rhs_evals": rhs_total / args.steps,
                     "step_attempts": attempts_total / args.steps, "steps_rejected": rej_total / args.steps,
                     "rhs_per_ray": rhs_total / rays_total, "flops_model": "383*rhs + 516*attempts (SURVEY.md 8d)"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf,
                         "frac_is": "MODEL flops (SURVEY 8d: 383 per RHS + 516 per attempt) / self-measured DFMA peak; see `executed` for what the hardware does",
                         "traffic": cap["dram_bytes_per_launch"] if cap is not None else None,
                         "executed": executed,
                         "peak_source": "self-measured register-resident DFMA chains on this GPU (MEASURED_PEAKS.json has no FP64 entry; nominal 148*64*2*1.965 GHz = 37.2)",
                         "kernel": "trace_kernel<KERR_SCHILD,AS_WRITTEN>", "kernel_ms": kernel_ms / args.steps,
                         "drain_ms": stats_sum["drain_ms"] / args.steps},
            "e2e": e2e, "cpu_baseline": cpu, "reference_scenes": small, "clocks": clocks,
            "gpu_launches": args.steps * world,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
