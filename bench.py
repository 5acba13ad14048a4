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

# One `ncu --set full` capture of trace_kernel on the default workload (config4, 3840x2160, 1 GPU):
# profiles/r01_trace_kernel_4k_ncu_raw.csv.  Static facts quoted beside the live numbers; they are
# NOT re-measured by this script.
NCU_4K = {
    "source": "profiles/r01z_trace_kernel_4k_ncu_raw.csv",
    "dram_bytes_per_launch": 493312 + 16817408,         # dram__bytes_read.sum + dram__bytes_write.sum
    "fp64_pipe_active_pct": 71.09,                      # sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
    "executed_tflops": 20.58,                           # (2*dfma + dmul + dadd thread-inst/cycle) * 1.962 GHz
    "fp64_thread_inst_per_attempt": 1166,               # dfma + dmul + dadd, per step attempt (6 RHS + the rest)
    "warp_execution_efficiency": 31.11 / 32,            # smsp__thread_inst_executed_per_inst_executed
    "capture_of": "the round-1 kernel as profiled in profiles/r01z (later hot-loop savings -- integer |x| in the error norm, "
                  "a 3-instruction coarse event bound, no fourth distance chain for <= 3 objects, 140 instead of 143 FP64 "
                  "instructions per RHS: about -43 FP64 instructions per attempt by static SASS count, tests/sass_mix.py "
                  "-- are newer than this capture)",
    "note": "the kernel executes fewer flops than the 383/516 model credits (leaner RHS than the model), so frac "
            "(model flops / peak) reads above the executed-flop fraction; three-register-operand DFMA code tops out "
            "at 69 % of the DFMA peak on this part (profiles/r01z_fp64_modes.log)",
}


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
        "config": workload_config(scene, "n/a (CPU)"),
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
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # FP64 roofline denominator: register-resident DFMA chains on this GPU (best of several launches)
    peak_tf = max(ctx.fp64_peak(0)[0] for _ in range(3))

    # ---------------- kernel path: inputs resident, nothing copied ----------------
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
        if allreduce(ok, dist.ReduceOp.MIN if world > 1 else None) < 1.0:
            # e.g. a sandbox without CUDA IPC: all ranks use the static deal (still the GPU path)
            if frame is not None:
                frame.close()
            frame, args.queue = None, "static"
            queue_note = "shared frame unavailable on this box (%s): static deal" % (why or "another rank failed to map it")
            if rank == 0:
                print("bench.py: " + queue_note, file=sys.stderr)

    def step_kernel():
        flush.zero_()
        torch.cuda.synchronize()
        if frame is not None:
            barrier()       # the frame protocol's barrier between consecutive frames (inside the timed region)
            return frame.render(scene)
        return ctx.render_resident(scene, tile_offset=rank, tile_stride=world)

    for _ in range(args.warmup):
        step_kernel()
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
    frame_checksum = None
    if frame is not None:
        barrier()
        if rank == 0:
            frame_checksum = int(frame.read().astype(np.uint64).sum())
        barrier()
        frame.close()
    value = rays_total / wall_max
    # roofline of the trace kernel on this rank (rank 0 reports its own kernel)
    my_flops = W_RHS * stats_sum["rhs_evals"] + W_STEP * (stats_sum["steps_accepted"] + stats_sum["steps_rejected"])
    achieved_tf = my_flops / (kernel_ms * 1e-3) / 1e12

    # ---------------- end to end: the trace_rays drop-in with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        p, objs, nobj, cam = pkg.scenes.to_abi(scene)
        # The caller's input: the whole canvas (Array{Pixel{Float64},2}, 88 B per pixel) in page-locked
        # host memory, built once outside the timed region.  Every rank traces its tiles of it in place.
        buf = pkg.PinnedArray((scene.nj, scene.ni, 11))
        buf.array[...] = ctx.make_canvas(p, cam).reshape(scene.nj, scene.ni, 11)
        canvas = buf.array

        def step_e2e():
            canvas[:, :, 8:] = 0.0           # results of the previous step cannot be reused
            flush.zero_()
            torch.cuda.synchronize()
            return ctx.trace_canvas(p, objs, nobj, canvas, tile_offset=rank, tile_stride=world)["stats"]

        for _ in range(args.warmup):
            step_e2e()
        barrier()
        e2e_call_s, my_rays = 0.0, 0
        for _ in range(args.steps):
            canvas[:, :, 8:] = 0.0
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            st = ctx.trace_canvas(p, objs, nobj, canvas, tile_offset=rank, tile_stride=world)["stats"]
            barrier()
            e2e_call_s += time.perf_counter() - t0
            my_rays = st["rays"]
        e2e_wall = allreduce(e2e_call_s, dist.ReduceOp.MAX if world > 1 else None)
        checksum = allreduce(float(canvas[:, :, 8:].sum()), dist.ReduceOp.SUM if world > 1 else None)
        e2e = {"value": n_rays * args.steps / e2e_wall, "unit": "rays/s", "ms_per_step": 1e3 * e2e_wall / args.steps,
               "h2d_bytes_per_step": int(my_rays * 64), "d2h_bytes_per_step": int(my_rays * 24),
               "api": "rtgr_trace_canvas (drop-in for trace_rays, src:483) on a page-locked host Array{Pixel}: the kernel "
                      "reads pos/normal (64 B/ray) from and writes rgb (24 B/ray) into HOST memory in place over PCIe "
                      "while it computes; bytes are per rank; timed from call to return, barrier on both sides",
               "rgb_checksum": checksum}
        # the same call on PAGEABLE host memory (staged: whole-canvas H2D, trace, D2H), N = 1 only
        if world == 1:
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
            ctx.render(scene, want=("rgb8",), tile_offset=rank, tile_stride=world, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.render(scene, want=("rgb8",), tile_offset=rank, tile_stride=world, out=out)
        barrier()
        r_wall = allreduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        e2e["render_rgb8"] = {"value": n_rays * args.steps / r_wall, "unit": "rays/s",
                              "api": "rtgr_render_tiles: device-side make_canvas, RGB8 frame copied to host",
                              "d2h_bytes_per_step": int(n_rays * 3)}
        buf.free()

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(pkg, scene, 0, steps=1, warmup=0, budget_s=15.0)
        cpu = {"value": r["rays_per_s"], "unit": "rays/s", "rhs_evals_per_s": r["rhs_per_s"], "cores": r["cores"],
               "kind": "port", "cpu": cpu_model(),
               "sample": "every %d-th pixel in i and j of the %dx%d frame = %d rays, %.1f s"
                         % (r["stride"], scene.ni, scene.nj, r["rays"], r["sec_per_step"]),
               "note": "C++ restatement of the reference path in its as-written operation order "
                       "(Julia unavailable in image), OpenMP schedule(static) over the pixel index"}

    if rank == 0:
        line = {
            "metric": "rays_per_s", "value": value, "unit": "rays/s",
            "rhs_evals_per_s": rhs_total / wall_max,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall_max / args.steps, "kernel_ms_per_step": kernel_ms_max / args.steps,
            "kernel_ms_per_rank": per_rank_kernel_ms, "rays_per_rank": per_rank_rays,
            "queue": args.queue, "queue_note": queue_note, "frame_rgb8_checksum": frame_checksum,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(scene, "256 MB device memset between steps (L2 is 126 MB); inputs are <1 KB of scene constants", args.queue),
            "work": {"rays": rays_total / args.steps, "rhs_evals": rhs_total / args.steps,
                     "step_attempts": attempts_total / args.steps, "steps_rejected": rej_total / args.steps,
                     "rhs_per_ray": rhs_total / rays_total, "flops_model": "383*rhs + 516*attempts (SURVEY.md 8d)"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf,
                         "traffic": NCU_4K["dram_bytes_per_launch"] if (scene.name == "ks_a0.99_4k_wide" and scene.ni == 3840 and world == 1) else None,
                         "ncu": NCU_4K if (scene.name == "ks_a0.99_4k_wide" and scene.ni == 3840 and world == 1) else None,
                         "peak_source": "self-measured register-resident DFMA chains on this GPU (MEASURED_PEAKS.json has no FP64 entry; nominal 148*64*2*1.965 GHz = 37.2)",
                         "kernel": "trace_kernel<KERR_SCHILD,AS_WRITTEN>", "kernel_ms": kernel_ms / args.steps,
                         "drain_ms": stats_sum["drain_ms"] / args.steps},
            "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
            "gpu_launches": args.steps * world,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
