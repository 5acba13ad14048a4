#!/usr/bin/env python3
"""Developer measurement (2-GPU box): what the RGB8 patch staging changes on the NVLink.  Device 0 owns a shared
frame and does NOT render; a second process on device 1 renders the whole frame, so every pixel is a peer store
into device 0's memory.  That process runs under ncu (single-pass metric groups only: a replayed pass would find
the queue drained) once with the staging kernel (the default for a remote image) and once with RTGR_RGB8_STAGING=0
(byte stores); the image must equal device 0's own render both times.  usage: peer_store_probe.py <out dir> [workload ni nj]"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

out_dir = sys.argv[1]
name, ni, nj = (sys.argv[2], int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("config4", 1920, 1080)
pkg = entry.load_package()
sc = pkg.scenes.BY_NAME[name]().with_size(ni, nj)
ctx = pkg.Context([0])
ref = ctx.render(sc, want=("rgb8",))["rgb8"]
GROUPS = {
    "stores": "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "nvlink": "nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_packet_request_data_protocol.sum",
}
for label, flag in (("default (remote image)", None), ("staged", "1"), ("bytestores", "0")):
    row = {"stores": label, "workload": sc.name, "ni": ni, "nj": nj}
    for group, metrics in GROUPS.items():
        frame = pkg.Frame(ctx, ni, nj)
        log = os.path.join(out_dir, "peer_store_%s_%s.csv" % (label.split()[0], group))
        env = dict(os.environ)
        env.pop("RTGR_RGB8_STAGING", None)
        if flag is not None:
            env["RTGR_RGB8_STAGING"] = flag
        cmd = ["ncu", "--metrics", metrics, "--clock-control", "none", "-k", "regex:trace_", "--csv", "--log-file", log,
               sys.executable, os.path.join(ROOT, "tests", "frame_peer.py"), "1", frame.handle.hex(), name, str(ni), str(nj), "1"]
        peer = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=ROOT, env=env)
        line = ""
        while True:                       # (ncu prints its own ==PROF== lines on the same stdout)
            line = peer.stdout.readline()
            if not line or line.strip() == "ready" or line.startswith("open-failed"):
                break
        if line.strip() != "ready":
            row[group] = "peer failed: " + line.strip()
            peer.kill()
            frame.close()
            continue
        peer.stdin.write("go\n"); peer.stdin.flush()
        while True:
            line = peer.stdout.readline()
            if not line or line.startswith("done"):
                break
        peer.stdin.close()
        peer.wait(timeout=120)
        img = frame.read()
        row["image_equals_single_gpu_render_" + group] = bool((img == ref).all())
        row["rays_" + group] = int(line.split()[1]) if line.startswith("done") else None
        try:
            for r in csv.DictReader(l for l in open(log) if l.startswith('"')):
                row[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
                row[r["Metric Name"] + "_unit"] = r["Metric Unit"]
        except Exception as e:      # noqa: BLE001
            row[group] = "no metrics: %s" % e
        frame.close()
    row["rgb8_sha"] = hashlib.sha256(ref.tobytes()).hexdigest()[:16]
    print(json.dumps(row), flush=True)
ctx.close()
