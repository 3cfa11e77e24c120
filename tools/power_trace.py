"""Evidence for the "power-limited, not HBM-limited" reading of the sustained exhaustive scan:
`nvidia-smi -lms 50` power / SM-clock / throttle-reason trace of (a) 60 back-to-back exhaustive MaxSim scans of the bench
shard (500k x 1030-token pages, 131.8 GB per launch) and (b) a plain device-to-device copy loop (three 20 GB copies per iteration:
60 GB read + 60 GB written), each preceded by 3 s of idle. Prints one JSON line with the per-phase
medians and the achieved GB/s; the raw trace goes to gpurun_out/<tag>_power_trace.csv.

    python tools/power_trace.py [tag] [pages]
"""
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))

import numpy as np
import torch

from visual_rag_b200.corpus import GpuCorpus, query_flags

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
pages = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
out_csv = os.path.join(ROOT, "gpurun_out", f"{tag}_power_trace.csv")
os.makedirs(os.path.dirname(out_csv), exist_ok=True)

rows, marks = [], []
q = "timestamp,power.draw,clocks.sm,clocks.mem,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu"
proc = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)


def reader():
    for line in proc.stdout:
        rows.append((time.time(), line.strip()))


threading.Thread(target=reader, daemon=True).start()

torch.cuda.set_device(0)
c = GpuCorpus(0)
c.add_synthetic_store("initial", pages, fixed_rows=1030, seed=1)
qd = torch.from_numpy(np.random.default_rng(0).standard_normal((20, 128)).astype(np.float32)).cuda()
sc = torch.empty((pages,), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
n_bytes = pages * (1030 * 256 + 1030 * 4)
copy_bytes = 20 << 30                      # the shard itself takes 134 GB: copy buffers of 20 GB each, 3 copies per iteration
src = torch.empty((copy_bytes,), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)


def phase(name, fn, iters, bytes_per_iter=None):
    torch.cuda.synchronize()
    time.sleep(3.0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    marks.append((name, t0, t1, e0.elapsed_time(e1) / iters, bytes_per_iter or n_bytes))


phase("scan", lambda: c.score_dev("initial", qd.data_ptr(), 20, query_flags(True, False), 0, pages, sc.data_ptr(), st), 60)
phase("scan_fp16_query", lambda: c.score_dev("initial", qd.data_ptr(), 20, query_flags(True, False, True), 0, pages, sc.data_ptr(), st), 60)
phase("copy", lambda: [dst.copy_(src) for _ in range(3)], 60, 3 * 2 * copy_bytes)
time.sleep(1.0)
proc.terminate()

with open(out_csv, "w") as f:
    f.write("t_rel_s,phase," + q + "\n")
    t_first = rows[0][0] if rows else 0.0
    for t, line in rows:
        ph = next((m[0] for m in marks if m[1] <= t <= m[2]), "idle")
        f.write(f"{t - t_first:.3f},{ph},{line}\n")

summary = {}
for name, t0, t1, ms, nb in marks:
    sel = [r[1].split(", ") for r in rows if t0 + 0.3 <= r[0] <= t1]
    pw = [float(x[1]) for x in sel if len(x) > 3]
    sm = [float(x[2]) for x in sel if len(x) > 3]
    cap = [x[4].strip().lower().startswith("active") for x in sel if len(x) > 4]
    summary[name] = {"ms_per_iter": ms, "bytes_per_iter": nb, "gbs": nb / (ms * 1e-3) / 1e9, "samples": len(pw),
                     "power_w_median": statistics.median(pw) if pw else None, "power_w_max": max(pw) if pw else None,
                     "sm_mhz_median": statistics.median(sm) if sm else None,
                     "sw_power_cap_active_frac": (sum(cap) / len(cap)) if cap else None}
print(json.dumps({"power_trace": summary, "bytes_per_iter": n_bytes, "csv": os.path.relpath(out_csv, ROOT)}))
c.close()
