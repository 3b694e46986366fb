#!/usr/bin/env python
"""ViBe throughput on one GPU: device-resident frames, CUDA events on the instance's stream, the scan kernel
timed per launch through lvb_vibe_set_profile. One JSON line. (The CPU restatement is timed beside the GPU by
tests/test_gpu_vibe.py::test_cpu_restatement_timed_beside_the_gpu: only tests/ and bench.py may execute oracle/.)
usage: python tools/bench_vibe.py [--size 1920x1080] [--channels 3] [--steps 300]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import litiv_b200 as lv
from litiv_b200.synth import SynthSequence

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--channels", type=int, default=3)
ap.add_argument("--steps", type=int, default=300)
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
C, NF = args.channels, 12
dev = torch.device("cuda", 0)
seq = SynthSequence(W, H, C, seed=4100)
host = [seq.frame(t) for t in range(NF)]
pitch = (W * C + 127) // 128 * 128
d_frames = torch.zeros((NF, H, pitch), dtype=torch.uint8, device=dev)
for t in range(NF):
    d_frames[t, :, :W * C] = torch.from_numpy(np.ascontiguousarray(host[t]).reshape(H, W * C)).to(dev)
d_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
v = (lv.BackgroundSubtractorViBe_1ch if C == 1 else lv.BackgroundSubtractorViBe_3ch)(seed=4100)
v.initialize(host[0])

def pp(i, n):
    k = i % (2 * (n - 1)); return k if k < n else 2 * (n - 1) - k

k = 0
for _ in range(40):
    k += 1; v.apply_device(d_frames[pp(k, NF)].data_ptr(), C, pitch, d_mask.data_ptr(), 16.0)
v.sync()
# scan depth / writes per pixel: an untimed instrumented pass (the counters cost atomics)
v.set_collect_stats(True)
s0 = v.stats()
for _ in range(24):
    k += 1; v.apply_device(d_frames[pp(k, NF)].data_ptr(), C, pitch, d_mask.data_ptr(), 16.0)
s1 = v.stats()
v.set_collect_stats(False)
st = torch.cuda.ExternalStream(v.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(args.steps):
    k += 1; v.apply_device(d_frames[pp(k, NF)].data_ptr(), C, pitch, d_mask.data_ptr(), 16.0)
e1.record(st)
v.sync()
ms = e0.elapsed_time(e1) / args.steps
px = s1["roi_px"] - s0["roi_px"]
depth = (s1["samples_scanned"] - s0["samples_scanned"]) / px
writes = (s1["sample_writes"] - s0["sample_writes"]) / px
v.set_profile(True)
for _ in range(50):
    k += 1; v.apply_device(d_frames[pp(k, NF)].data_ptr(), C, pitch, d_mask.data_ptr(), 16.0)
pms, pn = v.get_profile()
v.set_profile(False)
# end to end: host frame -> host mask through lvb_vibe_apply (pinned buffers)
hf = [lv.pinned_empty(host[0].shape) for _ in range(NF)]
for t in range(NF):
    hf[t][...] = host[t]
hm = lv.pinned_empty((H, W))
t0 = time.perf_counter()
for i in range(60):
    k += 1; v.apply(hf[pp(k, NF)], 16.0, out=hm)
e2e_ms = (time.perf_counter() - t0) / 60 * 1e3
# algorithmic bytes per pixel: frame C + mask 1 + sample bytes scanned / written (C each)
sample_b = C
b_alg = C + 1 + sample_b * (depth + writes)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 0) or 0) or 6459.0
scan_ms = pms / max(pn, 1)
print(json.dumps({
    "metric": "vibe_mpx_per_s", "value": W * H / ms / 1e3, "unit": "Mpx/s", "ms_per_frame": ms, "fps": 1e3 / ms, "frame": [W, H, C],
    "e2e": {"value": W * H / e2e_ms / 1e3, "unit": "Mpx/s", "api": "lvb_vibe_apply(host frame, host mask), synchronous, pinned buffers"},
    "scan_depth": depth, "sample_writes_per_px": writes, "alg_bytes_per_px": b_alg,
    "roofline": {"bound": "hbm", "kernel": "vibe_phaseA", "avg_launch_ms": scan_ms, "achieved": W * H * b_alg / (scan_ms * 1e-3) / 1e9, "peak": peak,
                 "unit": "GB/s", "frac": W * H * b_alg / (scan_ms * 1e-3) / 1e9 / peak, "kernel_share_of_step": scan_ms / ms},
}))
