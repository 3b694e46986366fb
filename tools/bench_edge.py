#!/usr/bin/env python
"""EdgeDetectorLBSP throughput on one GPU: device-resident frames through lvb_edge_apply_threshold_device with CUDA
events on the detector's stream, and host frames through lvb_edge_apply_threshold. One JSON line. First checks the device-resident
variant against the host variant (the entry point was added without a GPU at hand).
usage: python tools/bench_edge.py [--size 1920x1080] [--channels 3] [--levels 3] [--steps 200]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import litiv_b200 as lv
from litiv_b200.synth import SynthSequence

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--channels", type=int, default=3)
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
C, NF = args.channels, 8
dev = torch.device("cuda", 0)
seq = SynthSequence(W, H, C, seed=4400)
host = [np.ascontiguousarray(seq.frame(t)) for t in range(1, NF + 1)]
pitch = (W * C + 127) // 128 * 128
d_frames = torch.zeros((NF, H, pitch), dtype=torch.uint8, device=dev)
for t in range(NF):
    d_frames[t, :, :W * C] = torch.from_numpy(host[t].reshape(H, W * C)).to(dev)
d_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
e, ref = lv.EdgeDetectorLBSP(args.levels), lv.EdgeDetectorLBSP(args.levels)
for t in range(NF):   # device-resident variant == host variant (itself parity-tested against the oracle)
    e.apply_threshold_device(d_frames[t].data_ptr(), W, H, C, pitch, d_mask.data_ptr(), 0.5)
    want = ref.apply_threshold(host[t], 0.5)
    assert np.array_equal(d_mask.cpu().numpy(), want), f"device-resident variant differs from the host variant on frame {t}"
st = torch.cuda.ExternalStream(e.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = lv.kernel_launch_count()
e0.record(st)
for i in range(args.steps):
    e.apply_threshold_device(d_frames[i % NF].data_ptr(), W, H, C, pitch, d_mask.data_ptr(), 0.5)
e1.record(st)
e1.synchronize()
ms = e0.elapsed_time(e1) / args.steps
launches = (lv.kernel_launch_count() - l0) / args.steps
t0 = time.perf_counter()
for i in range(60):
    ref.apply_threshold(host[i % NF], 0.5)
e2e_ms = (time.perf_counter() - t0) / 60 * 1e3
# algorithmic bytes per pixel of one call: every level reads its image (C) and writes / re-reads its 4-byte map (gradient write, combine
# read + write, coarse read /4), level sizes 1 + 1/4 + 1/16; suppression reads the map once (4) and writes the mask (1); every flood
# sweep reads the mask (1); output reads the mask and writes the result (2)
lv_sum = sum(0.25 ** l for l in range(args.levels))
b_alg = lv_sum * (C + 4 + 8 + 1) + (lv_sum - 1) * C + 4 + 1 + e.flood_sweeps() + 2
peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = float((json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}).get("hbm_gbs", 0) or 0) or 6459.0
print(json.dumps({
    "metric": "edge_lbsp_mpx_per_s", "value": W * H / ms / 1e3, "unit": "Mpx/s", "ms_per_call": ms, "frame": [W, H, C], "levels": args.levels,
    "launches_per_call": launches, "flood_sweeps": e.flood_sweeps(),
    "e2e": {"value": W * H / e2e_ms / 1e3, "unit": "Mpx/s", "ms_per_call": e2e_ms, "api": "lvb_edge_apply_threshold(host image, host mask), synchronous, pageable buffers"},
    "roofline": {"bound": "hbm", "scope": "whole call (launch bound at small sizes)", "alg_bytes_per_px": b_alg, "achieved": W * H * b_alg / (ms * 1e-3) / 1e9, "peak": peak,
                 "unit": "GB/s", "frac": W * H * b_alg / (ms * 1e-3) / 1e9 / peak},
}))
