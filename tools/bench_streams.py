#!/usr/bin/env python
"""BASELINE.json configs[4]: SuBSENSE 640x480 RGB, S independent streams on one GPU (device-resident frames, round-robin
enqueue from one host thread, every stream on its own CUDA streams). Prints one JSON line (aggregate Mpx/s, streams x fps).
usage: python tools/bench_streams.py [--streams 64] [--steps 100] [--size 640x480]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import litiv_b200 as lv
from litiv_b200.synth import SynthSequence

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=64)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--size", default="640x480")
ap.add_argument("--algo", default="subsense", choices=["subsense", "lobster", "pawcs"])
ap.add_argument("--channels", type=int, default=3)
ap.add_argument("--batch", action="store_true", help="enqueue through lvb_apply_batch_device (pool of host threads) instead of one call per stream")
ap.add_argument("--unique", type=int, default=4, help="distinct sequences (streams share synthetic frames, each has its own model and seed)")
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
C, NF, BOOT = args.channels, 12, 60
dev = torch.device("cuda", 0)
pitch = (W * C + 127) // 128 * 128
seqs = [SynthSequence(W, H, C, seed=5000 + i) for i in range(args.unique)]
d_frames = torch.zeros((args.unique, NF, H, pitch), dtype=torch.uint8, device=dev)
host0 = []
for u, sq in enumerate(seqs):
    for t in range(NF):
        f = sq.frame(t)
        if t == 0: host0.append(f)
        d_frames[u, t, :, :W * C] = torch.from_numpy(np.ascontiguousarray(f).reshape(H, W * C)).to(dev)
d_masks = torch.zeros((args.streams, H, W), dtype=torch.uint8, device=dev)
subs = []
for s in range(args.streams):
    b = {"subsense": lv.BackgroundSubtractorSuBSENSE, "lobster": lv.BackgroundSubtractorLOBSTER, "pawcs": lv.BackgroundSubtractorPAWCS}[args.algo](device=0, seed=5000 + s)
    b.initialize(host0[s % args.unique])
    subs.append(b)

def pp(i, n):
    k = i % (2 * (n - 1)); return k if k < n else 2 * (n - 1) - k

batch = lv.DeviceBatch(subs) if args.batch else None
mask_ptrs = [d_masks[s].data_ptr() for s in range(args.streams)]
frame_ptrs = [[d_frames[s % args.unique, t].data_ptr() for s in range(args.streams)] for t in range(NF)]

def round_(k):
    lr = 16.0 if args.algo == "lobster" else (1.0 if k <= 50 else 0.0)
    t = pp(k, NF)
    if batch is not None:
        batch.apply(frame_ptrs[t], pitch, mask_ptrs, lr)
        return
    for s, b in enumerate(subs):
        b.apply_device(d_frames[s % args.unique, t].data_ptr(), pitch, d_masks[s].data_ptr(), lr)

k = 0
for _ in range(BOOT + 5):
    k += 1; round_(k)
torch.cuda.synchronize()
l0 = lv.kernel_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
import time; t0 = time.perf_counter()
for _ in range(args.steps):
    k += 1; round_(k)
host_s = time.perf_counter() - t0
torch.cuda.synchronize()
wall_s = time.perf_counter() - t0
px = W * H * args.streams * args.steps
print(json.dumps({"metric": args.algo + "_multistream_mpx_per_s", "value": px / wall_s / 1e6, "unit": "Mpx/s", "streams": args.streams, "frame": [W, H, C],
                  "steps": args.steps, "fps_per_stream": args.steps / wall_s, "streams_x_fps": args.streams * args.steps / wall_s,
                  "ms_per_round": wall_s / args.steps * 1e3, "host_enqueue_ms_per_round": host_s / args.steps * 1e3,
                  "gpu_launches": lv.kernel_launch_count() - l0, "enqueue": "lvb_apply_batch_device" if args.batch else "lvb_apply_device per stream", "timing": "wall clock around enqueue + device synchronize (work spans many CUDA streams)"}))
