#!/usr/bin/env python
"""Tuning helper (not part of the product): time phase A per launch at 1080p for the library named by LVB_SO.
usage: LVB_SO=... python tools/exp_phaseA.py [nsamples] [label]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import litiv_b200 as lv
from bench import make_frames, pingpong, lr_for, W, H, C, BOOT_FRAMES, N_UNIQUE

nsamples = int(sys.argv[1]) if len(sys.argv) > 1 else 50
label = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("LVB_SO", "default")
seq, frames = make_frames(4, N_UNIQUE)
dev = torch.device("cuda", 0)
pitch = (W * C + 127) // 128 * 128
d_frames = torch.zeros((N_UNIQUE, H, pitch), dtype=torch.uint8, device=dev)
for i, f in enumerate(frames):
    d_frames[i, :, :W * C] = torch.from_numpy(f.reshape(H, W * C)).to(dev)
d_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
sub = lv.BackgroundSubtractorSuBSENSE(nBGSamples=nsamples, device=0, seed=0)
sub.initialize(frames[0])
stream = torch.cuda.ExternalStream(sub.stream, device=dev)
k = 0
def step():
    global k
    k += 1
    sub.apply_device(d_frames[pingpong(k, N_UNIQUE)].data_ptr(), pitch, d_mask.data_ptr(), lr_for(k))
for _ in range(BOOT_FRAMES + 10):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(100):
    step()
sub.flush()
e1.record(stream)
torch.cuda.synchronize()
frame_ms = e0.elapsed_time(e1) / 100
sub.set_profile(True); sub.set_collect_stats(True)
for _ in range(50):
    step()
torch.cuda.synchronize()
ms, n = sub.get_profile()
st = sub.stats()
print(json.dumps({"label": label, "N": nsamples, "phaseA_us": ms / n * 1e3, "frame_us": frame_ms * 1e3,
                  "scan_depth": st["samples_scanned"] / max(st["roi_px"], 1), "fg_frac": st.get("fg_px", 0) / max(st["roi_px"], 1)}))
