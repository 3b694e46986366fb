#!/usr/bin/env python
"""Tuning helper (not part of the product): PAWCS frame time at 640x480 RGB (BASELINE.json configs[2]) or WxH given."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import litiv_b200 as lv
from litiv_b200.synth import SynthSequence

W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
nboot = int(sys.argv[3]) if len(sys.argv) > 3 else 60
C, NU = 3, 24
seq = SynthSequence(W, H, C, seed=4)
frames = [seq.frame(t) for t in range(NU)]
dev = torch.device("cuda", 0)
pitch = (W * C + 127) // 128 * 128
d_frames = torch.zeros((NU, H, pitch), dtype=torch.uint8, device=dev)
for i, f in enumerate(frames):
    d_frames[i, :, :W * C] = torch.from_numpy(f.reshape(H, W * C)).to(dev)
d_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
sub = lv.BackgroundSubtractorPAWCS(device=0, seed=0)
sub.initialize(frames[0])
stream = torch.cuda.ExternalStream(sub.stream, device=dev)
k = 0
def pp(i, n):
    period = 2 * (n - 1); j = i % period
    return j if j < n else period - j
def step():
    global k
    k += 1
    sub.apply_device(d_frames[pp(k, NU)].data_ptr(), pitch, d_mask.data_ptr(), 0.0)
for _ in range(nboot):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
n = 64
for _ in range(n):
    step()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
sub.set_profile(True); sub.set_collect_stats(True)
for _ in range(32):
    step()
torch.cuda.synchronize()
pms, pn = sub.get_profile()
st = sub.stats()
print(json.dumps({"W": W, "H": H, "frame_us": ms * 1e3, "mpx_s": W * H / ms / 1e3, "scan_plus_tail_us": pms / pn * 1e3,
                  "words_scanned_per_px": st["samples_scanned"] / max(st["roi_px"], 1), "fg_frac": st["fg_px"] / max(st["roi_px"], 1)}))
