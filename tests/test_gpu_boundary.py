"""GPU tests of the boundary additions of round 2 (run with -m gpu on the B200 box): validateROI, device-pointer getBackgroundImage,
lvb_apply_stream, the frame-level model reset under CUDA_LAUNCH_BLOCKING=1 (its in-kernel wait for the mask stream is bounded and must not
hang or time out when launches are serialised), caller-buffer validation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_validate_roi_and_set_roi(lv, oracle):
    w, h = 160, 120
    seq = SynthSequence(w, h, 3, seed=3)
    g = lv.BackgroundSubtractorSuBSENSE(seed=1)
    roi = np.full((h, w), 255, np.uint8)
    roi[40:50, 60:90] = 0
    v = g.validateROI(roi.copy())
    assert not v[:2].any() and not v[-2:].any() and not v[:, :2].any() and not v[:, -2:].any()
    assert np.array_equal(v[2:-2, 2:-2], roi[2:-2, 2:-2])
    g.initialize(seq.frame(0), roi)
    got = g.getROICopy()
    assert np.array_equal(got == 255, v == 255)   # (the final ROI also carries the 128-valued ring the reference adds around ROI borders)
    for t in range(1, 5):
        g.apply(seq.frame(t), 1.0)
    roi2 = np.full((h, w), 255, np.uint8)
    g.setROI(roi2)                                   # re-initialises from the current background image (BackgroundSubtractionUtils.cpp:38-48)
    assert (g.getROICopy()[2:-2, 2:-2] > 0).all()
    with pytest.raises(lv.LitivError, match="ROI"):
        g.setROI(np.zeros((h + 1, w), np.uint8))
    with pytest.raises(lv.LitivError, match="ROI"):
        g.validateROI(np.zeros((h, w), np.float32))


def test_get_background_image_device_matches_host(lv):
    torch = pytest.importorskip("torch")
    w, h = 320, 240
    seq = SynthSequence(w, h, 3, seed=5)
    g = lv.BackgroundSubtractorSuBSENSE(seed=2)
    g.initialize(seq.frame(0))
    for t in range(1, 8):
        g.apply(seq.frame(t), 1.0)
    d = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    g.getBackgroundImageDevice(d.data_ptr())
    assert np.array_equal(d.cpu().numpy(), g.getBackgroundImage())
    with pytest.raises(lv.LitivError, match="device pointer"):
        g.getBackgroundImageDevice(np.zeros((h, w, 3), np.uint8).ctypes.data)


def test_apply_stream_matches_synchronous_apply(lv):
    w, h, n = 320, 240, 14
    seq = SynthSequence(w, h, 3, seed=31)
    a, b = lv.BackgroundSubtractorSuBSENSE(seed=4), lv.BackgroundSubtractorSuBSENSE(seed=4)
    a.initialize(seq.frame(0)); b.initialize(seq.frame(0))
    frames = [seq.frame(t) for t in range(1, n + 1)]
    lrs = [1.0 if t < 6 else 0.0 for t in range(n)]
    want = [b.apply(f, lr) for f, lr in zip(frames, lrs)]
    got = a.apply_stream(frames, lrs)
    for t in range(n):
        assert np.array_equal(got[t], want[t]), f"frame {t}"
    for name in ("bg_color", "bg_desc", "R", "T", "lastfg", "unstable"):
        assert np.array_equal(a.state_get(name), b.state_get(name)), name
    with pytest.raises(lv.LitivError, match="output mask"):
        a.apply(frames[0], 0.0, out=np.zeros((h, w + 1), np.uint8))
    with pytest.raises(lv.LitivError, match="output mask"):
        a.apply_async(frames[0], 0.0, out=np.zeros((h, w), np.float32))


BLOCKING_SCRIPT = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import litiv_b200 as lv
from litiv_b200.synth import SynthSequence
from oracle import oracle as O
w, h, c = 320, 240, 3
seq_a, seq_b = SynthSequence(w, h, c, seed=21), SynthSequence(w, h, c, seed=22)
g = lv.BackgroundSubtractorSuBSENSE(seed=5)
o = O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_SNAPSHOT, seed=5)
f0 = seq_a.frame(0)
g.initialize(f0); o.initialize(f0)
e0 = o.state_get("scalars")[12]
for t in range(1, 96):
    f = seq_a.frame(t) if t < 60 else (255 - seq_b.frame(t))
    lr = 1.0 if t <= 10 else 0.0
    assert np.array_equal(g.apply(f, lr), o.apply(f, lr)), t
g.sync()   # raises if the frame tail's bounded wait for the mask chain timed out
assert o.state_get("scalars")[12] > e0 and g.state_get("scalars")[12] == o.state_get("scalars")[12]
assert np.array_equal(g.state_get("bg_color"), o.state_get("bg_color"))
print("BLOCKING_OK resets", int(g.state_get("scalars")[12] - e0))
'''


def test_model_reset_under_launch_blocking(lv):
    """the conditional refresh waits inside a kernel for the mask stream; with CUDA_LAUNCH_BLOCKING=1 every launch is synchronous, so the wait
    only terminates because the mask chain is always enqueued before the kernel that waits for it"""
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    out = subprocess.run([sys.executable, "-c", BLOCKING_SCRIPT % ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "BLOCKING_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
