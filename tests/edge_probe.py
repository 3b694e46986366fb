"""TEST INFRASTRUCTURE (it executes the oracle, so it lives under tests/): one-shot GPU probe of the edge detector against the oracle (a few sizes, printed mismatch counts and wall-clock times); the full suite is tests/test_gpu_edge.py."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import litiv_b200 as lv  # noqa: E402
from litiv_b200.synth import SynthSequence  # noqa: E402
from oracle import oracle as O  # noqa: E402

bad = 0
for (w, h), ch, levels in [((96, 72), 3, 3), ((97, 73), 1, 3), ((97, 73), 3, 1), ((320, 240), 3, 3), ((641, 479), 1, 2), ((1920, 1080), 3, 3)]:
    seq = SynthSequence(w, h, ch, seed=w + h + ch)
    o, e = O.EdgeDetectorLBSPOracle(levels=levels), lv.EdgeDetectorLBSP(levels)
    for t, thr in [(3, 0.5), (5, 0.25), (9, 0.0)]:
        f = np.ascontiguousarray(seq.frame(t))
        f = np.ascontiguousarray(f[..., 0]) if ch == 1 and f.ndim == 3 else f
        want = o.apply_threshold(f, thr)
        t0 = time.perf_counter()
        got = e.apply_threshold(f, thr)
        dt = time.perf_counter() - t0
        gd = int((e.gradient_map() != o.gradient_map(f.shape)).sum())
        md = int((got != want).sum())
        bad += (gd > 0) + (md > 0)
        print(f"{w}x{h}x{ch} L{levels} thr {thr}: grad diff {gd}, mask diff {md}, edge px {int((want > 0).sum())}, sweeps {e.flood_sweeps()}, {dt * 1e3:.2f} ms", flush=True)
    if w <= 320:
        f = np.ascontiguousarray(seq.frame(14))
        f = np.ascontiguousarray(f[..., 0]) if ch == 1 and f.ndim == 3 else f
        cd = int((e.apply(f) != o.apply(f)).sum())
        bad += cd > 0
        print(f"  confidence map diff {cd}", flush=True)
print("EDGE PROBE", "OK" if not bad else f"FAILED ({bad})", "launches", lv.kernel_launch_count(), flush=True)
