#!/usr/bin/env python
"""Parity fuzz of the PAWCS kernels (test infrastructure: it executes the oracle, so it lives under tests/; run once per kernel change on the GPU box): random frame sizes incl. tiny and
ragged ones, gray / RGB, random ROIs, learning-rate overrides; the CUDA path against the CPU oracle (snapshot mode), full dictionary state
after every frame. usage: python tests/fuzz_pawcs.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import litiv_b200 as lv
from oracle import oracle as O
from litiv_b200.synth import SynthSequence
import test_gpu_pawcs as TP

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for k in range(ncases):
    w, h = int(rng.integers(24, 360)), int(rng.integers(24, 260))
    c = int(rng.choice([1, 3]))
    nf = int(rng.integers(6, 14))
    seed = int(rng.integers(0, 1 << 30))
    roi = None
    if rng.random() < 0.5:
        roi = np.zeros((h, w), np.uint8)
        x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2))
        roi[min(y0, h - 12):max(y1, min(y0, h - 12) + 12), min(x0, w - 12):max(x1, min(x0, w - 12) + 12)] = 255
        if rng.random() < 0.5:
            roi[rng.random((h, w)) < 0.05] = 0          # holes in the ROI
    tag = f"case {k}: {w}x{h}x{c}, {nf} frames, seed {seed}, roi {'yes' if roi is not None else 'no'}"
    try:
        seq = SynthSequence(w, h, c, seed=seed & 0xFFFF)
        g, o = lv.BackgroundSubtractorPAWCS(seed=seed), O.Oracle(O.ALGO_PAWCS, mode=O.MODE_SNAPSHOT, seed=seed)
        f0 = seq.frame(0)
        g.initialize(f0, roi); o.initialize(f0, roi)
        TP._compare(g, o, "init", skip=("rawmask",))
        for t in range(1, nf + 1):
            f = seq.frame(t)
            if rng.random() < 0.1:
                f = 255 - f
            lr = float(rng.choice([0.0, 0.0, 0.0, 1.0, 3.0, 50.0]))
            mg, mo = g.apply(f, lr), o.apply(f, lr)
            assert np.array_equal(mg, mo), f"frame {t}: masks differ in {(mg != mo).sum()} px"
            TP._compare(g, o, f"frame {t}")
        print("ok  ", tag, flush=True)
    except Exception as e:     # noqa: BLE001
        bad += 1
        print("FAIL", tag, "->", str(e)[:300], flush=True)
print(f"{ncases - bad} of {ncases} cases bit-exact")
sys.exit(1 if bad else 0)
