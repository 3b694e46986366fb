#!/usr/bin/env python
"""Parity fuzz of the SuBSENSE / LOBSTER paths (test infrastructure: it executes the oracle, so it lives under tests/; run once per kernel change on the GPU box): random frame sizes,
gray / RGB, random ROIs, scene cuts (frame-level reset -> device-side refreshModel), host operations between frames that flush the queued
sample writes (state export, getBackgroundImage), host refreshModel, sample-model re-import, runs that cross the 128-frame rebuild of the
colour boxes; the CUDA path against the CPU oracle (snapshot mode), masks every frame and the full state at random frames.
usage: python tests/fuzz_subsense.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import litiv_b200 as lv
from oracle import oracle as O
from litiv_b200.synth import SynthSequence
import test_gpu_parity as TP

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for k in range(ncases):
    algo = "subsense" if rng.random() < 0.75 else "lobster"
    big = rng.random() < 0.3
    w, h = (int(rng.integers(320, 420)), int(rng.integers(240, 300))) if big else (int(rng.integers(24, 260)), int(rng.integers(24, 200)))
    c = int(rng.choice([1, 3]))
    nf = int(rng.integers(140, 200)) if rng.random() < 0.4 else int(rng.integers(20, 60))
    seed = int(rng.integers(0, 1 << 30))
    roi = None
    if rng.random() < 0.4:
        roi = np.zeros((h, w), np.uint8)
        roi[h // 7:h - h // 9, w // 6:w - 2] = 255
        roi[rng.random((h, w)) < 0.03] = 0
    tag = f"case {k}: {algo} {w}x{h}x{c}, {nf} frames, seed {seed}, roi {'yes' if roi is not None else 'no'}"
    try:
        seq = SynthSequence(w, h, c, seed=seed & 0xFFFF, fg_area=float(rng.choice([0.02, 0.09, 0.2])))
        g, o = TP._mk(lv, O, algo, seed=seed)
        ints = [n for n in TP.INT_STATE if algo == "subsense" or n in ("lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "rawmask")]
        flts = TP.FLT_STATE if algo == "subsense" else []
        f0 = seq.frame(0)
        g.initialize(f0, roi); o.initialize(f0, roi)
        cut = int(rng.integers(10, nf)) if rng.random() < 0.5 else -1
        for t in range(1, nf + 1):
            f = seq.frame(t)
            if cut > 0 and cut <= t < cut + 12:
                f = np.roll(255 - f, 17, axis=1)                  # scene cut: frame-level reset on the device
            lr = (16.0 if algo == "lobster" else (1.0 if t <= 8 else 0.0)) if rng.random() > 0.05 else float(rng.choice([1.0, 2.0, 8.0]))
            mg, mo = g.apply(f, lr), o.apply(f, lr)
            assert np.array_equal(mg, mo), f"frame {t}: masks differ in {(mg != mo).sum()} px"
            r = rng.random()
            if r < 0.04:
                TP._compare_state(g, o, ints, flts, f"frame {t}")
            elif r < 0.07:
                assert np.array_equal(g.getBackgroundImage(), o.get_background_image()), f"frame {t}: background image"
            elif r < 0.09:
                frac, force = float(rng.choice([0.1, 0.5, 1.0])), bool(rng.random() < 0.5)
                g.refreshModel(frac, force); o.refresh_model(frac, force)
            elif r < 0.11:
                for n in ("bg_color", "bg_desc"):              # re-import of the sample model (rebuilds the colour boxes)
                    g.state_set(n, o.state_get(n))
        TP._compare_state(g, o, ints, flts, "end")
        print("ok  ", tag, flush=True)
    except Exception as e:     # noqa: BLE001
        bad += 1
        print("FAIL", tag, "->", str(e)[:300], flush=True)
print(f"{ncases - bad} of {ncases} cases bit-exact")
sys.exit(1 if bad else 0)
