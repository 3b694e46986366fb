"""Generates tests/golden/sequence_golden.npz: the oracle's snapshot-mode results on small integer-generated sequences, committed as
regression fixtures (the reference holds no golden vector for apply(), so the parity anchor is our restatement; these
fixtures keep that anchor from drifting unnoticed when oracle and kernels are edited together).

Run from the repo root: python tests/golden/make_sequence_golden.py . The input frames use integer arithmetic only (PCG64 integers, box
sums, a moving rectangle), so they are identical on every platform; each case stores the SHA-256 of all masks of the run, the last mask
and the SHA-256 of the final sample model / word planes.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
W, H, NF = 96, 72, 36


def frames(channels, seed):
    """textured static background (integer box-filtered noise), one dark rectangle moving 2 px / frame, +-3 integer noise"""
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (H, W, channels)
    base = rng.integers(0, 256, (H + 4, W + 4, channels), dtype=np.int64)
    acc = np.zeros(shape, np.int64)
    for dy in range(5):
        for dx in range(5):
            acc += base[dy:dy + H, dx:dx + W]
    bg = (acc // 25).astype(np.int64)
    bg = (bg - 128) * 3 + 128 + (np.arange(W, dtype=np.int64)[None, :, None] * 64) // W     # more contrast + a horizontal ramp
    out = []
    for t in range(NF):
        f = bg + rng.integers(-3, 4, shape, dtype=np.int64)
        if t > 0:
            x0, y0 = (5 + 2 * t) % (W - 20), 20 + (t % 7)
            f[y0:y0 + 18, x0:x0 + 14] = 24 + 12 * np.arange(channels, dtype=np.int64)
        f = np.clip(f, 0, 255).astype(np.uint8)
        out.append(f[..., 0] if channels == 1 else f)
    return out


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_cases():
    from oracle import oracle as O
    res = {}

    def lbsp_family(name, algo, ch, lr_boot, lr, model):
        o = O.Oracle(algo, mode=O.MODE_SNAPSHOT, seed=3)
        fr = frames(ch, 100 + ch)
        o.initialize(fr[0])
        masks = [o.apply(f, lr_boot if t <= 5 else lr) for t, f in enumerate(fr[1:], start=1)]
        res[name] = (sha(*masks), masks[-1], sha(*[o.state_get(n) for n in model]))

    lbsp_family("subsense_3ch", O.ALGO_SUBSENSE, 3, 1.0, 0.0, ["bg_color", "bg_desc", "T", "R", "v", "Dlast"])
    lbsp_family("subsense_1ch", O.ALGO_SUBSENSE, 1, 1.0, 0.0, ["bg_color", "bg_desc", "T", "R", "v", "Dlast"])
    lbsp_family("lobster_1ch", O.ALGO_LOBSTER, 1, 16.0, 16.0, ["bg_color", "bg_desc"])
    lbsp_family("lobster_3ch", O.ALGO_LOBSTER, 3, 16.0, 16.0, ["bg_color", "bg_desc"])
    lbsp_family("pawcs_3ch", O.ALGO_PAWCS, 3, 1.0, 0.0, ["lw_first", "lw_last", "lw_occ", "lw_color", "lw_desc", "T", "R", "v"])
    for ch in (1, 3):
        fr = frames(ch, 100 + ch)
        v = O.ViBeOracle(ch, mode=O.MODE_SNAPSHOT, seed=3)
        v.initialize(fr[0])
        masks = [v.apply(f, 16.0) for f in fr[1:]]
        res[f"vibe_{ch}ch"] = (sha(*masks), masks[-1], sha(v.model()))
        p = O.PBASOracle(ch, mode=O.MODE_SNAPSHOT, seed=3)
        p.initialize(fr[0])
        masks = [p.apply(f) for f in fr[1:]]
        res[f"pbas_{ch}ch"] = (sha(*masks), masks[-1], sha(*[p.state_get(n) for n in ("bg_color", "bg_grad", "R", "T", "meanmin", "scalars")]))
        e = O.EdgeDetectorLBSPOracle()
        edges = [e.apply_threshold(fr[5], t / 16.0) for t in (4, 8, 12)] + [O.EdgeDetectorLBSPOracle().apply(fr[5])]
        res[f"edge_lbsp_{ch}ch"] = (sha(*edges), edges[1], sha(O.lbsp_gradient(fr[5])))
    return res


def oracle_factories():
    from oracle import oracle as O
    return {"subsense": lambda ch: O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_SNAPSHOT, seed=3),
            "lobster": lambda ch: O.Oracle(O.ALGO_LOBSTER, mode=O.MODE_SNAPSHOT, seed=3),
            "pawcs": lambda ch: O.Oracle(O.ALGO_PAWCS, mode=O.MODE_SNAPSHOT, seed=3),
            "vibe": lambda ch: O.ViBeOracle(ch, mode=O.MODE_SNAPSHOT, seed=3),
            "pbas": lambda ch: O.PBASOracle(ch, mode=O.MODE_SNAPSHOT, seed=3),
            "edge": lambda: O.EdgeDetectorLBSPOracle(), "lbsp_gradient": O.lbsp_gradient}


def run_mask_cases(make):
    """the protocol of run_cases() for ANY implementation with the reference's method names (make: kind -> factory, see
    oracle_factories): name -> (SHA-256 of all masks, last mask, SHA-256 of the LBSP gradient map for the edge cases, else None).
    tests/test_gpu_parity.py::test_gpu_reproduces_the_committed_sequence_fixtures runs it with the CUDA classes."""
    res = {}
    for name, kind, ch, lr_boot, lr in [("subsense_3ch", "subsense", 3, 1.0, 0.0), ("subsense_1ch", "subsense", 1, 1.0, 0.0), ("lobster_1ch", "lobster", 1, 16.0, 16.0),
                                        ("lobster_3ch", "lobster", 3, 16.0, 16.0), ("pawcs_3ch", "pawcs", 3, 1.0, 0.0)]:
        a, fr = make[kind](ch), frames(ch, 100 + ch)
        a.initialize(fr[0])
        masks = [a.apply(f, lr_boot if t <= 5 else lr).copy() for t, f in enumerate(fr[1:], start=1)]
        res[name] = (sha(*masks), masks[-1], None)
    for ch in (1, 3):
        fr = frames(ch, 100 + ch)
        v = make["vibe"](ch)
        v.initialize(fr[0])
        masks = [v.apply(f, 16.0).copy() for f in fr[1:]]
        res[f"vibe_{ch}ch"] = (sha(*masks), masks[-1], None)
        p = make["pbas"](ch)
        p.initialize(fr[0])
        masks = [p.apply(f).copy() for f in fr[1:]]
        res[f"pbas_{ch}ch"] = (sha(*masks), masks[-1], None)
        e = make["edge"]()
        edges = [e.apply_threshold(fr[5], t / 16.0).copy() for t in (4, 8, 12)] + [make["edge"]().apply(fr[5])]
        res[f"edge_lbsp_{ch}ch"] = (sha(*edges), edges[1], sha(make["lbsp_gradient"](fr[5])))
    return res


if __name__ == "__main__":
    res = run_cases()
    out = {}
    for k, (m, last, s) in res.items():
        out[k + "__masks_sha256"] = np.array(m)
        out[k + "__last_mask"] = last
        out[k + "__state_sha256"] = np.array(s)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sequence_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(res), "cases")
    for k, (m, last, s) in res.items():
        print(f"  {k:16s} fg {float((last > 0).mean()):.4f}  masks {m[:12]}  state {s[:12]}")
