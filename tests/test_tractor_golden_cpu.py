"""Real data: 20 frames of a 320x240 crop of the reference's 1080p sample sequence (samples/data/tractor.mp4) with the
outputs of the REFERENCE'S OWN CODE on them (tests/golden/tractor_crop.npz, written by tests/golden/make_tractor_golden.py from oracle/_ref).
The oracle in reference-order mode must reproduce them bit for bit — this pin travels with the repository (the GPU box has no /root/reference) —
and the snapshot mode (what the CUDA path implements) must stay within the stated tier-3 disagreement of it."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tractor_crop.npz"))
ALGOS = {"lobster": (O.ALGO_LOBSTER, 16.0), "subsense": (O.ALGO_SUBSENSE, None), "pawcs": (O.ALGO_PAWCS, 0.0)}
# tier 3 (north star: "within a stated per-pixel disagreement"): the reference itself is only reproducible up to its rand() seed, so the
# bound is derived, not guessed: snapshot-vs-reference disagreement <= 1.5 x the seed-to-seed disagreement of reference-order runs + 0.2 %
TIER3_SEED_FACTOR, TIER3_MARGIN = 1.5, 0.002


def frames_of(gray):
    fr = G["frames"]
    if not gray:
        return fr
    # cv2.cvtColor(BGR2GRAY) of the generating script, restated: (B*1868 + G*9617 + R*4899 + 8192) >> 14
    return ((fr[..., 0].astype(np.uint32) * 1868 + fr[..., 1].astype(np.uint32) * 9617 + fr[..., 2].astype(np.uint32) * 4899 + 8192) >> 14).astype(np.uint8)


def run(algo, lr, fr, mode, seed=0):
    a = O.Oracle(algo, mode=mode, seed=seed)
    a.initialize(fr[0])
    h = hashlib.sha256()
    last = None
    for t in range(1, len(fr)):
        last = a.apply(fr[t], lr if lr is not None else (1.0 if t <= 8 else 0.0))
        h.update(last.tobytes())
    return h.digest(), last, a.get_background_image()


@pytest.mark.parametrize("name", sorted(ALGOS))
@pytest.mark.parametrize("gray", [False, True])
def test_oracle_reference_order_reproduces_the_reference_on_real_frames(name, gray):
    algo, lr = ALGOS[name]
    key = f"{name}_{'gray' if gray else 'rgb'}"
    digest, last, bg = run(algo, lr, frames_of(gray), O.MODE_REFERENCE)
    assert np.array_equal(last, G[key + "_last_mask"]), f"{key}: last mask differs from the reference's in {(last != G[key + '_last_mask']).sum()} px"
    assert digest == G[key + "_masks_sha256"].tobytes(), f"{key}: mask sequence differs from the reference's"
    assert np.array_equal(bg, G[key + "_bg"]), f"{key}: background image differs from the reference's"


@pytest.mark.parametrize("name", sorted(ALGOS))
def test_snapshot_semantics_stay_within_tier3_tolerance_of_the_reference(name):
    algo, lr = ALGOS[name]
    ref = G[f"{name}_rgb_last_mask"]
    _, last, _ = run(algo, lr, frames_of(False), O.MODE_SNAPSHOT)
    d = float((last != ref).mean())
    noise = max(float((run(algo, lr, frames_of(False), O.MODE_REFERENCE, seed=s)[1] != ref).mean()) for s in (2, 3))   # (glibc: srand(1) == srand(0))
    assert noise > 0, "seeds 2 and 3 reproduce seed 0 exactly: the sequence does not exercise the stochastic updates"
    assert d <= TIER3_SEED_FACTOR * noise + TIER3_MARGIN, \
        f"{name}: snapshot-mode mask disagrees with the reference's in {100 * d:.2f} % of the pixels; seed-to-seed noise of the reference order is {100 * noise:.2f} %"
