"""GPU parity on REAL frames (run with -m gpu on the B200 box): the 320x240 crop of the reference's 1080p sample sequence
(tests/golden/tractor_crop.npz, see tests/test_tractor_golden_cpu.py). The CUDA path must equal the oracle's snapshot mode bit for bit, and its
last mask must stay within the tier-3 tolerance of the mask the REFERENCE'S OWN CODE produced on the same frames (stored in the fixture)."""
import os

import numpy as np
import pytest

import test_tractor_golden_cpu as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(T.ALGOS))
@pytest.mark.parametrize("gray", [False, True])
def test_cuda_path_on_real_frames(lv, oracle, name, gray):
    algo, lr = T.ALGOS[name]
    fr = T.frames_of(gray)
    cls = {"lobster": lv.BackgroundSubtractorLOBSTER, "subsense": lv.BackgroundSubtractorSuBSENSE, "pawcs": lv.BackgroundSubtractorPAWCS}[name]
    g, o = cls(seed=0), oracle.Oracle(algo, mode=oracle.MODE_SNAPSHOT, seed=0)
    g.initialize(fr[0]); o.initialize(fr[0])
    mg = None
    for t in range(1, len(fr)):
        rate = lr if lr is not None else (1.0 if t <= 8 else 0.0)
        mg, mo = g.apply(fr[t], rate), o.apply(fr[t], rate)
        assert np.array_equal(mg, mo), f"{name} frame {t}: CUDA mask differs from the oracle's in {(mg != mo).sum()} px"
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())
    if not gray:
        ref = T.G[f"{name}_rgb_last_mask"]
        d = float((mg != ref).mean())
        noise = max(float((T.run(algo, lr, fr, oracle.MODE_REFERENCE, seed=s)[1] != ref).mean()) for s in (2, 3))
        assert d <= T.TIER3_SEED_FACTOR * noise + T.TIER3_MARGIN, f"{name}: {100 * d:.2f} % vs the reference's mask (seed noise {100 * noise:.2f} %)"
