"""GPU parity tests for PAWCS (run with -m gpu on the B200 box): the CUDA path through the C ABI against the CPU oracle
(oracle/lvo_pawcs.hpp, snapshot mode, same Philox seed) on identical inputs. Every integer/byte buffer (local word
dictionaries, global dictionary, per-pixel sort LUTs, masks) must be bit-exact after every frame; float maps within 1e-5
relative (we observe bit-identical)."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-5

INT_STATE = ["lastfg", "lastcolor", "lastdesc", "lut", "unstable", "illum", "blinks", "lastraw", "lastrawblink", "dil", "dilinv", "rawmask",
             "lw_first", "lw_last", "lw_occ", "lw_color", "lw_desc", "gw_bits", "gw_color", "gw_desc", "gdict", "glut"]
FLT_STATE = ["T", "R", "v", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST", "gw_weight", "gw_map"]


def _mk(lv, oracle, seed):
    return lv.BackgroundSubtractorPAWCS(seed=seed), oracle.Oracle(oracle.ALGO_PAWCS, mode=oracle.MODE_SNAPSHOT, seed=seed)


def _compare(g, o, tag, skip=()):
    roi = o.state_get("roi") > 0
    for n in INT_STATE:
        if n in skip:
            continue
        a, b = g.state_get(n), o.state_get(n)
        if n.startswith("lw_") or n == "glut":     # per-pixel dictionaries only exist inside the ROI
            k = a.size // roi.size
            a, b = a.reshape(roi.size, k)[roi], b.reshape(roi.size, k)[roi]
        nbad = int((a != b).sum())
        assert nbad == 0, f"{tag}: integer state '{n}' differs in {nbad} of {a.size} entries (first at {np.argwhere(a != b)[:4].tolist()})"
    for n in FLT_STATE:
        a, b = g.state_get(n), o.state_get(n)
        assert np.allclose(a, b, rtol=FLOAT_RTOL, atol=1e-7), f"{tag}: float map '{n}' max abs diff {np.abs(a - b).max()} at {np.argmax(np.abs(a - b))}"
    sa, sb = g.state_get("scalars"), o.state_get("scalars")
    assert np.allclose(sa[:13], sb[:13], rtol=1e-6), f"{tag}: scalars differ {sa[:13]} vs {sb[:13]}"


@pytest.mark.parametrize("w,h,c,nframes,roi", [(160, 120, 3, 20, None), (160, 120, 1, 16, None), (320, 240, 3, 10, None), (96, 72, 3, 12, "roi"),
                                                  (330, 250, 3, 8, None), (75, 61, 1, 10, "roi")])   # sizes that are not multiples of 8 (general INTER_AREA)
def test_pawcs_state_parity(lv, oracle, w, h, c, nframes, roi):
    seq = SynthSequence(w, h, c, seed=13)
    roi_img = None
    if roi:
        roi_img = np.zeros((h, w), np.uint8)
        roi_img[h // 6:h - h // 8, w // 5:w - 3] = 255
        roi_img[h // 2:h // 2 + 5, w // 2:w // 2 + 9] = 0
    g, o = _mk(lv, oracle, seed=7)
    f0 = seq.frame(0)
    g.initialize(f0, roi_img)
    o.initialize(f0, roi_img)
    assert np.array_equal(g.getROICopy().ravel(), o.state_get("roi"))
    _compare(g, o, "init", skip=("rawmask",))
    for t in range(1, nframes + 1):
        f = seq.frame(t)
        mg, mo = g.apply(f, 0.0), o.apply(f, 0.0)
        _compare(g, o, f"frame {t}")
        assert np.array_equal(mg, mo), f"frame {t}: final masks differ in {(mg != mo).sum()} px"
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())
    assert np.array_equal(g.getBackgroundDescriptorsImage(), o.get_background_descriptors_image())


def test_pawcs_import_oracle_snapshot_and_refresh(lv, oracle):
    """evolve the ORACLE, import its whole state into the GPU object, then step both; then refreshModel on both"""
    w, h, c = 160, 120, 3
    seq = SynthSequence(w, h, c, seed=5)
    g, o = _mk(lv, oracle, seed=3)
    f0 = seq.frame(0)
    g.initialize(f0); o.initialize(f0)
    for t in range(1, 40):
        o.apply(seq.frame(t), 0.0)
    for n in INT_STATE + FLT_STATE + ["scalars"]:
        if n != "rawmask":
            g.state_set(n, o.state_get(n))
    for t in range(40, 46):
        f = seq.frame(t)
        mg, mo = g.apply(f, 0.0), o.apply(f, 0.0)
        _compare(g, o, f"after import, frame {t}")
        assert np.array_equal(mg, mo)
    g.refreshModel(125, 0.0, True); o.pawcs_refresh_model(125, 0.0, True)
    _compare(g, o, "refreshModel(125,0,true)")
    g.refreshModel(1, 1.0, False); o.pawcs_refresh_model(1, 1.0, False)
    _compare(g, o, "refreshModel(1,1,false)")
    f = seq.frame(46)
    assert np.array_equal(g.apply(f, 0.0), o.apply(f, 0.0))
    _compare(g, o, "frame after refresh")


@pytest.mark.parametrize("w,h", [(64, 48), (70, 51)])
def test_pawcs_learning_rate_override_and_long_run(lv, oracle, w, h):
    """learning-rate override (incl. +inf), the end of the bootstrap window (frame 500: model check, maintenance recalc at
    256/512) and a scene change that triggers the frame-level reset; masks compared every frame, full state at checkpoints.
    70x51: neither dimension divides by 8 nor by 2 (general INTER_AREA for the motion analysis, the ROI and the model check)"""
    c = 3
    seq = SynthSequence(w, h, c, seed=21)
    g, o = _mk(lv, oracle, seed=9)
    f0 = seq.frame(0)
    g.initialize(f0); o.initialize(f0)
    n = 530
    for t in range(1, n):
        f = seq.frame(t)
        if 300 <= t < 320:
            f = (255 - f)                      # abrupt scene change -> frame-level reset path
        lr = 2.0 if t < 10 else (float("inf") if 10 <= t < 14 else 0.0)
        mg, mo = g.apply(f, lr), o.apply(f, lr)
        assert np.array_equal(mg, mo), f"frame {t}: masks differ in {(mg != mo).sum()} px"
        if t in (9, 13, 64, 256, 299, 301, 307, 312, 321, 499, 500, 501, 512, n - 1):
            _compare(g, o, f"long run frame {t}")
    assert o.state_get("scalars")[12] >= 2, "the scene change was expected to trigger a model reset in the oracle"


def test_pawcs_large_counters_take_the_division_path(lv, oracle):
    """the bubble-pass kernel compares word weights by exact 64-bit cross-multiplication while every counter is below 2^24
    (exact in float) and redoes a pixel with the reference's float divisions otherwise, near ties included: occurrence counts
    around 2^24..2^26 (rounded by the int -> float conversion) and equal-ratio words must sort exactly like the oracle"""
    w, h, c = 96, 72, 3
    seq = SynthSequence(w, h, c, seed=17)
    g, o = _mk(lv, oracle, seed=11)
    f0 = seq.frame(0)
    g.initialize(f0); o.initialize(f0)
    for t in range(1, 8):
        o.apply(seq.frame(t), 0.0)
    occ = o.state_get("lw_occ").copy().reshape(h, w, -1)
    rng = np.random.default_rng(3)
    big = rng.integers(1 << 24, 1 << 26, size=occ.shape, dtype=np.int64).astype(np.uint32)
    occ[:, : w // 3] = big[:, : w // 3]                               # >= 2^24: float conversion rounds, division path
    occ[:, w // 3: 2 * w // 3] = ((occ[:, w // 3: 2 * w // 3].astype(np.int64) + 1) * 4099).astype(np.uint32)   # large products, many near ties
    o.state_set("lw_occ", occ.ravel())
    for n in INT_STATE + FLT_STATE + ["scalars"]:
        if n != "rawmask":
            g.state_set(n, o.state_get(n))
    for t in range(8, 20):
        f = seq.frame(t)
        mg, mo = g.apply(f, 0.0), o.apply(f, 0.0)
        _compare(g, o, f"large counters, frame {t}")
        assert np.array_equal(mg, mo)


def test_pawcs_api_errors(lv):
    with pytest.raises(lv.LitivError, match="56 local words"):
        lv.BackgroundSubtractorPAWCS(nMaxNbWords=60).initialize(np.zeros((48, 64, 3), np.uint8))
    g = lv.BackgroundSubtractorPAWCS()
    with pytest.raises(lv.LitivError, match="initialized"):
        g.apply(np.zeros((48, 64, 3), np.uint8))
    g.initialize(np.zeros((50, 70, 3), np.uint8))   # not a multiple of 8: general INTER_AREA path
    assert g.apply(np.zeros((50, 70, 3), np.uint8)).shape == (50, 70)
    g.initialize(np.zeros((48, 64, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="fraction"):
        g.refreshModel(1, 1.5)
    assert g.getDefaultLearningRate() == 0
