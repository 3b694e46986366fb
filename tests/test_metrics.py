"""CDnet-style evaluation: lv::BinClassif::accumulate + BinClassifMetrics
(reference modules/datasets/src/metrics.cpp:21-61, datasets/include/litiv/datasets/metrics.hpp:23-67, 213-257).
CPU part: the oracle restatement against an independent numpy statement of the rule and hand-computed known answers.
GPU part: the device kernel against the oracle, standalone and on a subtractor's latest mask."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

LABELS = np.array([0, 50, 85, 170, 255, 17, 254], np.uint8)  # the five CDnet labels + two values outside the protocol


def _random_case(rng, h, w):
    classif = rng.choice(np.array([0, 255, 128, 1], np.uint8), size=(h, w), p=[0.55, 0.4, 0.03, 0.02])  # anything but 255 is negative
    gt = rng.choice(LABELS, size=(h, w))
    roi = rng.choice(np.array([0, 255, 128], np.uint8), size=(h, w), p=[0.2, 0.7, 0.1])
    return classif, gt, roi


def _numpy_rule(classif, gt, roi):
    scored = (gt != 85) & (gt != 170)
    if roi is not None:
        scored &= roi != 0
    pos, gpos = classif == 255, gt == 255
    return np.array([(scored & pos & gpos).sum(), (scored & ~pos & ~gpos).sum(), (scored & pos & ~gpos).sum(), (scored & ~pos & gpos).sum(),
                     (scored & pos & (gt == 50)).sum(), (~scored).sum()], np.uint64)


def test_oracle_binclassif_matches_rule_and_known_answers():
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    for h, w in [(1, 1), (7, 33), (240, 320)]:
        c, g, r = _random_case(rng, h, w)
        assert np.array_equal(O.binclassif(c, g, r), _numpy_rule(c, g, r))
        assert np.array_equal(O.binclassif(c, g, None), _numpy_rule(c, g, None))
        assert np.array_equal(O.binclassif(c, None, r), np.array([0, 0, 0, 0, 0, h * w], np.uint64))   # metrics.cpp:26-29
    # accumulation over calls (BinClassif::accumulate adds to the counters)
    c, g, r = _random_case(rng, 5, 9)
    once = O.binclassif(c, g, r)
    assert np.array_equal(O.binclassif(c, g, r, counters=once), 2 * once)
    # hand-computed: one pixel of every (input, gt) combination
    c = np.array([[255, 255, 255, 255, 255, 0, 0, 0, 0, 0]], np.uint8)
    g = np.array([[255, 0, 50, 85, 170, 255, 0, 50, 85, 170]], np.uint8)
    assert O.binclassif(c, g).tolist() == [1, 2, 2, 1, 1, 4]   # TP TN FP FN SE DC
    m = O.binclassif_metrics(np.array([6, 80, 4, 10, 0, 0], np.uint64))
    assert m["dRecall"] == 6 / 16 and m["dPrecision"] == 6 / 10 and m["dSpecificity"] == 80 / 84
    assert m["dPBC"] == 100.0 * 14 / 100 and abs(m["dFMeasure"] - 2 * (6 / 16 * 0.6) / (6 / 16 + 0.6)) < 1e-15
    assert abs(m["dMCC"] - (6 * 80 - 4 * 10) / np.sqrt(10 * 16 * 84 * 90)) < 1e-15
    assert O.binclassif_metrics(np.zeros(6, np.uint64)) == dict(dRecall=0.0, dSpecificity=0.0, dFPR=0.0, dFNR=0.0, dPBC=0.0, dPrecision=0.0, dFMeasure=0.0, dMCC=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (5, 31), (61, 75), (240, 320), (1080, 1920)])
def test_gpu_binclassif_matches_oracle(lv, oracle, shape):
    rng = np.random.default_rng(11)
    c, g, r = _random_case(rng, *shape)
    for gt, roi in [(g, r), (g, None), (None, r)]:
        b = lv.BinClassif().accumulate(c, gt, roi)
        want = oracle.binclassif(c, gt, roi)
        assert np.array_equal(b.counters, want), f"{shape}: {b.counters} vs {want}"
    b.accumulate(c, g, r)   # adds to the counters
    assert np.array_equal(b.counters, oracle.binclassif(c, g, r, counters=want))
    m, mo = b.metrics(), oracle.binclassif_metrics(b.counters)
    assert all(abs(m[k] - mo[k]) <= 1e-15 for k in mo), (m, mo)
    assert b.total() == int(b.counters[:4].sum()) and b.total(True) == b.total() + b.nDC
    with pytest.raises(lv.LitivError):
        lv.BinClassif().accumulate(c, g[:, :-1] if g.shape[1] > 1 else np.zeros((3, 3), np.uint8))


@pytest.mark.gpu
def test_gpu_binclassif_on_subtractor_mask(lv, oracle):
    """score the instance's latest foreground mask where it lives (no read-back) against the synthetic ground truth"""
    w, h = 320, 240
    seq = SynthSequence(w, h, 3, seed=4)
    g = lv.BackgroundSubtractorSuBSENSE(seed=2)
    g.initialize(seq.frame(0))
    dev, host = lv.BinClassif(), np.zeros(6, np.uint64)
    roi = np.full((h, w), 255, np.uint8); roi[:20] = 0
    for t in range(1, 40):
        f, fg = seq.frame(t, with_gt=True)
        mask = g.apply(f, 1.0 if t <= 20 else 0.0)
        gt = np.where(fg, 255, 0).astype(np.uint8)
        gt[-5:] = 85; gt[:, :3] = 170; gt[100:110, 100:110] = 50
        dev.accumulate(g, gt, roi)
        host = oracle.binclassif(mask, gt, roi, counters=host)
    assert np.array_equal(dev.counters, host)
    assert dev.nTP > 0 and dev.nTN > 0
