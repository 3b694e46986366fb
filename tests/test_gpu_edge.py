"""GPU parity of the LBSP edge detector (lvb_edge_*, litiv_b200/csrc/edge.cuh) against the CPU oracle, through the
C ABI: edge masks, gradient maps and confidence maps bit-exact for 1 / 3 pyramid levels, gray / RGB, odd and even sizes up to 1080p,
sequences of calls on one object (its maps persist, like the reference's). First run on a B200: 18 passed (gpurun_out/edge_gpu_tests.log,
profiles/r01z_edge_probe.md)."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _frame(seq, t, ch):
    f = np.ascontiguousarray(seq.frame(t))
    if ch == 1 and f.ndim == 3:
        return np.ascontiguousarray(f[..., 0])
    if ch in (2, 4):   # the reference instantiates the detector for 1 to 4 channels (EdgeDetectorLBSP.cpp:144-160)
        extra = (f[..., :1].astype(np.int32) * 3 + f[..., 1:2] * 5 + 17 * t) % 256
        return np.ascontiguousarray(np.concatenate([f, extra.astype(np.uint8)], axis=2)[..., :ch] if ch == 4 else f[..., :2])
    return f


@pytest.mark.parametrize("size", [(96, 72), (97, 73), (320, 240), (641, 479)])
@pytest.mark.parametrize("ch", [1, 2, 3, 4])
@pytest.mark.parametrize("levels", [1, 3])
def test_edge_masks_and_gradient_map_match_oracle(lv, oracle, size, ch, levels):
    w, h = size
    seq = SynthSequence(w, h, 1 if ch == 1 else 3, seed=w + h + ch)
    o, e = oracle.EdgeDetectorLBSPOracle(levels=levels), lv.EdgeDetectorLBSP(levels)
    for t, thr in [(3, 0.5), (5, 0.25), (7, 0.75), (9, 0.0), (12, -1.0)]:
        f = _frame(seq, t, ch)
        want, got = o.apply_threshold(f, thr), e.apply_threshold(f, thr)
        assert np.array_equal(e.gradient_map(), o.gradient_map(f.shape)), (t, thr)
        assert np.array_equal(got, want), (t, thr, int((got != want).sum()))
    f = _frame(seq, 14, ch)
    want = o.apply(f)
    assert np.array_equal(e.apply(f), want) and want.any()


def test_edge_1080p_and_size_change(lv, oracle):
    e, o = lv.EdgeDetectorLBSP(), oracle.EdgeDetectorLBSPOracle()
    f = SynthSequence(1920, 1080, 3, seed=4).frame(20)
    assert np.array_equal(e.apply_threshold(f), o.apply_threshold(f))
    assert e.flood_sweeps() == 0     # hysteresis by union-find on row runs: no relaxation sweeps, no host round trip
    small = SynthSequence(160, 120, 3, seed=4).frame(20)
    assert np.array_equal(e.apply_threshold(small, 0.3), oracle.EdgeDetectorLBSPOracle().apply_threshold(small, 0.3))


@pytest.mark.parametrize("size,ch", [((160, 120), 3), ((97, 73), 1), ((320, 240), 4)])
def test_edge_normalized_confidence_map(lv, oracle, size, ch):
    """bNormalizeOutput=true (EdgeDetectorLBSP.cpp:431-432): cv::normalize(NORM_MINMAX) of the confidence map; the oracle's restatement of
    it is pinned against cv2 on the CPU (tests/test_edge_oracle_cpu.py)"""
    w, h = size
    seq = SynthSequence(w, h, 1 if ch == 1 else 3, seed=31)
    e, o = lv.EdgeDetectorLBSP(3, 0.5, True), oracle.EdgeDetectorLBSPOracle(normalize_output=True)
    plain = oracle.EdgeDetectorLBSPOracle()
    for t in (4, 9):
        f = _frame(seq, t, ch)
        want, got = o.apply(f), e.apply(f)
        assert np.array_equal(got, want), int((got != want).sum())
        assert want.max() == 255 and not np.array_equal(want, plain.apply(f))
    # a map whose minimum is not zero (every pixel an edge at some threshold) cannot come out of real images easily: a flat image instead
    flat = np.full((h, w) if ch == 1 else (h, w, ch), 90, np.uint8)
    assert np.array_equal(e.apply(flat), o.apply(flat)) and not o.apply(flat).any()


def test_edge_hysteresis_by_sweeps_equals_union_find(lv, oracle, monkeypatch):
    """LVB_EDGE_SWEEPS=1 (read when the detector is created) selects round 1's relaxation sweeps; both forms of the hysteresis must give
    the oracle's masks on images with long thin edges, blobs and noise"""
    rng = np.random.default_rng(8)
    img = np.full((241, 333), 90, np.uint8)
    img[40:200, 50:53] = 200; img[100:103, 20:300] = 30                         # long thin structures
    img[150:230, 200:320] = rng.integers(0, 256, (80, 120), dtype=np.int64).astype(np.uint8)   # noise block: many small components
    yy, xx = np.mgrid[0:241, 0:333]
    img[(yy - 60) ** 2 + (xx - 250) ** 2 < 900] = 170                            # disc
    uf = lv.EdgeDetectorLBSP(3)
    monkeypatch.setenv("LVB_EDGE_SWEEPS", "1")
    sw = lv.EdgeDetectorLBSP(3)
    monkeypatch.delenv("LVB_EDGE_SWEEPS")
    o = oracle.EdgeDetectorLBSPOracle(3)
    for thr in (0.5, 0.2, 0.8, 0.05):
        want = o.apply_threshold(img, thr)
        assert np.array_equal(uf.apply_threshold(img, thr), want), thr
        assert np.array_equal(sw.apply_threshold(img, thr), want), thr
    assert uf.flood_sweeps() == 0 and sw.flood_sweeps() >= 4 and want.any()


def test_edge_argument_checks(lv):
    with pytest.raises(lv.LitivError):
        lv.EdgeDetectorLBSP(0)
    with pytest.raises(lv.LitivError):
        lv.EdgeDetectorLBSP(3, 1.0)
    with pytest.raises(lv.LitivError):
        lv.EdgeDetectorLBSP(3).apply_threshold(np.zeros((16, 16), np.uint8))   # 16 -> 8 -> 4: smaller than the LBSP patch
