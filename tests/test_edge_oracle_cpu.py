"""CPU tests of the EdgeDetectorLBSP oracle (oracle/lvo_edge_lbsp.hpp; the CUDA counterpart is lvb_edge_*, tests/test_gpu_edge.py).
The reference has no test or golden vector for the detector (the restatement itself is pinned to the reference's own source by
tests/test_ref_pin_cpu.py); these tests pin the restatement's documented properties,
including the three observable quirks of the source listed in DESIGN.md (row shift, unwritten mask rows, little-endian initial value)."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence


def _square(n=64, lo=60, hi=200, a=20, b=44):
    img = np.full((n, n), lo, np.uint8)
    img[a:b, a:b] = hi
    return img


def test_flat_image_has_no_edges_and_square_has_a_closed_outline(oracle):
    O = oracle
    assert not O.EdgeDetectorLBSPOracle().apply_threshold(np.full((48, 64, 3), 90, np.uint8)).any()
    m = O.EdgeDetectorLBSPOracle(levels=1).apply_threshold(_square(), 0.3)
    assert set(np.unique(m)) == {0, 255}
    ys, xs = np.nonzero(m)
    assert xs.min() in (18, 19, 20) and xs.max() in (43, 44, 45)        # columns: around the square's sides
    # rows: the non-maximum-suppression loop writes gradient row r+2 into mask row r (EdgeDetectorLBSP.cpp:263, 270-272)
    assert ys.min() in (16, 17, 18) and ys.max() in (41, 42, 43)
    assert not m[30, 25:40].any()                                       # nothing inside the square


def test_gradient_map_keeps_the_little_endian_initial_value_quirk(oracle):
    """(CHAR_MAX<<24)|(CHAR_MAX<<16)|(UCHAR_MAX<<8) stored as uint32 -> per-pixel bytes (gradX, gradY, mag, pad) = (0, -1, 127, 127):
    with the min-|.| combination gradX stays 0 and |gradY| <= 1; the magnitude is the minimum over the scales"""
    O = oracle
    f = SynthSequence(160, 120, 3, seed=3).frame(10)
    e1 = O.EdgeDetectorLBSPOracle(levels=1)
    e1.apply_threshold(f)
    g = e1.gradient_map(f.shape)
    own = O.lbsp_gradient(f)
    assert not g[..., 0].any() and np.abs(g[..., 1].view(np.int8)).max() <= 1
    assert np.array_equal(g[..., 2], own[..., 2])                       # one scale: min(own, 127) = own
    gy_own = own[..., 1].view(np.int8).astype(int)
    want_gy = np.where(np.abs(gy_own) <= 1, gy_own, -1)                 # std::min(new, -1, |a| < |b|)
    assert np.array_equal(g[..., 1].view(np.int8).astype(int), want_gy)
    e3 = O.EdgeDetectorLBSPOracle(levels=3)
    e3.apply_threshold(f)
    assert (e3.gradient_map(f.shape)[..., 2] <= g[..., 2]).all()        # more scales can only lower the magnitude


def test_thresholds_are_monotone_and_confidence_map_counts_them(oracle):
    O = oracle
    f = SynthSequence(128, 96, 1, seed=8).frame(12)
    masks = [O.EdgeDetectorLBSPOracle().apply_threshold(f, t / 16.0) > 0 for t in range(16)]   # fresh objects
    for a, b in zip(masks[:-1], masks[1:]):
        assert not (b & ~a).any()                                       # a higher threshold never adds an edge pixel
    assert masks[2].any() and masks[2].sum() > masks[10].sum()
    conf = O.EdgeDetectorLBSPOracle().apply(f)                          # apply(): sixteen passes on ONE object, 16 per hit, saturated
    assert set(np.unique(conf)) <= set(list(range(0, 256, 16)) + [255])
    assert (conf > 0).any() and conf.max() <= 255
    assert np.array_equal(O.EdgeDetectorLBSPOracle().apply_threshold(f, -1.0), O.EdgeDetectorLBSPOracle().apply_threshold(f, 0.5))   # default


def test_normalized_output_restatement_equals_cv2_normalize(oracle):
    """bNormalizeOutput (EdgeDetectorLBSP.cpp:431-432): cv::normalize(x, x, 0, UCHAR_MAX, NORM_MINMAX). The confidence map only holds
    0, 16, .., 240, 255: every (min, max) pair of those values, array lengths that exercise OpenCV's vector body and its scalar tail,
    random maps over that value set, and the detector's own output, against cv2"""
    cv2 = pytest.importorskip("cv2")
    vals = [16 * i for i in range(16)] + [255]
    for lo in vals:
        for hi in vals:
            if hi < lo:
                continue
            inner = [v for v in vals if lo <= v <= hi]
            for n in (1, 3, 7, 16, 33, 64, 100):
                src = np.array((inner * (n // len(inner) + 1))[:n], np.uint8)
                src[0] = lo
                if n > 1:
                    src[-1] = hi
                src = src.reshape(1, -1)
                assert np.array_equal(oracle.normalize_minmax_u8(src), cv2.normalize(src, None, 0, 255, cv2.NORM_MINMAX)), (lo, hi, n)
    # (arbitrary 8-bit maps are out of scope: there OpenCV's result depends on whether its build fuses the multiply-add, e.g. 190 in a
    # map spanning 181..199 comes out as 127 or 128; on the detector's value set fused and unfused arithmetic agree everywhere)
    rng = np.random.default_rng(5)
    for _ in range(40):
        src = np.array(vals, np.uint8)[rng.integers(0, 17, size=(rng.integers(1, 40), rng.integers(1, 70)))]
        assert np.array_equal(oracle.normalize_minmax_u8(src), cv2.normalize(src, None, 0, 255, cv2.NORM_MINMAX))
    img = _square()
    plain, norm = oracle.EdgeDetectorLBSPOracle().apply(img), oracle.EdgeDetectorLBSPOracle(normalize_output=True).apply(img)
    assert np.array_equal(norm, cv2.normalize(plain, None, 0, 255, cv2.NORM_MINMAX)) and norm.max() == 255


def test_two_and_four_channel_images(oracle):
    """the reference instantiates the detector for 1 to 4 channels (EdgeDetectorLBSP.cpp:144-160); the gradient takes the channel with the
    largest response (LBSP.hpp:235-256), so replicating a gray image into 2 or 4 channels must not change the result"""
    img = _square()
    want = oracle.EdgeDetectorLBSPOracle().apply_threshold(img, 0.5)
    for c in (2, 4):
        assert np.array_equal(oracle.EdgeDetectorLBSPOracle().apply_threshold(np.repeat(img[..., None], c, axis=2), 0.5), want)
    with pytest.raises(Exception):
        oracle.EdgeDetectorLBSPOracle().apply_threshold(np.repeat(img[..., None], 5, axis=2), 0.5)


def test_repeatable_on_one_object_and_errors(oracle):
    O = oracle
    f = SynthSequence(96, 80, 3, seed=1).frame(7)
    e = O.EdgeDetectorLBSPOracle()
    a = e.apply_threshold(f, 0.4)
    assert np.array_equal(a, e.apply_threshold(f, 0.4))                 # the buffers persist between calls, the same input reproduces
    assert np.array_equal(a, O.EdgeDetectorLBSPOracle().apply_threshold(f, 0.4))
    with pytest.raises(O.OracleError, match="too small"):
        O.EdgeDetectorLBSPOracle(levels=3).apply_threshold(np.zeros((12, 40), np.uint8))       # 12 -> 6 -> 3 rows
    with pytest.raises(O.OracleError):
        O.EdgeDetectorLBSPOracle(levels=0)
    with pytest.raises(O.OracleError):
        O.EdgeDetectorLBSPOracle(hyst_low_factor=1.0)
