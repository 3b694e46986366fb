"""Pins the oracle to the REFERENCE ITSELF (CPU only, no GPU).

oracle/_ref/liblitiv_ref.so is built from the reference's own, unmodified sources (modules/video/src/BackgroundSubtractorSuBSENSE.cpp,
...LOBSTER.cpp, ...PAWCS.cpp, ...LBSP.cpp, BackgroundSubtractionUtils.cpp, modules/features2d/src/LBSP.cpp and the litiv/utils headers),
compiled where they lie under /root/reference against oracle/cvcompat (recipe: oracle/Makefile target `_ref`). The oracle's
reference-order mode (same raster order, a clone of glibc rand()) must reproduce it BIT FOR BIT: masks, every model sample, every
float map, the LUT and the frame-level scalars, from initialize() through apply(), refreshModel() and getBackgroundImage().
With this, `SuBSENSE / LOBSTER / PAWCS::apply` are no longer "parity unpinned": GPU == oracle(snapshot) is tested on the device,
oracle(reference order) == reference source is tested here, and the two oracle modes share every per-pixel function.
The same holds for ViBe, PBAS and EdgeDetectorLBSP (BackgroundSubtractorViBe.cpp, BackgroundSubtractorPBAS.cpp and
imgproc/src/EdgeDetectorLBSP.cpp are part of the same library).
"""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="neither oracle/_ref/liblitiv_ref.so nor the reference tree is present")

SUB = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "T", "R", "v", "Dlast", "DminLT", "DminST", "rawLT", "rawST",
       "finLT", "finST", "dsLT", "dsST", "unstable", "blinks", "lastraw", "lastrawblink", "dilinv"]
LOB = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc"]
PAW_MAPS = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "T", "R", "v", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST",
            "unstable", "illum", "blinks", "lastraw", "lastrawblink", "dil", "dilinv"]


def _same(a, b, tag):
    assert a.shape == b.shape, f"{tag}: shapes {a.shape} vs {b.shape}"
    assert np.array_equal(a, b), f"{tag}: {int((a != b).sum())} of {a.size} entries differ (first at {np.flatnonzero(a != b)[:5]})"


def _compare(r, o, names, tag, scalar_idx):
    for n in names:
        _same(r.state_get(n), o.state_get(n), f"{tag}: '{n}'")   # bit-exact, floats included
    sa, sb = r.state_get("scalars"), o.state_get("scalars")
    assert np.array_equal(sa[scalar_idx], sb[scalar_idx]), f"{tag}: scalars {sa[:12]} vs {sb[:12]}"


def _compare_pawcs(r, o, tag):
    _compare(r, o, PAW_MAPS, tag, list(range(12)))
    valid = r.state_get("lw_valid") > 0
    for n in ["lw_first", "lw_last", "lw_occ", "lw_color", "lw_desc"]:      # local dictionaries, dictionary order, where a word exists
        a, b = r.state_get(n), o.state_get(n)
        k = a.size // valid.size
        _same(a.reshape(valid.size, k)[valid], b.reshape(valid.size, k)[valid], f"{tag}: '{n}'")
    gd = o.state_get("gdict")                                                # oracle: word identity per dictionary position
    ng = gd.size
    for n in ["gw_weight", "gw_bits", "gw_color", "gw_desc", "gw_map"]:      # reference side is exported by dictionary position
        a = r.state_get(n).reshape(ng, -1)
        _same(a, o.state_get(n).reshape(-1, a.shape[1])[gd], f"{tag}: '{n}'")
    roi = o.state_get("roi") > 0
    pos = np.empty(ng, np.int64)
    pos[gd] = np.arange(ng)
    _same(r.state_get("glut").reshape(-1, ng)[roi].astype(np.int64), pos[o.state_get("glut").reshape(-1, ng)[roi]], f"{tag}: per-pixel global sort LUT")


def _roi(w, h):
    roi = np.zeros((h, w), np.uint8)
    roi[h // 6:h - h // 8, w // 5:w - 3] = 255
    roi[h // 2:h // 2 + 5, w // 2:w // 2 + 9] = 0
    return roi


@pytest.mark.parametrize("w,h,c,n,roi,seed", [
    (320, 240, 3, 64, False, 0),    # BASELINE config #1 shape, through the lr switch of samples/changedet (override 1 up to frame 50)
    (320, 240, 1, 20, False, 1),
    (200, 150, 3, 14, True, 2),     # user ROI, width not a multiple of 8 or 32
    (330, 250, 3, 10, False, 3),    # general INTER_AREA path of the motion analysis
    (96, 72, 3, 12, False, 4),      # "small" branch: no frame-level analysis, other T caps
    (640, 480, 3, 6, False, 5),     # config #5 per-stream shape: 5x5 spread, median 13
])
def test_subsense_oracle_reference_order_equals_reference_source(w, h, c, n, roi, seed):
    seq = SynthSequence(w, h, c, seed=10 + seed)
    r, o = R.Reference(O.ALGO_SUBSENSE, seed=seed), O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_REFERENCE, seed=seed)
    roi_img = _roi(w, h) if roi else None
    f0 = seq.frame(0)
    r.initialize(f0, roi_img)
    o.initialize(f0, roi_img)
    _same(r.get_roi().ravel(), o.state_get("roi"), "ROI after initialize")
    _compare(r, o, SUB, "init", list(range(12)))
    for t in range(1, n):
        f = seq.frame(t)
        lr = 1.0 if t <= (50 if n > 55 else n // 2) else 0.0
        _same(r.apply(f, lr), o.apply(f, lr), f"frame {t}: final mask")
        if t in (1, 2, n // 2, n - 1):
            _compare(r, o, SUB, f"frame {t}", list(range(12)))
    _same(r.get_background_image(), o.get_background_image(), "getBackgroundImage")
    _same(r.get_background_descriptors_image(), o.get_background_descriptors_image(), "getBackgroundDescriptorsImage")
    r.refresh_model(0.5); o.refresh_model(0.5)
    _compare(r, o, SUB, "refreshModel(0.5)", list(range(12)))
    r.refresh_model(1.0, True); o.refresh_model(1.0, True)
    f = seq.frame(n)
    _same(r.apply(f, 0.0), o.apply(f, 0.0), "frame after refreshModel")
    _compare(r, o, SUB, "after refreshModel + frame", list(range(12)))


def test_subsense_scene_change_reset_equals_reference_source():
    """the frame-level reset (SuBSENSE.cpp:584-600): refreshModel(0.1) fired from inside apply(), T(x) set to 1, cooldown, shrinking caps"""
    w, h, c = 320, 240, 3   # the frame-level analysis only runs on frames of at least 320x240 (SuBSENSE.cpp:112-128)
    seq_a, seq_b = SynthSequence(w, h, c, seed=21), SynthSequence(w, h, c, seed=22)
    r, o = R.Reference(O.ALGO_SUBSENSE, seed=5), O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_REFERENCE, seed=5)
    f0 = seq_a.frame(0)
    r.initialize(f0); o.initialize(f0)
    caps = set()
    for t in range(1, 80):
        f = seq_a.frame(t) if t < 40 else (seq_b.frame(t) // 5)
        lr = 1.0 if t <= 10 else 0.0
        _same(r.apply(f, lr), o.apply(f, lr), f"frame {t}: final mask")
        sc = r.state_get("scalars")
        caps.add((sc[7], sc[8]))
        if t in (39, 41, 45, 52, 60, 79):
            _compare(r, o, SUB, f"frame {t}", list(range(12)))
    assert len(caps) > 2, "the learning-rate caps never moved: the sequence does not exercise the frame-level analysis"
    assert r.state_get("scalars")[2] > 0 or r.state_get("scalars")[1] < 50, "the reset never fired in the reference"


@pytest.mark.parametrize("w,h,c,n,roi", [(320, 240, 1, 40, False), (320, 240, 3, 16, False), (75, 61, 1, 10, True)])
def test_lobster_oracle_reference_order_equals_reference_source(w, h, c, n, roi):
    seq = SynthSequence(w, h, c, seed=2)
    r, o = R.Reference(O.ALGO_LOBSTER, seed=3), O.Oracle(O.ALGO_LOBSTER, mode=O.MODE_REFERENCE, seed=3)
    roi_img = _roi(w, h) if roi else None
    f0 = seq.frame(0)
    r.initialize(f0, roi_img); o.initialize(f0, roi_img)
    idx = [3, 10, 11]   # the reference's LOBSTER keeps no frame counter (ours only indexes the Philox stream with it)
    _compare(r, o, LOB, "init", idx)
    for t in range(1, n):
        f = seq.frame(t)
        _same(r.apply(f, 16.0), o.apply(f, 16.0), f"frame {t}: final mask")
        if t in (1, n // 2, n - 1):
            _compare(r, o, LOB, f"frame {t}", idx)
    _same(r.get_background_image(), o.get_background_image(), "getBackgroundImage")
    _same(r.get_background_descriptors_image(), o.get_background_descriptors_image(), "getBackgroundDescriptorsImage")
    r.refresh_model(0.3); o.refresh_model(0.3)
    _compare(r, o, LOB, "refreshModel(0.3)", idx)


@pytest.mark.parametrize("w,h,c,n,roi", [(160, 120, 3, 40, False), (160, 120, 1, 24, False), (96, 72, 3, 20, True), (330, 250, 3, 5, False), (640, 480, 3, 3, False)])
def test_pawcs_oracle_reference_order_equals_reference_source(w, h, c, n, roi):
    """local word dictionaries, the global dictionary with its occupancy maps, the per-pixel sort LUTs, the maintenance frames (8, 16, 32)"""
    seq = SynthSequence(w, h, c, seed=3)
    r, o = R.Reference(O.ALGO_PAWCS, seed=7), O.Oracle(O.ALGO_PAWCS, mode=O.MODE_REFERENCE, seed=7)
    roi_img = _roi(w, h) if roi else None
    f0 = seq.frame(0)
    r.initialize(f0, roi_img); o.initialize(f0, roi_img)
    _compare_pawcs(r, o, "init")
    for t in range(1, n):
        f = seq.frame(t)
        _same(r.apply(f, 0.0), o.apply(f, 0.0), f"frame {t}: final mask")
        if t in (1, 8, 16, 32, n - 1):
            _compare_pawcs(r, o, f"frame {t}")
    _same(r.get_background_image(), o.get_background_image(), "getBackgroundImage")
    _same(r.get_background_descriptors_image(), o.get_background_descriptors_image(), "getBackgroundDescriptorsImage")
    if w <= 160:
        r.pawcs_refresh_model(125, 0.0, True); o.pawcs_refresh_model(125, 0.0, True)
        _compare_pawcs(r, o, "refreshModel(125,0,true)")


def test_pawcs_bootstrap_exit_and_reset_equals_reference_source():
    """past frame 500 (end of the bootstrap window: model check against the background image, maintenance recalculation at 256 / 512) and a
    scene change that triggers the frame-level reset, plus learning-rate overrides incl. +inf"""
    w, h, c = 64, 48, 3
    seq = SynthSequence(w, h, c, seed=21)
    r, o = R.Reference(O.ALGO_PAWCS, seed=9), O.Oracle(O.ALGO_PAWCS, mode=O.MODE_REFERENCE, seed=9)
    f0 = seq.frame(0)
    r.initialize(f0); o.initialize(f0)
    for t in range(1, 530):
        f = seq.frame(t)
        if 300 <= t < 320:
            f = 255 - f
        lr = 2.0 if t < 10 else (float("inf") if 10 <= t < 14 else 0.0)
        _same(r.apply(f, lr), o.apply(f, lr), f"frame {t}: final mask")
        if t in (9, 13, 256, 301, 312, 499, 500, 501, 512, 529):
            _compare_pawcs(r, o, f"frame {t}")


@pytest.mark.parametrize("shape", [(37, 53, 3), (64, 64, 1), (240, 320, 3)])
@pytest.mark.parametrize("mode", ["abs", "rel", "rel_ref"])
def test_lbsp_oracle_equals_reference_extractor(shape, mode):
    """LBSP::compute2 of the reference (features2d/src/LBSP.cpp:102-152, SSE2 threshold path of LBSP.hpp:203-223) vs the oracle"""
    rng = np.random.RandomState(hash((shape, mode)) & 0xFFFF)
    img = rng.randint(0, 256, shape).astype(np.uint8)
    if shape[2] == 1:
        img = img[..., 0]
    ref = np.clip(img.astype(int) + rng.randint(-20, 21, img.shape), 0, 255).astype(np.uint8) if "ref" in mode else None
    kw = dict(thr=25) if mode == "abs" else dict(rel=0.333, thr=3)
    _same(R.lbsp_compute(img, ref=ref, **kw), O.lbsp_compute(img, ref=ref, **kw), "dense LBSP map")


def test_helpers_equal_reference_headers():
    """lv::cdist<3>, lv::L1dist<3> (uint8 wrap, quirk Q1), lv::hdist<3>, the 7x7 sampling walk and both neighbour patterns, straight from
    the reference's utils/math.hpp and utils/opencv.hpp"""
    import ctypes as C
    L, Lo = R.lib(), O.lib()
    rng = np.random.RandomState(5)
    for _ in range(3000):
        a, b = rng.randint(0, 256, 3).astype(np.uint8), rng.randint(0, 256, 3).astype(np.uint8)
        pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
        assert L.ref_cdist3(pa, pb) == Lo.lvo_cdist3(pa, pb)
        assert L.ref_L1dist3_u8(pa, pb) == (int(np.abs(a.astype(int) - b.astype(int)).sum()) & 0xFF)
        da, db = rng.randint(0, 65536, 3).astype(np.uint16), rng.randint(0, 65536, 3).astype(np.uint16)
        assert L.ref_hdist3(da.ctypes.data_as(C.c_void_p), db.ctypes.data_as(C.c_void_p)) == sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(da, db))
    x, y = C.c_int(), C.c_int()
    xy = (C.c_int * 2)()
    for rnd in list(range(0, 1100)) + [2 ** 31 - 1, 123456789]:
        for ox, oy in ((0, 0), (5, 7), (63, 47), (30, 2)):
            L.ref_sample_pos_7x7(rnd, ox, oy, 2, 64, 48, C.byref(x), C.byref(y))
            Lo.lvo_sample_pos_7x7(rnd, ox, oy, 2, 64, 48, xy)
            assert (x.value, y.value) == (xy[0], xy[1])
            L.ref_neighbor_pos_3x3(rnd, ox, oy, 2, 64, 48, C.byref(x), C.byref(y))
            Lo.lvo_neighbor_pos(0, rnd, ox, oy, 2, 64, 48, xy)
            assert (x.value, y.value) == (xy[0], xy[1])
            L.ref_neighbor_pos_5x5(rnd, ox, oy, 2, 64, 48, C.byref(x), C.byref(y))
            Lo.lvo_neighbor_pos(1, rnd, ox, oy, 2, 64, 48, xy)
            assert (x.value, y.value) == (xy[0], xy[1])


def test_reference_errors():
    r = R.Reference(O.ALGO_SUBSENSE)
    with pytest.raises(R.ReferenceError_, match="0 or 255"):
        r.initialize(np.zeros((40, 40, 3), np.uint8), np.full((40, 40), 7, np.uint8))
    with pytest.raises(R.ReferenceError_, match="no useful pixels"):
        r.initialize(np.zeros((40, 40, 3), np.uint8), np.zeros((40, 40), np.uint8))


# ---- ViBe / PBAS: the reference's own BackgroundSubtractorViBe.cpp / BackgroundSubtractorPBAS.cpp, compiled unmodified ----
def _vp_frame(seq, t, c_in):
    f = seq.frame(t)
    return np.ascontiguousarray(f[..., 0]) if c_in == 1 and f.ndim == 3 else f


@pytest.mark.parametrize("model_c,c_in,w,h,n,seed", [(3, 3, 160, 120, 24, 0), (1, 1, 160, 120, 24, 1), (3, 1, 97, 73, 12, 2), (3, 3, 64, 48, 40, 3), (1, 1, 33, 7, 16, 4)])
def test_vibe_oracle_reference_order_equals_reference_source(model_c, c_in, w, h, n, seed):
    """masks every frame, the whole sample model every few frames, getBackgroundImage; learning rates 16 (default), 2 and 1; gray frames
    into the 3-channel model (cvtColor GRAY2BGR); the uint16-wrapping L2 distance of the 3-channel class included (math.hpp:301-306)"""
    seq = SynthSequence(w, h, 3 if c_in == 3 else 1, seed=40 + seed)
    r = R.ReferenceViBe(model_c, seed=seed)
    o = O.ViBeOracle(model_c, mode=O.MODE_REFERENCE, seed=seed)
    f0 = _vp_frame(seq, 0, c_in)
    r.initialize(f0); o.initialize(f0)
    _same(r.model(), o.model(), "ViBe model after initialize")
    for t in range(1, n + 1):
        f = _vp_frame(seq, t, c_in)
        lr = 16.0 if t % 5 else (2.0 if t % 10 else 1.0)
        _same(r.apply(f, lr), o.apply(f, lr), f"ViBe mask, frame {t}")
        if t % 4 == 0 or t == n:
            _same(r.model(), o.model(), f"ViBe model, frame {t}")
    _same(r.get_background_image(), o.get_background_image(), "ViBe getBackgroundImage")


@pytest.mark.parametrize("model_c,c_in,w,h,n,seed", [(3, 3, 160, 120, 20, 0), (1, 1, 160, 120, 20, 1), (3, 1, 97, 73, 10, 2), (3, 3, 64, 48, 40, 3), (1, 1, 40, 9, 16, 4)])
def test_pbas_oracle_reference_order_equals_reference_source(model_c, c_in, w, h, n, seed):
    """masks every frame; colour and gradient models, R(x), T(x), mean-min-distance maps (floats bit for bit) and m_fFormerMeanGradDist every
    few frames; learning-rate overrides; the gradient image goes through the cvcompat stages (blur, Scharr, convertScaleAbs, addWeighted),
    whose composition is pinned against cv2 in tests/test_pbas_oracle_cpu.py"""
    seq = SynthSequence(w, h, 3 if c_in == 3 else 1, seed=60 + seed)
    r = R.ReferencePBAS(model_c, seed=seed)
    o = O.PBASOracle(model_c, mode=O.MODE_REFERENCE, seed=seed)
    f0 = _vp_frame(seq, 0, c_in)
    r.initialize(f0); o.initialize(f0)

    def compare(tag):
        for name in ("bg_color", "bg_grad", "R", "T", "meanmin"):
            _same(r.state_get(name), o.state_get(name), f"PBAS '{name}', {tag}")
        assert r.state_get("scalars")[1] == o.state_get("scalars")[1], f"PBAS former mean gradient distance, {tag}"
    compare("after initialize")
    for t in range(1, n + 1):
        f = _vp_frame(seq, t, c_in)
        lr = -1.0 if t % 6 else (4.0 if t % 12 else 1.0)
        _same(r.apply(f, lr), o.apply(f, lr), f"PBAS mask, frame {t}")
        if t % 4 == 0 or t == n:
            compare(f"frame {t}")
    _same(r.get_background_image(), o.get_background_image(), "PBAS getBackgroundImage")


# ---- EdgeDetectorLBSP: the reference's own imgproc/src/EdgeDetectorLBSP.cpp, compiled unmodified ----
def _edge_frame(seq, t, ch):
    f = np.ascontiguousarray(seq.frame(t))
    if ch == 1 and f.ndim == 3:
        return np.ascontiguousarray(f[..., 0])
    if ch in (2, 4):
        extra = (f[..., :1].astype(np.int32) * 3 + f[..., 1:2] * 5 + 17 * t) % 256
        return np.ascontiguousarray(np.concatenate([f, extra.astype(np.uint8)], axis=2)[..., :ch] if ch == 4 else f[..., :2])
    return f


@pytest.mark.parametrize("w,h,ch,levels,hyst", [(96, 72, 3, 3, 0.5), (97, 73, 1, 3, 0.5), (96, 73, 3, 2, 0.25), (97, 72, 1, 1, 0.75), (160, 120, 4, 3, 0.5),
                                                 (43, 41, 2, 2, 0.5), (320, 240, 3, 3, 0.5)])
def test_edge_oracle_equals_reference_source(w, h, ch, levels, hyst):
    """one object per side, a sequence of calls (the detector's maps persist between calls): edge masks for several thresholds incl. out-of-range
    ones, the confidence map of apply(), and after every call the detector's two persistent buffers byte for byte (padded gradient map with the
    little-endian initial value, padded edge mask with the two rows the suppression never writes)"""
    seq = SynthSequence(w, h, 1 if ch == 1 else 3, seed=w + h + ch)
    r, o = R.ReferenceEdgeDetectorLBSP(levels, hyst), O.EdgeDetectorLBSPOracle(levels, hyst)
    for t, thr in [(3, 0.5), (5, 0.25), (7, 0.75), (9, 0.0), (11, 0.9), (12, -1.0), (13, 1.0)]:
        f = _edge_frame(seq, t, ch)
        _same(r.apply_threshold(f, thr), o.apply_threshold(f, thr), f"edge mask, frame {t}, threshold {thr}")
        _same(r.raw(0), o.raw(0), f"gradient map buffer, frame {t}")
        _same(r.raw(1), o.raw(1), f"edge mask buffer, frame {t}")
    f = _edge_frame(seq, 14, ch)
    a, b = r.apply(f), o.apply(f)
    _same(a, b, "confidence map of apply()")
    assert b.any()
    _same(r.raw(1), o.raw(1), "edge mask buffer after apply()")


def test_edge_normalized_output_equals_reference_source():
    """bNormalizeOutput = true: cv::normalize(NORM_MINMAX) of the confidence map (the cvcompat call forwards to the restatement that
    tests/test_edge_oracle_cpu.py pins against cv2)"""
    seq = SynthSequence(160, 120, 3, seed=31)
    r, o = R.ReferenceEdgeDetectorLBSP(3, 0.5, True), O.EdgeDetectorLBSPOracle(3, 0.5, normalize_output=True)
    for t in (4, 9):
        _same(r.apply(seq.frame(t)), o.apply(seq.frame(t)), f"normalised confidence map, frame {t}")
