"""CPU tests (no GPU): the oracle is pinned against every golden vector / known answer the reference's own tests hold
for this path, its OpenCV-equivalent mask ops are cross-checked against cv2 when available, and the
two oracle modes are compared statistically."""
import ctypes as C
import os

import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_lbsp_golden_vector(oracle):
    """modules/features2d/test/lbsp.cpp:21-79 + test/data/test_lbsp.bin (bit-exact on the valid zone)"""
    g = np.load(os.path.join(GOLDEN, "lbsp_golden.npz"))
    d = oracle.lbsp_compute(g["crop"], thr=int(g["abs_threshold"]))
    assert d.shape == g["desc"].shape == (65, 65, 3)
    assert np.array_equal(d[2:63, 2:63], g["desc"][2:63, 2:63])


def test_lbsp_threshold_sse_equals_scalar(oracle):
    L = oracle.lib()
    rng = np.random.RandomState(0)
    for _ in range(2000):
        vals = rng.randint(0, 256, 16).astype(np.uint8)
        ref, t = int(rng.randint(0, 256)), int(rng.randint(0, 256))
        p = vals.ctypes.data_as(C.c_void_p)
        assert L.lvo_lbsp_threshold(p, ref, t, 0) == L.lvo_lbsp_threshold(p, ref, t, 1)
    # pattern order (LBSP.hpp:292-294): a single bright neighbour sets exactly its own bit
    img = np.zeros((5, 5), np.uint8)
    dx = [-2, 2, 0, 0, -2, 2, 2, -2, 0, -1, 0, 1, -1, 1, 1, -1]
    dy = [0, 0, -2, 2, 2, -2, 2, -2, 1, 0, -1, 0, -1, 1, -1, 1]
    for n in range(16):
        im = img.copy(); im[2 + dy[n], 2 + dx[n]] = 200
        assert oracle.lbsp_compute(im, thr=20)[2, 2] == 1 << n


def test_l1dist_hdist_cdist_known_answers(oracle):
    """modules/utils/test/math.cpp:287-420 (L1dist), :762-860 (cdist); plus the uint8 wrap of quirk Q1"""
    L = oracle.lib()
    u8 = lambda *v: np.array(v, np.uint8).ctypes.data_as(C.c_void_p)
    assert L.lvo_l1dist3_u8(u8(0, 0, 0), u8(1, 2, 3)) == 6
    assert L.lvo_l1dist3_u8(u8(200, 200, 200), u8(0, 0, 0)) == 600 % 256 == 88      # Q1: wraps mod 256
    assert L.lvo_cdist2(u8(255, 0), u8(0, 255)) == 255
    assert L.lvo_cdist3(u8(7, 7, 7), u8(9, 9, 9)) == 0                                 # gray vs gray
    assert L.lvo_cdist3(u8(10, 20, 30), u8(10, 20, 30)) == 0                           # equal
    assert L.lvo_cdist4(u8(0, 255, 0, 255), u8(255, 0, 255, 0)) == int(np.floor(np.sqrt(2 * 255.0 ** 2)))


def test_sampling_patterns(oracle):
    """modules/utils/test/opencv.cpp:369-429: clamp arithmetic, the 'r = 1+rand%tot, subtract until <=0' walk, neighbour != centre"""
    L = oracle.lib()
    xy = (C.c_int * 2)()
    pat = [[2, 4, 6, 7, 6, 4, 2], [4, 8, 12, 14, 12, 8, 4], [6, 12, 21, 25, 21, 12, 6], [7, 14, 25, 28, 25, 14, 7],
           [6, 12, 21, 25, 21, 12, 6], [4, 8, 12, 14, 12, 8, 4], [2, 4, 6, 7, 6, 4, 2]]
    assert sum(map(sum, pat)) == 512
    hist = np.zeros((7, 7), int)
    for r in range(512):
        L.lvo_sample_pos_7x7(r, 50, 50, 2, 100, 100, xy)
        hist[xy[1] - 47, xy[0] - 47] += 1
    assert np.array_equal(hist, np.array(pat))
    L.lvo_sample_pos_7x7(0, 0, 0, 2, 100, 100, xy); assert (xy[0], xy[1]) == (2, 2)          # clamped to the border
    L.lvo_sample_pos_7x7(511, 99, 99, 2, 100, 100, xy); assert (xy[0], xy[1]) == (97, 97)
    n3 = [(-1, 1), (0, 1), (1, 1), (-1, 0), (1, 0), (-1, -1), (0, -1), (1, -1)]
    for r in range(16):
        L.lvo_neighbor_pos(0, r, 50, 50, 2, 100, 100, xy)
        assert (xy[0] - 50, xy[1] - 50) == n3[r % 8]
    seen = set()
    for r in range(24):
        L.lvo_neighbor_pos(1, r, 50, 50, 2, 100, 100, xy)
        seen.add((xy[0] - 50, xy[1] - 50))
    assert len(seen) == 24 and (0, 0) not in seen and all(abs(a) <= 2 and abs(b) <= 2 for a, b in seen)


def test_glibc_rand_clone_matches_libc(oracle):
    libc = C.CDLL(None)
    for seed in (0, 1, 42, 123456789):
        out = (C.c_int * 64)()
        oracle.lib().lvo_glibc_rand_seq(seed, 64, out)
        libc.srand(seed)
        assert [libc.rand() for _ in range(64)] == list(out)


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10"""
    L = oracle.lib()

    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.lvo_philox(c, k, o)
        return list(o)
    assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_lut_values(oracle):
    """BackgroundSubtractorLBSP.cpp:29-30,42-43 with cv::saturate_cast rounding (SURVEY Appendix E examples)"""
    lut = np.zeros(256, np.uint8)
    oracle.lib().lvo_build_lut(3, C.c_float(0.333), 0, lut.ctypes.data_as(C.c_void_p))
    assert [int(lut[t]) for t in (1, 2, 3, 100, 200, 255)] == [0, 1, 1, 33, 67, 85]
    oracle.lib().lvo_build_lut(1, C.c_float(0.333), 0, lut.ctypes.data_as(C.c_void_p))
    assert int(lut[255]) == 28 and int(lut[0]) == 0


def test_mask_ops_match_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    L = oracle.lib()
    rng = np.random.RandomState(3)
    for (h, w) in ((60, 80), (37, 53), (120, 161)):
        m = ((rng.rand(h, w) < 0.4) * 255).astype(np.uint8)
        blobs = np.zeros((h, w), np.uint8)
        cv2.circle(blobs, (w // 2, h // 2), min(h, w) // 3, 255, 3)
        cv2.rectangle(blobs, (5, 5), (w // 3, h // 3), 255, 2)
        for src in (m, blobs):
            src = src.copy(); src[:2] = 0; src[-2:] = 0; src[:, :2] = 0; src[:, -2:] = 0
            out = np.empty_like(src)
            p, q = src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)
            for it in (1, 3):
                L.lvo_morph_rect(p, q, w, h, it, 1); assert np.array_equal(out, cv2.dilate(src, None, iterations=it))
                L.lvo_morph_rect(p, q, w, h, it, 0); assert np.array_equal(out, cv2.erode(src, None, iterations=it))
            for k in (3, 9, 11, 13):
                L.lvo_median_binary(p, q, w, h, k); assert np.array_equal(out, cv2.medianBlur(src, k))
            fl = src.copy()
            L.lvo_floodfill_origin(fl.ctypes.data_as(C.c_void_p), w, h)
            ref = src.copy(); cv2.floodFill(ref, None, (0, 0), 255)
            assert np.array_equal(fl, ref)
    img = rng.randint(0, 256, (64, 96, 3)).astype(np.uint8)
    ds = np.empty((8, 12, 3), np.uint8)
    L.lvo_resize_area_exact(img.ctypes.data_as(C.c_void_p), 96, 64, 3, 8, ds.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ds, cv2.resize(img, (12, 8), interpolation=cv2.INTER_AREA))
    # non-integer shrink factors (CDnet frame sizes that are not multiples of 8): OpenCV's general area path, bit-exact
    for (w, h) in ((570, 340), (640, 364), (595, 245), (700, 450), (624, 420), (480, 295), (323, 243)):
        for cn in (1, 3):
            img = rng.randint(0, 256, (h, w, cn)).astype(np.uint8)
            ds = np.empty((h // 8, w // 8, cn), np.uint8)
            L.lvo_resize_area_general(img.ctypes.data_as(C.c_void_p), w, h, cn, w // 8, h // 8, ds.ctypes.data_as(C.c_void_p))
            ref = cv2.resize(img if cn == 3 else img[..., 0], (w // 8, h // 8), interpolation=cv2.INTER_AREA).reshape(ds.shape)
            assert np.array_equal(ds, ref), f"INTER_AREA {w}x{h}x{cn}: {(ds != ref).sum()} mismatches"


def _fmeasure(m, gt):
    tp = ((m > 0) & gt).sum(); fp = ((m > 0) & ~gt).sum(); fn = ((m == 0) & gt).sum()
    return 2 * tp / max(2 * tp + fp + fn, 1)


@pytest.mark.parametrize("algo,c,lr", [("subsense", 3, None), ("lobster", 1, 16.0)])
def test_snapshot_mode_within_seed_noise_of_reference_order(oracle, algo, c, lr):
    """Tier-3 tolerance derivation: the snapshot semantics the GPU implements must disagree with the
    reference-order semantics by no more than reference-order runs with different seeds disagree among themselves
    (+ margin), and the F-measure against the synthetic ground truth must stay within the seed-to-seed spread."""
    w, h, n = 160, 120, 90
    seq = SynthSequence(w, h, c, seed=3)
    frames = [seq.frame(t, with_gt=True) for t in range(n)]
    A = oracle.ALGO_SUBSENSE if algo == "subsense" else oracle.ALGO_LOBSTER

    def run(mode, seed):
        o = oracle.Oracle(A, mode=mode, seed=seed)
        o.initialize(frames[0][0])
        masks = []
        for t in range(1, n):
            masks.append(o.apply(frames[t][0], lr if lr else (1.0 if t <= 50 else 0.0)))
        return masks
    ref = [run(oracle.MODE_REFERENCE, s) for s in (1, 2, 3)]
    snap = [run(oracle.MODE_SNAPSHOT, s) for s in (1, 2)]
    tail = range(60, n - 1)
    dis = lambda a, b: float(np.mean([(a[t] != b[t]).mean() for t in tail]))
    fm = lambda a: float(np.mean([_fmeasure(a[t], frames[t + 1][1]) for t in tail]))
    seed_noise = max(dis(ref[0], ref[1]), dis(ref[0], ref[2]), dis(ref[1], ref[2]))
    cross = max(dis(snap[0], ref[0]), dis(snap[1], ref[1]), dis(snap[0], ref[2]))
    assert cross <= 1.5 * seed_noise + 0.002, (cross, seed_noise)
    f_ref = [fm(r) for r in ref]; f_snap = [fm(s) for s in snap]
    spread = max(f_ref) - min(f_ref)
    assert abs(np.mean(f_snap) - np.mean(f_ref)) <= spread + 0.02, (f_snap, f_ref)
    assert np.mean(f_snap) > 0.5


def test_oracle_determinism_and_api_errors(oracle):
    seq = SynthSequence(64, 48, 3, seed=1)
    a, b = (oracle.Oracle(oracle.ALGO_SUBSENSE, mode=oracle.MODE_SNAPSHOT, seed=5) for _ in range(2))
    a.initialize(seq.frame(0)); b.initialize(seq.frame(0))
    for t in range(1, 6):
        assert np.array_equal(a.apply(seq.frame(t), 1.0), b.apply(seq.frame(t), 1.0))
    assert np.array_equal(a.get_background_image(), b.get_background_image())
    with pytest.raises(oracle.OracleError, match="0 or 255"):
        a.initialize(seq.frame(0), np.full((48, 64), 9, np.uint8))
    lob = oracle.Oracle(oracle.ALGO_LOBSTER)
    lob.initialize(seq.frame(0))
    with pytest.raises(oracle.OracleError, match="positive"):
        lob.apply(seq.frame(1), 0.0)
    # small frame -> "small" branch of SuBSENSE.cpp:121-128
    sc = a.state_get("scalars")
    assert sc[4] == 0 and sc[7] == 4.0 and sc[8] == 512.0 and sc[6] == 9


def test_lbsp_gradient_known_answers(oracle):
    """LBSP::computeDescriptor_gradient (features2d LBSP.hpp:235-256) on patterns whose answer follows from the masks of :288-291"""
    O = oracle
    flat = np.full((9, 9), 100, np.uint8)
    assert not O.lbsp_gradient(flat).any()                       # no neighbour differs by more than t = ((100 >> 2) + 20) / 2 = 22
    step = flat.copy(); step[:, 5:] = 200                        # vertical edge right of the centre column
    g = O.lbsp_gradient(step)[4, 4]
    # neighbours with dx > 0 differ (bits 1, 5, 6 at dx = +2; 11, 13, 14 at dx = +1): six bits, all in the GradX_Neg mask
    assert g[2] == 6 and np.int8(g[0]) == -6 and np.int8(g[1]) == 0
    g = O.lbsp_gradient(np.ascontiguousarray(step.T))[4, 4]      # horizontal edge below the centre row: dy > 0 bits = GradY_Pos
    assert g[2] == 6 and np.int8(g[0]) == 0 and np.int8(g[1]) == 6
    # three channels: the channel with the most differing neighbours wins; ties keep the LAST channel
    rgb = np.stack([flat, step, flat], axis=2)
    assert np.array_equal(O.lbsp_gradient(rgb)[4, 4], O.lbsp_gradient(step)[4, 4])
    assert not O.lbsp_gradient(rgb)[:2].any() and not O.lbsp_gradient(rgb)[:, -2:].any()   # 2-px border: zero pattern


def test_cdist_l1dist_full_known_answer_lists_of_the_reference(oracle):
    """every uchar-applicable known answer of modules/utils/test/math.cpp: cdist<2> (:782-795, the same list again for std::array at
    :799-812), the cdist range / self-distance properties of :764-781 on random arrays for 2, 3 and 4 channels, and the L1dist<3> array
    cases of :349-353 / :357-361"""
    L = oracle.lib()
    L.lvo_cdist2.restype = L.lvo_cdist3.restype = L.lvo_cdist4.restype = C.c_uint64
    u8 = lambda *v: np.array(v, np.uint8).ctypes.data_as(C.c_void_p)
    cd2 = [((0, 0), (0, 0), 0), ((1, 0), (0, 0), 1), ((255, 0), (0, 0), 255), ((0, 0), (1, 0), 0), ((0, 0), (255, 0), 0), ((1, 0), (0, 1), 1),
           ((0, 1), (1, 0), 1), ((255, 0), (0, 255), 255), ((0, 255), (255, 0), 255), ((255, 0), (0, 1), 255), ((1, 1), (1, 1), 0),
           ((255, 255), (255, 255), 0), ((1, 1), (255, 255), 0), ((255, 255), (1, 1), 0)]
    for a, b, want in cd2:
        assert L.lvo_cdist2(u8(*a), u8(*b)) == want, (a, b)
    assert L.lvo_l1dist3_u8(u8(1, 2, 3), u8(0, 0, 0)) == 6
    rng = np.random.default_rng(17)
    for hi in (2, 256):                                               # genarray(0,1) and genarray(0,255)
        v = rng.integers(0, hi, 20004, dtype=np.uint8)
        for i in range(0, 10000, 7):
            a, b = v[i:i + 4].copy(), v[i + 10000:i + 10004].copy()
            for c, f in ((2, L.lvo_cdist2), (3, L.lvo_cdist3), (4, L.lvo_cdist4)):
                d = f(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
                assert 0 <= d <= (hi - 1) * c, (a, b, c, d)
                assert f(a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) == 0


def test_clamp_and_neighbour_known_answers_of_the_reference(oracle):
    """modules/utils/test/opencv.cpp:369-388 (clampImageCoords at 640x480 with borders 0 and 5, reached through the neighbour walk) and
    :415-429 (10 000 draws of the 8-neighbour pattern stay within +-1 of the origin and never return it)"""
    L = oracle.lib()
    xy = (C.c_int * 2)()
    L.lvo_neighbor_pos(0, 5, 0, 0, 0, 640, 480, xy); assert (xy[0], xy[1]) == (0, 0)          # (-1,-1) clamped, border 0
    L.lvo_neighbor_pos(0, 5, 0, 0, 5, 640, 480, xy); assert (xy[0], xy[1]) == (5, 5)          # border 5
    L.lvo_neighbor_pos(0, 2, 639, 479, 0, 640, 480, xy); assert (xy[0], xy[1]) == (639, 479)  # (+1,+1) from the last pixel
    L.lvo_neighbor_pos(0, 2, 639, 479, 5, 640, 480, xy); assert (xy[0], xy[1]) == (634, 474)
    L.lvo_neighbor_pos(0, 4, 320, 240, 5, 640, 480, xy); assert (xy[0], xy[1]) == (321, 240)  # interior: untouched by the clamp
    rng = np.random.default_rng(23)
    for r in rng.integers(0, 2 ** 31 - 1, 10000):
        L.lvo_neighbor_pos(0, int(r), 320, 240, 0, 640, 480, xy)
        assert 319 <= xy[0] <= 321 and 239 <= xy[1] <= 241 and (xy[0], xy[1]) != (320, 240)
