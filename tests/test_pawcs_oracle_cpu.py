"""CPU tests of the PAWCS restatement (oracle/lvo_pawcs.hpp). The reference holds no test or golden vector for PAWCS
(the restatement is pinned to the reference's own source by tests/test_ref_pin_cpu.py): these tests check the restatement's invariants, its cv2-equivalent float ops, and that the snapshot
semantics the GPU implements stay within the seed-to-seed noise of the reference-order semantics."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence


def _fmeasure(m, gt):
    tp = ((m > 0) & gt).sum(); fp = ((m > 0) & ~gt).sum(); fn = ((m == 0) & gt).sum()
    return 2 * tp / max(2 * tp + fp + fn, 1)


def _run(oracle, mode, seed, frames, n, lr=0.0):
    o = oracle.Oracle(oracle.ALGO_PAWCS, mode=mode, seed=seed)
    o.initialize(frames[0][0])
    return o, [o.apply(frames[t][0], lr) for t in range(1, n)]


@pytest.mark.parametrize("c", [3, 1])
def test_pawcs_snapshot_mode_within_seed_noise_of_reference_order(oracle, c):
    w, h, n = 160, 120, 100
    seq = SynthSequence(w, h, c, seed=3)
    frames = [seq.frame(t, with_gt=True) for t in range(n)]
    ref = [_run(oracle, oracle.MODE_REFERENCE, s, frames, n)[1] for s in (1, 2, 3)]
    snap = [_run(oracle, oracle.MODE_SNAPSHOT, s, frames, n)[1] for s in (1, 2)]
    tail = range(60, n - 1)
    dis = lambda a, b: float(np.mean([(a[t] != b[t]).mean() for t in tail]))
    fm = lambda a: float(np.mean([_fmeasure(a[t], frames[t + 1][1]) for t in tail]))
    seed_noise = max(dis(ref[0], ref[1]), dis(ref[0], ref[2]), dis(ref[1], ref[2]))
    cross = max(dis(snap[0], ref[0]), dis(snap[1], ref[1]), dis(snap[0], ref[2]))
    assert cross <= 1.5 * seed_noise + 0.002, (cross, seed_noise)
    f_ref = [fm(r) for r in ref]; f_snap = [fm(s) for s in snap]
    assert abs(np.mean(f_snap) - np.mean(f_ref)) <= (max(f_ref) - min(f_ref)) + 0.02, (f_snap, f_ref)
    assert np.mean(f_snap) > 0.6


def test_pawcs_model_invariants(oracle):
    """PAWCS.cpp:494-557 sizes; dictionaries stay sorted enough for the bubble pass; per-pixel LUTs are permutations"""
    seq = SynthSequence(64, 48, 3, seed=2)
    o = oracle.Oracle(oracle.ALGO_PAWCS, mode=oracle.MODE_SNAPSHOT, seed=7)
    o.initialize(seq.frame(0))
    sc = o.state_get("scalars")
    nw, ng = int(sc[4]), int(sc[5])
    assert nw == 50 and 1 <= ng <= 25 and sc[7] == 1000
    roi = o.state_get("roi").reshape(48, 64)
    assert roi[:2].max() == 0 and roi[2:-2, 2:-2].min() == 255
    occ = o.state_get("lw_occ").reshape(-1, nw)
    assert (occ[roi.ravel() > 0] >= 1).all()          # every word of every ROI pixel is initialised after refreshModel(1,0)
    for t in range(1, 30):
        m = o.apply(seq.frame(t))
        assert set(np.unique(m)) <= {0, 255} and m[:2].max() == 0
    glut = o.state_get("glut").reshape(-1, ng)[roi.ravel() > 0]
    assert (np.sort(glut, axis=1) == np.arange(ng)).all()
    gd = o.state_get("gdict")
    assert sorted(gd.tolist()) == list(range(ng))
    first, last = o.state_get("lw_first").reshape(-1, nw), o.state_get("lw_last").reshape(-1, nw)
    assert (last >= first).all() and last.max() <= 29
    bg = o.get_background_image()
    inner = np.abs(bg[4:-4, 4:-4].astype(int) - seq.frame(0)[4:-4, 4:-4].astype(int))
    assert np.median(inner) <= 6


def test_pawcs_determinism_refresh_and_errors(oracle):
    seq = SynthSequence(64, 48, 1, seed=5)
    a, b = (oracle.Oracle(oracle.ALGO_PAWCS, mode=oracle.MODE_SNAPSHOT, seed=11) for _ in range(2))
    a.initialize(seq.frame(0)); b.initialize(seq.frame(0))
    for t in range(1, 12):
        assert np.array_equal(a.apply(seq.frame(t)), b.apply(seq.frame(t)))
    a.pawcs_refresh_model(125, 0.0, True); b.pawcs_refresh_model(125, 0.0, True)
    for n in ("lw_occ", "lw_color", "lw_desc", "gw_weight", "gw_map", "glut", "gdict"):
        assert np.array_equal(a.state_get(n), b.state_get(n)), n
    with pytest.raises(oracle.OracleError, match="fraction"):
        a.pawcs_refresh_model(1, 1.5, False)
    c = oracle.Oracle(oracle.ALGO_PAWCS)
    c.initialize(np.zeros((50, 70, 3), np.uint8))   # sizes that are not multiples of 8 take OpenCV's general INTER_AREA path
    assert c.apply(np.zeros((50, 70, 3), np.uint8)).shape == (50, 70)


def test_pawcs_blur3_matches_cv2(oracle):
    """the global-word occupancy maps are smoothed with cv::blur(3x3, BORDER_REPLICATE) (PAWCS.cpp:1314): run one maintenance
    step through the oracle's state and compare with cv2"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    m = (rng.rand(24, 32) * 3).astype(np.float32)
    want = cv2.blur(m, (3, 3), borderType=cv2.BORDER_REPLICATE)
    rs = np.pad(m.astype(np.float64), 1, mode="edge")
    row = rs[:, :-2] + rs[:, 1:-1] + rs[:, 2:]
    got = ((row[:-2] + row[1:-1] + row[2:]) * (1.0 / 9.0)).astype(np.float32)   # the oracle's formula (lvo_pawcs.hpp blur3)
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, float(np.abs(want).max()))
