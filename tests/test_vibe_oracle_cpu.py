"""CPU tests of the ViBe oracle (oracle/lvo_vibe.hpp): known answers of the distance quirk, the two oracle modes against each other,
and the restated getBackgroundImage. The reference has no test for ViBe (the restatement itself is pinned to the reference's source by tests/test_ref_pin_cpu.py): these pin it to the
source's arithmetic (video/src/BackgroundSubtractorViBe.cpp, utils/math.hpp:301-306, 391-397)."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence


def test_l2dist_accumulates_in_uint16(oracle):
    O = oracle
    # lv::L2dist<3,uchar>: sum of squares in uint16, float sqrt, `< thr*3` (ViBe.cpp:167-171)
    assert O.vibe_match(3, 20, [0, 0, 0], [34, 34, 34])            # sqrt(3468) = 58.9 < 60
    assert not O.vibe_match(3, 20, [0, 0, 0], [35, 35, 35])        # sqrt(3675) = 60.6
    assert not O.vibe_match(3, 20, [0, 0, 0], [60, 0, 0])          # strict '<'
    assert O.vibe_match(3, 20, [0, 0, 0], [59, 10, 0])             # 3581
    assert not O.vibe_match(3, 20, [0, 0, 0], [255, 10, 10])       # 65225: no wrap yet
    assert O.vibe_match(3, 20, [0, 0, 0], [255, 22, 6])            # 65545 wraps to 9: a far colour that "matches"
    assert not O.vibe_match(3, 20, [0, 0, 0], [200, 200, 200])     # 120000 mod 65536 = 54464
    # 1 channel: L1 < thr (ViBe.cpp:93)
    assert O.vibe_match(1, 20, [100], [119]) and not O.vibe_match(1, 20, [100], [120]) and O.vibe_match(1, 20, [100], [81])


def test_wrap_quirk_matches_numpy_restatement(oracle):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (4000, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (4000, 3), dtype=np.uint8)
    d = a.astype(np.int64) - b.astype(np.int64)
    acc = (d * d).sum(1) & 0xFFFF
    want = np.sqrt(acc.astype(np.float32)) < np.float32(60)
    got = np.array([oracle.vibe_match(3, 20, a[i], b[i]) for i in range(len(a))])
    assert np.array_equal(got, want)
    assert (want != ((d * d).sum(1) < 3600)).any()   # the wrap is observable on random colours


@pytest.mark.parametrize("ch", [1, 3])
def test_snapshot_mode_within_seed_noise_of_reference_order(oracle, ch):
    """tier 3: the parallel semantics (queued neighbour writes, Philox) stay within the seed-to-seed spread of the reference order, in
    per-pixel disagreement and in F-measure against the synthetic ground truth"""
    O = oracle
    seq = SynthSequence(160, 120, ch, seed=11)
    frames = [seq.frame(t, with_gt=True) for t in range(60)]

    def run(mode, seed):
        v = O.ViBeOracle(ch, mode=mode, seed=seed)
        v.initialize(frames[0][0])
        return np.stack([v.apply(f) for f, _ in frames[1:]])[20:]

    gts = np.stack([g for _, g in frames[1:]])[20:]
    fm = lambda m: 2 * ((m > 0) & gts).sum() / max(2 * ((m > 0) & gts).sum() + ((m > 0) & ~gts).sum() + ((m == 0) & gts).sum(), 1)
    refs = [run(O.MODE_REFERENCE, s) for s in (1, 2, 3)]
    snaps = [run(O.MODE_SNAPSHOT, s) for s in (1, 2)]
    noise = max((refs[i] != refs[j]).mean() for i in range(3) for j in range(i + 1, 3))
    gap = max((s != r).mean() for s in snaps for r in refs)
    assert gap <= 1.5 * noise + 0.002, (gap, noise)
    f_ref, f_snap = [fm(r) for r in refs], [fm(s) for s in snaps]
    assert abs(np.mean(f_snap) - np.mean(f_ref)) <= (max(f_ref) - min(f_ref)) + 0.02, (f_snap, f_ref)

def test_gray_input_to_3ch_model_equals_replicated_bgr(oracle):
    O = oracle
    seq = SynthSequence(64, 48, 1, seed=2)
    a, b = O.ViBeOracle(3, seed=4), O.ViBeOracle(3, seed=4)
    f0 = seq.frame(0)
    a.initialize(f0); b.initialize(np.repeat(f0[..., None], 3, axis=2))
    for t in range(1, 12):
        f = seq.frame(t)
        assert np.array_equal(a.apply(f), b.apply(np.repeat(f[..., None], 3, axis=2)))
    assert np.array_equal(a.model(), b.model())
    with pytest.raises(O.OracleError):
        O.ViBeOracle(1).initialize(np.zeros((8, 8, 3), np.uint8))


def test_init_samples_come_from_the_7x7_neighbourhood_border0(oracle):
    O = oracle
    h, w = 20, 24
    img = (np.arange(h * w, dtype=np.uint32).reshape(h, w) % 251).astype(np.uint8)
    # encode the position instead: two runs on x / y coordinate images
    xs = np.tile(np.arange(w, dtype=np.uint8), (h, 1)); ys = np.tile(np.arange(h, dtype=np.uint8)[:, None], (1, w))
    vx, vy = O.ViBeOracle(1, seed=9), O.ViBeOracle(1, seed=9)
    vx.initialize(xs); vy.initialize(ys)
    mx, my = vx.model()[..., 0].astype(int), vy.model()[..., 0].astype(int)
    assert (np.abs(mx - xs[None]) <= 3).all() and (np.abs(my - ys[None]) <= 3).all()
    assert mx.min() == 0 and mx.max() == w - 1 and my.min() == 0 and my.max() == h - 1   # border 0: edge pixels are sampled
    assert (mx != xs[None]).any() and (my != ys[None]).any()
    del img


def test_background_image_is_float_mean_round_half_even(oracle):
    O = oracle
    seq = SynthSequence(48, 40, 3, seed=6)
    v = O.ViBeOracle(3, seed=1)
    v.initialize(seq.frame(0))
    for t in range(1, 6):
        v.apply(seq.frame(t), 2.0)
    m = v.model().astype(np.float32)
    acc = np.zeros(m.shape[1:], np.float32)
    for s in range(m.shape[0]):
        acc = acc + m[s] / np.float32(m.shape[0])
    assert np.array_equal(v.get_background_image(), np.clip(np.rint(acc), 0, 255).astype(np.uint8))


def test_integer_form_of_the_l2_test_is_exact():
    """the kernel evaluates `(float)sqrt(n) < T` (ViBe.cpp:171 through lv::L2dist) as `n < T*T`: exhaustive over every uint16 sum and every
    threshold the API accepts (T = 3 * nColorDistThreshold <= 765)"""
    n = np.arange(65536, dtype=np.float32)
    r = np.sqrt(n)                                   # correctly rounded float32 square root, as std::sqrt(float)
    ni = np.arange(65536, dtype=np.int64)
    for T in range(0, 766):
        assert np.array_equal(r < np.float32(T), ni < T * T), T
