"""CPU tests of the PBAS oracle (oracle/lvo_pbas.hpp). The reference has no test for PBAS (the restatement itself is pinned to the reference's source by tests/test_ref_pin_cpu.py); these pin the restated
OpenCV arithmetic of the gradient image against cv2 and measure the gap between the two oracle modes."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence


@pytest.mark.parametrize("shape", [(37, 53, 3), (20, 31), (5, 4, 3), (3, 3), (2, 7), (1, 9, 3), (6, 1), (240, 320, 3), (295, 480)])
def test_gradient_image_matches_cv2(oracle, shape):
    """PBAS.cpp:125-134: GaussianBlur(3x3) -> Scharr x/y (16S) -> convertScaleAbs -> addWeighted(0.5, 0.5), all BORDER_DEFAULT"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(hash(shape) & 0xFFFF)
    for img in (rng.integers(0, 256, shape, dtype=np.uint8), (rng.integers(0, 2, shape) * 255).astype(np.uint8)):
        bl = cv2.GaussianBlur(img, (3, 3), 0, 0, borderType=cv2.BORDER_DEFAULT)
        gx = cv2.Scharr(bl, cv2.CV_16S, 1, 0, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
        gy = cv2.Scharr(bl, cv2.CV_16S, 0, 1, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
        want = cv2.addWeighted(cv2.convertScaleAbs(gx), 0.5, cv2.convertScaleAbs(gy), 0.5, 0)
        assert np.array_equal(oracle.pbas_gradient_image(img), want)


@pytest.mark.parametrize("ch", [1, 3])
def test_snapshot_mode_within_seed_noise_of_reference_order(oracle, ch):
    """tier 3: per-pixel disagreement and F-measure against the synthetic ground truth stay within the seed-to-seed spread of the
    reference order"""
    O = oracle
    seq = SynthSequence(160, 120, ch, seed=12)
    frames = [seq.frame(t, with_gt=True) for t in range(70)]

    def run(mode, seed):
        v = O.PBASOracle(ch, mode=mode, seed=seed)
        v.initialize(frames[0][0])
        return np.stack([v.apply(f) for f, _ in frames[1:]])[25:], v

    gts = np.stack([g for _, g in frames[1:]])[25:]
    fm = lambda m: 2 * ((m > 0) & gts).sum() / max(2 * ((m > 0) & gts).sum() + ((m > 0) & ~gts).sum() + ((m == 0) & gts).sum(), 1)
    refs = [run(O.MODE_REFERENCE, s)[0] for s in (1, 2, 3)]
    snaps = [run(O.MODE_SNAPSHOT, s) for s in (1, 2)]
    noise = max((refs[i] != refs[j]).mean() for i in range(3) for j in range(i + 1, 3))
    gap = max((s != r).mean() for s, _ in snaps for r in refs)
    assert gap <= 1.5 * noise + 0.002, (gap, noise)
    f_ref, f_snap = [fm(r) for r in refs], [fm(s) for s, _ in snaps]
    assert abs(np.mean(f_snap) - np.mean(f_ref)) <= (max(f_ref) - min(f_ref)) + 0.02, (f_snap, f_ref)
    v = snaps[0][1]
    R, T, mm = v.state_get("R"), v.state_get("T"), v.state_get("meanmin")
    assert R.min() >= 0.6 * 0.95 - 1e-6 and R.max() <= 99 * 1.05 and T.min() >= 2.0 and T.max() <= 200.0 and mm.min() >= 0 and mm.max() <= 1.0
    assert v.state_get("scalars")[1] >= 20.0   # m_fFormerMeanGradDist floor (PBAS.cpp:224)

def test_self_diffusion_writes_the_neighbours_own_pixel(oracle):
    """BGSPBAS_USE_SELF_DIFFUSION (PBAS.cpp:190-191): every model sample of a pixel is one of that pixel's own past colours or an
    initial 7x7 sample, never a neighbour's current colour"""
    O = oracle
    h, w = 24, 32
    ys, xs = np.mgrid[0:h, 0:w]
    base = ((xs * 8) % 256).astype(np.uint8)          # columns differ by 8 grey levels: a neighbour's colour is never the pixel's own
    v = O.PBASOracle(1, mode=O.MODE_SNAPSHOT, seed=3)
    v.initialize(base)
    m0 = v.state_get("bg_color")[..., 0].copy()
    for t in range(30):
        v.apply(base, 1.0)   # lr override 1: every background pixel writes itself and one neighbour each frame
    m = v.state_get("bg_color")[..., 0]
    changed = m != m0
    assert changed.any()
    assert (m[changed] == np.broadcast_to(base, m.shape)[changed]).all()


def test_gray_input_to_3ch_model_and_errors(oracle):
    O = oracle
    seq = SynthSequence(64, 48, 1, seed=2)
    a, b = O.PBASOracle(3, seed=4), O.PBASOracle(3, seed=4)
    f0 = seq.frame(0)
    a.initialize(f0); b.initialize(np.repeat(f0[..., None], 3, axis=2))
    for t in range(1, 10):
        f = seq.frame(t)
        assert np.array_equal(a.apply(f), b.apply(np.repeat(f[..., None], 3, axis=2)))
    assert np.array_equal(a.state_get("bg_grad"), b.state_get("bg_grad"))
    with pytest.raises(O.OracleError):
        O.PBASOracle(1).initialize(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(O.OracleError):
        O.PBASOracle(3, update_rate=0.0)
