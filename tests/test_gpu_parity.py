"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
(oracle/, snapshot mode, same Philox seed) on identical inputs.

Tier 1: LBSP descriptors bit-exact (incl. the reference's golden vector).
Tier 2: raw per-pixel classification + every integer/byte state buffer bit-exact from an identical state snapshot.
Tier 3: end-to-end masks over a sequence: per-pixel disagreement <= TOL_MASK (we observe 0), float maps within 1e-5 rel.
"""
import os

import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FLOAT_RTOL = 1e-5   # north_star: "float feedback maps within 1e-5 relative"
TOL_MASK = 0.0      # fraction of pixels allowed to differ GPU vs oracle(snapshot), same seed

INT_STATE = ["lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "unstable", "blinks", "lastraw", "lastrawblink", "dilinv", "rawmask"]
FLT_STATE = ["T", "R", "v", "Dlast", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST"]


def test_lbsp_golden_vector(lv):
    g = np.load(os.path.join(GOLDEN, "lbsp_golden.npz"))
    d = lv.LBSP(int(g["abs_threshold"])).compute2(g["crop"])
    assert np.array_equal(d[2:63, 2:63], g["desc"][2:63, 2:63])
    assert not d[:2].any() and not d[:, :2].any()  # border untouched


@pytest.mark.parametrize("shape", [(37, 53, 3), (64, 64, 1), (240, 320, 3), (5, 5, 3), (9, 130, 1), (1080, 1920, 3)])
@pytest.mark.parametrize("mode", ["abs", "rel", "rel_ref", "abs_ref"])
def test_lbsp_matches_oracle(lv, oracle, shape, mode):
    rng = np.random.RandomState(hash((shape, mode)) & 0xFFFF)
    img = rng.randint(0, 256, shape).astype(np.uint8)
    if shape[2] == 1:
        img = img[..., 0]
    ref = (np.clip(img.astype(int) + rng.randint(-20, 21, img.shape), 0, 255)).astype(np.uint8) if "ref" in mode else None
    if mode.startswith("abs"):
        ext, kw = lv.LBSP(25), dict(thr=25)
    else:
        ext, kw = lv.LBSP(0.333, 3), dict(rel=0.333, thr=3)
    ext.setReference(ref)
    got = ext.compute2(img)
    want = oracle.lbsp_compute(img, ref=ref, **kw)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape", [(37, 53, 3), (64, 64, 1), (5, 5, 3), (9, 130, 1), (240, 320, 3), (1080, 1920, 3), (61, 97, 2), (48, 80, 4)])
def test_lbsp_gradient_matches_oracle(lv, oracle, shape):
    """dense LBSP::computeDescriptor_gradient (the per-pixel primitive of EdgeDetectorLBSP): bit-exact gradX / gradY / magnitude"""
    rng = np.random.RandomState(hash(shape) & 0xFFFF)
    img = rng.randint(0, 256, shape).astype(np.uint8)
    if shape[2] == 1:
        img = img[..., 0]
    got, want = lv.lbsp_gradient(img), oracle.lbsp_gradient(img)
    assert np.array_equal(got, want)
    smooth = np.ascontiguousarray(np.clip(np.cumsum(np.cumsum(rng.randint(-3, 4, shape), axis=0), axis=1) + 128, 0, 255).astype(np.uint8))
    if shape[2] == 1:
        smooth = smooth[..., 0]
    assert np.array_equal(lv.lbsp_gradient(smooth), oracle.lbsp_gradient(smooth))


def _mk(lv, oracle, algo, seed, **kw):
    if algo == "subsense":
        return lv.BackgroundSubtractorSuBSENSE(seed=seed, **kw), oracle.Oracle(oracle.ALGO_SUBSENSE, mode=oracle.MODE_SNAPSHOT, seed=seed)
    return lv.BackgroundSubtractorLOBSTER(seed=seed, **kw), oracle.Oracle(oracle.ALGO_LOBSTER, mode=oracle.MODE_SNAPSHOT, seed=seed)


def _compare_state(g, o, names_int, names_flt, tag):
    for n in names_int:
        a, b = g.state_get(n), o.state_get(n)
        nbad = int((a != b).sum())
        assert nbad == 0, f"{tag}: integer state '{n}' differs in {nbad} of {a.size} entries (first at {np.flatnonzero(a != b)[:5]})"
    for n in names_flt:
        a, b = g.state_get(n), o.state_get(n)
        assert np.allclose(a, b, rtol=FLOAT_RTOL, atol=1e-7), f"{tag}: float map '{n}' max abs diff {np.abs(a - b).max()}"
    sa, sb = g.state_get("scalars"), o.state_get("scalars")
    idx = list(range(13)) if names_flt else [0, 3, 10, 11, 12]
    assert np.allclose(sa[idx], sb[idx], rtol=1e-6), f"{tag}: scalars differ {sa[:13]} vs {sb[:13]}"


CASES = [
    ("subsense", 320, 240, 3, 12, None),
    ("subsense", 320, 240, 1, 8, None),
    ("subsense", 96, 72, 3, 8, None),       # "small" branch: no frame-level analysis, T in [4,512]
    ("subsense", 200, 150, 3, 6, "roi"),     # ragged width (not a multiple of 32) + user ROI
    ("subsense", 330, 250, 3, 8, None),      # frame-level analysis on a size that is not a multiple of 8 (general INTER_AREA path)
    ("subsense", 570, 340, 1, 6, None),      # CDnet twoPositionPTZCam size, 1 channel
    ("subsense", 9, 7, 3, 6, None),          # smallest useful frames: the ROI left by the 2-px border is 5x3 / 27x1
    ("subsense", 31, 5, 1, 5, None),
    ("subsense", 4100, 6, 1, 4, None),       # very wide and flat: 129 tiles in x, one (partial) tile row
    ("subsense", 37, 1030, 3, 3, None),      # narrow and tall: 129 tile rows, ragged width
    ("lobster", 33, 5, 3, 6, None),
    ("lobster", 320, 240, 1, 10, None),
    ("lobster", 320, 240, 3, 8, None),
    ("lobster", 75, 61, 1, 6, "roi"),
]


@pytest.mark.parametrize("algo,w,h,c,nframes,roi", CASES)
def test_tier2_state_parity(lv, oracle, algo, w, h, c, nframes, roi):
    """identical snapshot -> identical raw classification and integer state, frame after frame"""
    seq = SynthSequence(w, h, c, seed=11)
    roi_img = None
    if roi:
        roi_img = np.zeros((h, w), np.uint8)
        roi_img[h // 6:h - h // 8, w // 5:w - 3] = 255
        roi_img[h // 2:h // 2 + 5, w // 2:w // 2 + 9] = 0
    g, o = _mk(lv, oracle, algo, seed=7)
    f0 = seq.frame(0)
    g.initialize(f0, roi_img)
    o.initialize(f0, roi_img)
    ints = [n for n in INT_STATE if algo == "subsense" or n in ("lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "rawmask")]
    ints_init = [n for n in ints if n != "rawmask"]
    flts = FLT_STATE if algo == "subsense" else []
    assert np.array_equal(g.getROICopy().ravel(), o.state_get("roi"))
    _compare_state(g, o, ints_init, flts, f"{algo} init")
    for t in range(1, nframes + 1):
        f = seq.frame(t)
        lr = (1.0 if t <= 3 else 0.0) if algo == "subsense" else 16.0
        mg = g.apply(f, lr)
        mo = o.apply(f, lr)
        _compare_state(g, o, ints, flts, f"{algo} frame {t}")
        assert np.array_equal(mg, mo), f"{algo} frame {t}: final masks differ in {(mg != mo).sum()} px"
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())
    assert np.array_equal(g.getBackgroundDescriptorsImage(), o.get_background_descriptors_image())


def test_tier2_import_oracle_snapshot(lv, oracle):
    """run the ORACLE for a while, import its whole state into the GPU object, then classify one frame on both"""
    w, h, c = 320, 240, 3
    seq = SynthSequence(w, h, c, seed=5)
    g, o = _mk(lv, oracle, "subsense", seed=3)
    f0 = seq.frame(0)
    g.initialize(f0)
    o.initialize(f0)
    for t in range(1, 30):
        o.apply(seq.frame(t), 1.0 if t <= 20 else 0.0)
    for n in ["lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "unstable", "blinks", "lastraw", "lastrawblink", "dilinv"] + FLT_STATE + ["scalars"]:
        g.state_set(n, o.state_get(n))
    f = seq.frame(30)
    mg, mo = g.apply(f, 0.0), o.apply(f, 0.0)
    assert np.array_equal(g.state_get("rawmask"), o.state_get("rawmask"))
    assert np.array_equal(mg, mo)
    _compare_state(g, o, INT_STATE, FLT_STATE, "after import")


@pytest.mark.parametrize("check_every_frame", [True, False])
def test_subsense_scene_change_reset(lv, oracle, check_every_frame):
    """frame-level reset (SuBSENSE.cpp:584-600): a sudden scene change makes the frame tail request refreshModel(0.1) on the
    device. The conditional refresh then has to (1) wait for this frame's final mask, which is still being produced on the mask
    stream, and (2) apply the frame's queued neighbour writes before it resamples. With check_every_frame=False nothing reads the
    state between frames, so the kernels of consecutive frames really overlap (pipelined path); state is compared at the end."""
    w, h, c = 320, 240, 3
    seq_a, seq_b = SynthSequence(w, h, c, seed=21), SynthSequence(w, h, c, seed=22)
    g, o = _mk(lv, oracle, "subsense", seed=5)
    f0 = seq_a.frame(0)
    g.initialize(f0)
    o.initialize(f0)
    epoch0 = o.state_get("scalars")[12]
    masks_g, masks_o = [], []
    for t in range(1, 96):
        f = seq_a.frame(t) if t < 60 else (255 - seq_b.frame(t))   # hard cut at frame 60: the caps shrink, the reset fires around frame 79
        lr = 1.0 if t <= 10 else 0.0
        masks_g.append(g.apply(f, lr))
        masks_o.append(o.apply(f, lr))
        if check_every_frame:
            _compare_state(g, o, INT_STATE, FLT_STATE, f"scene change, frame {t}")
    assert o.state_get("scalars")[12] > epoch0, "the sequence did not trigger a model reset: the test does not test anything"
    for t, (mg, mo) in enumerate(zip(masks_g, masks_o), 1):
        assert np.array_equal(mg, mo), f"scene change, frame {t}: final masks differ in {(mg != mo).sum()} px"
    _compare_state(g, o, INT_STATE, FLT_STATE, "scene change, end")


@pytest.mark.parametrize("w,h", [(1920, 1080), (640, 480)])
def test_pipelined_runs_are_deterministic(lv, w, h):
    """race detector for the three-stream pipeline: the same sequence through two instances (one fed synchronously, one with two
    frames in flight and nothing reading state in between) and a third run of the first must end in byte-identical state"""
    seq = SynthSequence(w, h, 3, seed=77)
    frames = [seq.frame(t) for t in range(7)]
    order = [1, 2, 3, 4, 5, 6, 5, 4, 3, 2] * 12          # 120 frames of continuous motion

    def run(mode):
        g = lv.BackgroundSubtractorSuBSENSE(seed=3)
        g.initialize(frames[0])
        last = None
        for i, k in enumerate(order):
            lr = 1.0 if i < 30 else 0.0
            if mode == "sync":
                last = g.apply(frames[k], lr)
            else:
                g.apply_async(frames[k], lr)
                if i > 0:
                    last = g.sync_next()
        if mode != "sync":
            last = g.sync()
        return last, {n: g.state_get(n) for n in ("bg_color", "bg_desc", "T", "R", "v", "DminLT", "rawST", "finLT", "lastfg", "blinks", "unstable", "lut", "scalars")}

    m1, s1 = run("sync")
    m2, s2 = run("async")
    m3, s3 = run("sync")
    assert np.array_equal(m1, m2) and np.array_equal(m1, m3)
    for n in s1:
        assert np.array_equal(s1[n], s2[n]), f"sync vs async: {n}"
        assert np.array_equal(s1[n], s3[n]), f"run to run: {n}"


def test_subsense_host_operations_between_pipelined_frames(lv, oracle):
    """SuBSENSE leaves the neighbour writes of the latest frame queued for the next frame's scan and runs its mask chain on a side
    stream: host-side operations issued between frames (getBackgroundImage, refreshModel, state import, setAutomaticModelReset,
    a second instance interleaved) must see / produce exactly the oracle's state. No state is read between most frames, so the
    frames really are pipelined."""
    w, h, c = 320, 240, 3
    seq = SynthSequence(w, h, c, seed=33)
    g, o = _mk(lv, oracle, "subsense", seed=9)
    g2, o2 = _mk(lv, oracle, "subsense", seed=10)   # an unrelated instance interleaved on the same device
    f0 = seq.frame(0)
    for a in (g, o, g2, o2):
        a.initialize(f0)
    for t in range(1, 25):
        f = seq.frame(t)
        lr = 1.0 if t <= 6 else 0.0
        mg, mo = g.apply(f, lr), o.apply(f, lr)
        assert np.array_equal(mg, mo), f"frame {t}"
        if t % 3 == 0:
            assert np.array_equal(g2.apply(f, 0.0), o2.apply(f, 0.0))
        if t == 5:
            assert np.array_equal(g.getBackgroundImage(), o.get_background_image())          # flushes the queued neighbour writes
        if t == 9:
            g.refreshModel(0.5); o.refresh_model(0.5)                                        # host refresh with writes still queued
        if t == 12:
            g.setAutomaticModelReset(False); o.set_auto_model_reset(False)
        if t == 15:
            fg = o.state_get("lastfg").copy(); fg[::7] = 255
            g.state_set("lastfg", fg); o.state_set("lastfg", fg)                             # import into the ping-pong planes
            r = o.state_get("R").copy(); r[::5] += 0.25
            g.state_set("R", r); o.state_set("R", r)                                         # also refreshes the compact R plane
        if t == 18:
            assert np.array_equal(g.getBackgroundDescriptorsImage(), o.get_background_descriptors_image())
            g.refreshModel(1.0, True); o.refresh_model(1.0, True)
    _compare_state(g, o, INT_STATE, FLT_STATE, "host operations, end")
    _compare_state(g2, o2, INT_STATE, FLT_STATE, "interleaved instance, end")


@pytest.mark.parametrize("algo,w,h,c,n", [("subsense", 320, 240, 3, 130), ("lobster", 320, 240, 1, 80)])
def test_tier3_sequence(lv, oracle, algo, w, h, c, n):
    """end-to-end masks over a longer sequence with the samples/changedet learning-rate protocol (main.cpp:56)"""
    seq = SynthSequence(w, h, c, seed=1 if algo == "subsense" else 2)
    g, o = _mk(lv, oracle, algo, seed=0)
    f0 = seq.frame(0)
    g.initialize(f0)
    o.initialize(f0)
    worst, fm_g, fm_o = 0.0, [], []
    for t in range(1, n):
        f, gt = seq.frame(t, with_gt=True)
        lr = (1.0 if t <= 50 else 0.0) if algo == "subsense" else 16.0
        mg, mo = g.apply(f, lr), o.apply(f, lr)
        worst = max(worst, float((mg != mo).mean()))
        if t > 60:
            for m, acc in ((mg, fm_g), (mo, fm_o)):
                tp = ((m > 0) & gt).sum(); fp = ((m > 0) & ~gt).sum(); fn = ((m == 0) & gt).sum()
                acc.append(2 * tp / max(2 * tp + fp + fn, 1))
    assert worst <= TOL_MASK, f"per-pixel mask disagreement {worst}"
    if fm_g:
        assert abs(np.mean(fm_g) - np.mean(fm_o)) <= 1e-9
    flts = FLT_STATE if algo == "subsense" else []
    for nme in flts:
        a, b = g.state_get(nme), o.state_get(nme)
        assert np.allclose(a, b, rtol=FLOAT_RTOL, atol=1e-7), nme


def test_full_size_properties(lv):
    """1080p (BASELINE config #4 shape): size-independent properties instead of the (slow) oracle"""
    w, h = 1920, 1080
    seq = SynthSequence(w, h, 3, seed=4)
    g = lv.BackgroundSubtractorSuBSENSE(seed=0)
    f0 = seq.frame(0)
    g.initialize(f0)
    sc = g.state_get("scalars")
    assert sc[5] == 0 and sc[6] == 13  # 5x5 spread, median 13 (SuBSENSE.cpp:115-117)
    g2 = lv.BackgroundSubtractorSuBSENSE(seed=0)
    g2.initialize(f0)
    for t in range(1, 6):
        f, gt = seq.frame(t, with_gt=True)
        m, m2 = g.apply(f, 1.0), g2.apply(f, 1.0)
        assert np.array_equal(m, m2)                       # determinism (Philox, ordered neighbour writes)
        assert set(np.unique(m)) <= {0, 255}
        assert not m[:2].any() and not m[-2:].any() and not m[:, :2].any() and not m[:, -2:].any()  # 2-px border stays 0
    # a static scene stays background: feed the same frame repeatedly
    for _ in range(3):
        m = g.apply(f, 1.0)
    assert (m > 0).mean() < 0.02
    # the moving objects were detected at some point
    assert (gt & (m2 > 0)).sum() > 0.3 * gt.sum()


def test_async_batch_and_device_paths(lv, oracle):
    import ctypes
    w, h = 160, 120
    seqs = [SynthSequence(w, h, 3, seed=20 + i) for i in range(3)]
    subs = [lv.BackgroundSubtractorSuBSENSE(seed=i) for i in range(3)]
    refs = [lv.BackgroundSubtractorSuBSENSE(seed=i) for i in range(3)]
    for s, r, q in zip(subs, refs, seqs):
        s.initialize(q.frame(0)); r.initialize(q.frame(0))
    for t in range(1, 5):
        frames = [q.frame(t) for q in seqs]
        masks = lv.apply_batch(subs, frames, 1.0)
        for r, f, m in zip(refs, frames, masks):
            assert np.array_equal(r.apply(f, 1.0), m)
    s, r = subs[0], refs[0]
    f = seqs[0].frame(5)
    s.apply_async(f, 0.0)
    assert np.array_equal(s.sync(), r.apply(f, 0.0))


@pytest.mark.parametrize("algo", [0, 1, 2, 3, 4])
def test_cpp_drop_in_matches_python_path(lv, tmp_path, algo):
    """the header-only C++ drop-in (include/litiv_b200.hpp: reference class and method names over the C ABI) produces the same
    masks as the Python mirror on the same frames and seed"""
    import subprocess
    from litiv_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "shim_gpu"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp_shim_gpu.cpp"),
                           "-L" + os.path.dirname(build.SO), "-llitiv_b200", "-Wl,-rpath," + os.path.dirname(build.SO), "-o", str(exe)])
    w, h, c, n = 160, 120, 3, 12
    seq = SynthSequence(w, h, c, seed=61)
    frames = np.stack([seq.frame(t) for t in range(n)])
    frames.tofile(tmp_path / "in.raw")
    out = subprocess.run([str(exe), str(algo), str(w), str(h), str(c), str(n), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ran on GPU" in out.stdout, out.stdout + out.stderr
    got = np.fromfile(tmp_path / "out.raw", np.uint8).reshape(n - 1, h, w)
    cls = [lv.BackgroundSubtractorLOBSTER, lv.BackgroundSubtractorSuBSENSE, lv.BackgroundSubtractorPAWCS, lv.BackgroundSubtractorViBe_3ch,
           lv.BackgroundSubtractorPBAS_3ch][algo]
    g = cls(seed=5)
    g.initialize(frames[0])
    for t in range(1, n):
        lr = g.getDefaultLearningRate() if algo in (0, 3, 4) else (1.0 if t <= 5 else g.getDefaultLearningRate())
        assert np.array_equal(g.apply(frames[t], lr), got[t - 1]), f"algo {algo}, frame {t}"


def test_batch_device_enqueue_pool(lv):
    """lvb_apply_batch_device: many instances fed with device-resident frames from a pool of host threads; every stream must end
    up in exactly the state of an instance driven alone through the host API"""
    torch = pytest.importorskip("torch")
    w, h, n = 160, 120, 12
    seqs = [SynthSequence(w, h, 3, seed=40 + i) for i in range(n)]
    subs = [lv.BackgroundSubtractorSuBSENSE(seed=i) for i in range(n)]
    refs = [lv.BackgroundSubtractorSuBSENSE(seed=i) for i in range(n)]
    for s, r, q in zip(subs, refs, seqs):
        s.initialize(q.frame(0)); r.initialize(q.frame(0))
    pitch = (w * 3 + 127) // 128 * 128
    d_frames = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
    d_masks = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    batch = lv.DeviceBatch(subs)
    want = None
    for t in range(1, 9):
        frames = [q.frame(t) for q in seqs]
        for i, f in enumerate(frames):
            d_frames[i, :, :w * 3] = torch.from_numpy(f.reshape(h, w * 3)).cuda()
        torch.cuda.synchronize()
        batch.apply([d_frames[i].data_ptr() for i in range(n)], pitch, [d_masks[i].data_ptr() for i in range(n)], 1.0 if t < 5 else 0.0)
        want = [r.apply(f, 1.0 if t < 5 else 0.0) for r, f in zip(refs, frames)]
        for s in subs:
            s.sync()
        got = d_masks.cpu().numpy()
        for i in range(n):
            assert np.array_equal(got[i], want[i]), f"stream {i}, frame {t}"
    for s, r in zip(subs, refs):
        for name in ("bg_color", "bg_desc", "R", "T", "lastfg"):
            assert np.array_equal(s.state_get(name), r.state_get(name)), name
    with pytest.raises(lv.LitivError):
        batch.apply([0] * n, pitch, None, 1.0)   # null frame pointer


@pytest.mark.parametrize("pinned", [False, True])
def test_two_deep_pipeline_matches_synchronous_apply(lv, pinned):
    """lvb_apply_async keeps two frames in flight (upload of k+1 overlaps the kernels of k); masks come back in order and are
    identical to the synchronous call's"""
    w, h = 320, 240
    seq = SynthSequence(w, h, 3, seed=31)
    a, b = lv.BackgroundSubtractorSuBSENSE(seed=4), lv.BackgroundSubtractorSuBSENSE(seed=4)
    a.initialize(seq.frame(0)); b.initialize(seq.frame(0))
    n = 12
    frames = [seq.frame(t) for t in range(1, n + 1)]
    want = [b.apply(f, 1.0 if t < 6 else 0.0) for t, f in enumerate(frames)]
    if pinned:
        bufs = [lv.pinned_empty((h, w, 3)) for _ in range(2)]
        outs = [lv.pinned_empty((h, w)) for _ in range(2)]
    got = []
    for t, f in enumerate(frames):
        if pinned:
            bufs[t % 2][...] = f
            a.apply_async(bufs[t % 2], 1.0 if t < 6 else 0.0, out=outs[t % 2])
        else:
            a.apply_async(f, 1.0 if t < 6 else 0.0)
        if t >= 1:
            got.append(a.sync_next().copy())
    got.append(a.sync_next().copy())
    assert len(got) == n
    for t in range(n):
        assert np.array_equal(got[t], want[t]), f"frame {t}"
    with pytest.raises(lv.LitivError, match="no frame in flight"):
        a.sync_next()
    a.apply_async(frames[0]); a.apply_async(frames[1])
    with pytest.raises(lv.LitivError, match="already in flight"):
        a.apply_async(frames[2])
    a.sync()


def test_errors_match_reference_messages(lv):
    s = lv.BackgroundSubtractorSuBSENSE()
    with pytest.raises(lv.LitivError, match="initialized first"):
        s.apply(np.zeros((10, 10, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="initialized first"):
        s.getBackgroundImage()
    s.initialize(np.zeros((40, 40, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="mismatch"):
        s.apply(np.zeros((41, 40, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="0 or 255"):
        s.initialize(np.zeros((40, 40, 3), np.uint8), np.full((40, 40), 7, np.uint8))
    with pytest.raises(lv.LitivError, match="no useful pixels"):
        s.initialize(np.zeros((40, 40, 3), np.uint8), np.zeros((40, 40), np.uint8))
    lob = lv.BackgroundSubtractorLOBSTER()
    lob.initialize(np.zeros((40, 40), np.uint8))
    with pytest.raises(lv.LitivError, match="positive"):
        lob.apply(np.zeros((40, 40), np.uint8), 0.0)
    with pytest.raises(lv.LitivError, match="more sample matches"):
        lv.BackgroundSubtractorSuBSENSE(nBGSamples=2, nRequiredBGSamples=3)


def _shapes(h, w, rng):
    """nasty masks for the hole filler: nested rings, a spiral, random noise, blobs touching the border zone"""
    m = np.zeros((h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for r0, r1 in ((30, 34), (18, 22), (6, 10)):
        d = np.hypot(yy - h // 2, xx - w // 3)
        m[(d >= r0) & (d < r1)] = 255
    # square spiral
    x0, y0, x1, y1 = w // 2 + 4, 6, w - 6, h - 6
    while x1 - x0 > 8 and y1 - y0 > 8:
        m[y0, x0:x1] = 255; m[y0:y1, x1 - 1] = 255; m[y1 - 1, x0 + 4:x1] = 255; m[y0 + 4:y1, x0 + 4] = 255
        x0 += 4; y0 += 4; x1 -= 4; y1 -= 4
        m[y0, x0:x1 - 4] = 255
    m[rng.rand(h, w) < 0.02] = 255
    m[:2] = 0; m[-2:] = 0; m[:, :2] = 0; m[:, -2:] = 0
    return m


@pytest.mark.parametrize("shape", [(96, 130), (240, 320), (61, 75), (120, 33), (1080, 1920)])
def test_mask_ops_match_oracle(lv, oracle, shape):
    import ctypes as C
    h, w = shape
    rng = np.random.RandomState(h * 7 + w)
    L = oracle.lib()
    for name, m in (("shapes", _shapes(h, w, rng)), ("noise", ((rng.rand(h, w) < 0.45) * 255).astype(np.uint8)), ("empty", np.zeros((h, w), np.uint8))):
        if name != "shapes":
            m[:2] = 0; m[-2:] = 0; m[:, :2] = 0; m[:, -2:] = 0
        for op, r, dil in ((lv.MASK_DILATE, 1, 1), (lv.MASK_DILATE, 3, 1), (lv.MASK_ERODE, 1, 0), (lv.MASK_ERODE, 3, 0)):
            want = np.empty_like(m)
            L.lvo_morph_rect(m.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), w, h, r, dil)
            assert np.array_equal(lv.mask_op(op, m, r), want), (name, op, r)
        for k in (3, 9, 13):
            want = np.empty_like(m)
            L.lvo_median_binary(m.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), w, h, k)
            assert np.array_equal(lv.mask_op(lv.MASK_MEDIAN, m, k), want), (name, "median", k)
        flooded = m.copy()
        L.lvo_floodfill_origin(flooded.ctypes.data_as(C.c_void_p), w, h)
        got = lv.mask_op(lv.MASK_HOLES, m)
        assert np.array_equal(got, 255 - flooded), (name, "holes", int((got != 255 - flooded).sum()))
