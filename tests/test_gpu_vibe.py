"""GPU parity tests of ViBe (run with -m gpu): the CUDA path through the C ABI (lvb_vibe_*) against the CPU
oracle (oracle/lvo_vibe.hpp, snapshot mode, same Philox seed). Integer work only: masks and the whole sample model must be
bit-exact after every frame."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = pytest.mark.gpu


def _pair(lv, oracle, ch, seed, **kw):
    cls = lv.BackgroundSubtractorViBe_1ch if ch == 1 else lv.BackgroundSubtractorViBe_3ch
    okw = dict(color_dist_threshold=kw.get("nColorDistThreshold", 20), n_samples=kw.get("nBGSamples", 20), n_required=kw.get("nRequiredBGSamples", 2))
    return cls(seed=seed, **kw), oracle.ViBeOracle(ch, mode=oracle.MODE_SNAPSHOT, seed=seed, **okw)


@pytest.mark.parametrize("shape", [(48, 64, 3), (48, 64, 1), (37, 53, 3), (9, 7, 1), (1, 40, 3), (33, 1, 1), (5, 130, 3), (240, 320, 3)])
def test_masks_and_model_bit_exact_every_frame(lv, oracle, shape):
    h, w, ch = shape
    seq = SynthSequence(w, h, ch, seed=21)
    g, o = _pair(lv, oracle, ch, seed=7)
    g.set_collect_stats(True)
    f0 = seq.frame(0)
    g.initialize(f0); o.initialize(f0)
    assert np.array_equal(g.model(), o.model()), "model after initialize"
    for t in range(1, 25):
        f = seq.frame(t)
        lr = 1.0 if t < 4 else (16.0 if t % 5 else 2.0)
        mg, mo = g.apply(f, lr), o.apply(f, lr)
        assert np.array_equal(mg, mo), f"mask differs at frame {t}: {(mg != mo).sum()} px"
        assert np.array_equal(g.model(), o.model()), f"model differs after frame {t}"
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())
    assert g.stats() == o.stats()


def test_non_default_parameters_and_wrap_quirk_colours(lv, oracle):
    """N / #min / threshold other than the defaults; saturated colours far apart exercise the uint16 wrap of lv::L2dist"""
    rng = np.random.default_rng(3)
    g, o = _pair(lv, oracle, 3, seed=5, nColorDistThreshold=35, nBGSamples=7, nRequiredBGSamples=3)
    base = rng.integers(0, 256, (40, 72, 3), dtype=np.uint8)
    g.initialize(base); o.initialize(base)
    for t in range(12):
        f = base.copy()
        m = rng.random((40, 72)) < 0.3
        f[m] = rng.integers(0, 256, (int(m.sum()), 3), dtype=np.uint8)   # random far colours: some wrap into a "match"
        assert np.array_equal(g.apply(f, 3.0), o.apply(f, 3.0)), t
        assert np.array_equal(g.model(), o.model()), t


def test_gray_frames_into_the_3ch_model(lv, oracle):
    seq = SynthSequence(80, 60, 1, seed=8)
    g, o = _pair(lv, oracle, 3, seed=2)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    for t in range(1, 10):
        f = seq.frame(t) if t % 2 else np.repeat(seq.frame(t)[..., None], 3, axis=2)   # 8UC1 and 8UC3 frames may alternate
        assert np.array_equal(g.apply(f), o.apply(f))
    assert np.array_equal(g.model(), o.model())
    for t, lr in enumerate((1000.0, float("inf"), 1.0, 0.5), start=10):   # large, "never", and fractional (ceil -> 1) learning rates
        f = seq.frame(t)
        assert np.array_equal(g.apply(f, lr), o.apply(f, lr)), (t, lr)
    assert np.array_equal(g.model(), o.model())
    with pytest.raises(lv.LitivError):
        lv.BackgroundSubtractorViBe_1ch().initialize(np.zeros((8, 8, 3), np.uint8))


def test_model_import_continues_like_the_oracle(lv, oracle):
    """classification from an identical model snapshot: evolve the oracle, import its model, both must continue identically"""
    seq = SynthSequence(96, 72, 3, seed=30)
    g, o = _pair(lv, oracle, 3, seed=11)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    for t in range(1, 15):
        o.apply(seq.frame(t))
    g.set_model(o.model(), frame_idx=14)
    for t in range(15, 22):
        assert np.array_equal(g.apply(seq.frame(t)), o.apply(seq.frame(t)))
    assert np.array_equal(g.model(), o.model())


def test_errors_mirror_the_reference_asserts(lv):
    v = lv.BackgroundSubtractorViBe_3ch()
    with pytest.raises(lv.LitivError, match="initialized"):
        v.apply(np.zeros((8, 8, 3), np.uint8))
    v.initialize(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="learning rate"):
        v.apply(np.zeros((8, 8, 3), np.uint8), 0.0)
    with pytest.raises(lv.LitivError, match="sample"):
        lv.BackgroundSubtractorViBe_3ch(nBGSamples=2, nRequiredBGSamples=3)


def test_full_hd_device_resident_path_matches_oracle(lv, oracle):
    """BASELINE size (1920x1080 RGB) through lvb_vibe_apply_device: frames and masks stay in HBM"""
    import torch
    seq = SynthSequence(1920, 1080, 3, seed=4)
    g, o = _pair(lv, oracle, 3, seed=1)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    d_mask = torch.empty((1080, 1920), dtype=torch.uint8, device="cuda")
    for t in range(1, 7):
        f = seq.frame(t)
        d = torch.from_numpy(f).cuda()
        torch.cuda.synchronize()
        g.apply_device(d.data_ptr(), 3, 1920 * 3, d_mask.data_ptr(), 16.0)
        g.sync()
        assert np.array_equal(d_mask.cpu().numpy(), o.apply(f, 16.0)), t
    assert np.array_equal(g.model(), o.model())
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())


def test_cpu_restatement_timed_beside_the_gpu(lv, oracle, capsys):
    """the reference-order CPU restatement on one host core beside the CUDA path on the same 1080p RGB frames (a reported baseline,
    printed with -s; the only assertion is that the GPU path is not the slower one)"""
    import json
    import time
    seq = SynthSequence(1920, 1080, 3, seed=4100)
    frames = [seq.frame(t) for t in range(5)]
    o = oracle.ViBeOracle(3, mode=oracle.MODE_REFERENCE, seed=1)
    o.initialize(frames[0])
    t0 = time.perf_counter()
    for f in frames[1:]:
        o.apply(f, 16.0)
    cpu_s = (time.perf_counter() - t0) / 4
    g = lv.BackgroundSubtractorViBe_3ch(seed=1)
    g.initialize(frames[0])
    hf = [lv.pinned_empty(frames[0].shape) for _ in frames]
    for a, b in zip(hf, frames):
        a[...] = b
    hm = lv.pinned_empty((1080, 1920))
    for k in range(10):
        g.apply(hf[k % 5], 16.0, out=hm)
    t0 = time.perf_counter()
    for k in range(40):
        g.apply(hf[k % 5], 16.0, out=hm)
    gpu_s = (time.perf_counter() - t0) / 40
    line = {"algo": "vibe", "frame": [1920, 1080, 3], "cpu_baseline": {"value": 1920 * 1080 / cpu_s / 1e6, "unit": "Mpx/s", "cores": 1, "kind": "port",
            "sample": "4 frames, oracle reference-order mode"}, "gpu_end_to_end": {"value": 1920 * 1080 / gpu_s / 1e6, "unit": "Mpx/s",
            "api": "synchronous apply, host frame -> host mask, pinned buffers"}}
    with capsys.disabled():
        print("\n" + json.dumps(line))
    assert gpu_s < cpu_s


def test_interleaved_instances_are_independent(lv, oracle):
    """several ViBe / PBAS instances fed in turn (one CUDA stream and one Philox key each, no shared rand()) end in exactly the state
    of the same instance driven alone; the reference's instances perturb each other through libc rand()"""
    seqs = [SynthSequence(96, 64, 3, seed=50 + i) for i in range(3)]
    mk = [lambda s: lv.BackgroundSubtractorViBe_3ch(seed=s), lambda s: lv.BackgroundSubtractorPBAS_3ch(seed=s)]
    for make in mk:
        together = [make(i) for i in range(3)]
        for g, q in zip(together, seqs):
            g.initialize(q.frame(0))
        masks = [[] for _ in range(3)]
        for t in range(1, 12):
            for i, (g, q) in enumerate(zip(together, seqs)):
                masks[i].append(g.apply(q.frame(t)))
        for i, q in enumerate(seqs):
            alone = make(i)
            alone.initialize(q.frame(0))
            for t in range(1, 12):
                assert np.array_equal(alone.apply(q.frame(t)), masks[i][t - 1]), (i, t)
            assert np.array_equal(alone.getBackgroundImage(), together[i].getBackgroundImage())
