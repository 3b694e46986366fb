"""GPU parity tests of PBAS (run with -m gpu): the CUDA path through the C ABI (lvb_pbas_*) against the CPU oracle
(oracle/lvo_pbas.hpp, snapshot mode, same Philox seed). Masks, raw masks, gradient images and the sample model must be bit-exact;
the float maps R(x), T(x), mean-min-distance and m_fFormerMeanGradDist within FLOAT_RTOL (observed: bit-identical)."""
import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-5   # north_star: "float feedback maps within 1e-5 relative"


def _pair(lv, oracle, ch, seed, **kw):
    cls = lv.BackgroundSubtractorPBAS_1ch if ch == 1 else lv.BackgroundSubtractorPBAS_3ch
    okw = dict(color_dist_threshold=kw.get("nInitColorDistThreshold", 30), update_rate=kw.get("fInitUpdateRate", 16.0),
               n_samples=kw.get("nBGSamples", 35), n_required=kw.get("nRequiredBGSamples", 2))
    return cls(seed=seed, **kw), oracle.PBASOracle(ch, mode=oracle.MODE_SNAPSHOT, seed=seed, **okw)


def _compare(g, o, tag):
    for n in ("bg_color", "bg_grad", "rawmask", "lastgrad"):
        a, b = g.state_get(n), o.state_get(n)
        assert np.array_equal(a, b), f"{tag}: '{n}' differs in {(a != b).sum()} of {a.size} entries"
    for n in ("R", "T", "meanmin", "scalars"):
        a, b = g.state_get(n), o.state_get(n)
        assert np.allclose(a, b, rtol=FLOAT_RTOL, atol=0), f"{tag}: '{n}' max rel err {np.abs(a - b).max()}"


@pytest.mark.parametrize("shape", [(48, 64, 3), (48, 64, 1), (37, 53, 3), (9, 7, 1), (1, 40, 3), (33, 1, 1), (5, 130, 3), (2, 2, 3), (240, 320, 3)])
def test_masks_and_state_every_frame(lv, oracle, shape):
    h, w, ch = shape
    seq = SynthSequence(w, h, ch, seed=22)
    g, o = _pair(lv, oracle, ch, seed=9)
    g.set_collect_stats(True)
    f0 = seq.frame(0)
    g.initialize(f0); o.initialize(f0)
    assert np.array_equal(g.state_get("bg_color"), o.state_get("bg_color")) and np.array_equal(g.state_get("bg_grad"), o.state_get("bg_grad"))
    for t in range(1, 30):
        f = seq.frame(t)
        lr = 1.0 if t < 4 else (-1.0 if t % 5 else 3.0)
        mg, mo = g.apply(f, lr), o.apply(f, lr)
        assert np.array_equal(mg, mo), f"mask differs at frame {t}: {(mg != mo).sum()} px"
        _compare(g, o, f"frame {t}")
    assert np.array_equal(g.getBackgroundImage(), o.get_background_image())
    assert g.stats() == o.stats()


def test_float_maps_bit_identical_over_a_longer_run(lv, oracle):
    seq = SynthSequence(160, 120, 3, seed=31)
    g, o = _pair(lv, oracle, 3, seed=2)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    for t in range(1, 90):
        assert np.array_equal(g.apply(seq.frame(t)), o.apply(seq.frame(t))), t
    for n in ("R", "T", "meanmin", "scalars"):
        assert np.array_equal(g.state_get(n), o.state_get(n)), n
    _compare(g, o, "end")


def test_non_default_parameters_gray_frames_and_errors(lv, oracle):
    seq = SynthSequence(80, 60, 1, seed=8)
    g, o = _pair(lv, oracle, 3, seed=2, nInitColorDistThreshold=20, fInitUpdateRate=4.0, nBGSamples=11, nRequiredBGSamples=3)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    for t in range(1, 14):
        f = seq.frame(t) if t % 2 else np.repeat(seq.frame(t)[..., None], 3, axis=2)
        assert np.array_equal(g.apply(f), o.apply(f)), t
    _compare(g, o, "gray into 3ch")
    for t, lr in enumerate((1000.0, float("inf"), 257.0, 2.0), start=14):   # overrides beyond the in-kernel modulo table, and "never"
        f = seq.frame(t)
        assert np.array_equal(g.apply(f, lr), o.apply(f, lr)), (t, lr)
    _compare(g, o, "large learning-rate overrides")
    with pytest.raises(lv.LitivError):
        lv.BackgroundSubtractorPBAS_1ch().initialize(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="initialized"):
        lv.BackgroundSubtractorPBAS_3ch().apply(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(lv.LitivError, match="update rate"):
        lv.BackgroundSubtractorPBAS_3ch(fInitUpdateRate=0.0)


def test_state_import_continues_like_the_oracle(lv, oracle):
    seq = SynthSequence(96, 72, 3, seed=30)
    g, o = _pair(lv, oracle, 3, seed=11)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    for t in range(1, 20):
        o.apply(seq.frame(t))
    for n in ("bg_color", "bg_grad", "R", "T", "meanmin", "scalars"):
        g.state_set(n, o.state_get(n))
    for t in range(20, 28):
        assert np.array_equal(g.apply(seq.frame(t)), o.apply(seq.frame(t))), t
    _compare(g, o, "after import")


def test_full_hd_device_resident_path_matches_oracle(lv, oracle):
    import torch
    seq = SynthSequence(1920, 1080, 3, seed=4)
    g, o = _pair(lv, oracle, 3, seed=1)
    g.initialize(seq.frame(0)); o.initialize(seq.frame(0))
    d_mask = torch.empty((1080, 1920), dtype=torch.uint8, device="cuda")
    for t in range(1, 5):
        f = seq.frame(t)
        d = torch.from_numpy(f).cuda()
        torch.cuda.synchronize()
        g.apply_device(d.data_ptr(), 3, 1920 * 3, d_mask.data_ptr())
        g.sync()
        assert np.array_equal(d_mask.cpu().numpy(), o.apply(f)), t
    _compare(g, o, "1080p")


def test_cpu_restatement_timed_beside_the_gpu(lv, oracle, capsys):
    """the reference-order CPU restatement on one host core beside the CUDA path on the same 1080p RGB frames (a reported baseline,
    printed with -s; the only assertion is that the GPU path is not the slower one)"""
    import json
    import time
    seq = SynthSequence(1920, 1080, 3, seed=4100)
    frames = [seq.frame(t) for t in range(5)]
    o = oracle.PBASOracle(3, mode=oracle.MODE_REFERENCE, seed=1)
    o.initialize(frames[0])
    t0 = time.perf_counter()
    for f in frames[1:]:
        o.apply(f, -1.0)
    cpu_s = (time.perf_counter() - t0) / 4
    g = lv.BackgroundSubtractorPBAS_3ch(seed=1)
    g.initialize(frames[0])
    hf = [lv.pinned_empty(frames[0].shape) for _ in frames]
    for a, b in zip(hf, frames):
        a[...] = b
    hm = lv.pinned_empty((1080, 1920))
    for k in range(10):
        g.apply(hf[k % 5], -1.0, out=hm)
    t0 = time.perf_counter()
    for k in range(40):
        g.apply(hf[k % 5], -1.0, out=hm)
    gpu_s = (time.perf_counter() - t0) / 40
    line = {"algo": "pbas", "frame": [1920, 1080, 3], "cpu_baseline": {"value": 1920 * 1080 / cpu_s / 1e6, "unit": "Mpx/s", "cores": 1, "kind": "port",
            "sample": "4 frames, oracle reference-order mode"}, "gpu_end_to_end": {"value": 1920 * 1080 / gpu_s / 1e6, "unit": "Mpx/s",
            "api": "synchronous apply, host frame -> host mask, pinned buffers"}}
    with capsys.disabled():
        print("\n" + json.dumps(line))
    assert gpu_s < cpu_s
