"""Frame ingest: CDnet sequence layout, precaching reader, the sandbox's evaluation loop.
Reference: modules/datasets/include/litiv/datasets/impl/CDnet.hpp:66-110, utils.hpp:456-486, apps/changedet/src/main.cpp:346-428."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _datasets():
    # imported lazily: the package loads liblitiv_b200.so (built by __graft_entry__.build(), no GPU needed to load it)
    from litiv_b200 import datasets
    return datasets


def test_cdnet_layout_and_precacher(tmp_path):
    D = _datasets()
    d = D.write_synthetic_cdnet(str(tmp_path), "highway", 96, 72, 9, seed=3)
    seq = D.CDnetSequence(d)
    assert len(seq) == 9 and seq.frame_size == (72, 96) and seq.channels == 3 and not seq.grayscale
    assert os.path.basename(seq.input_paths[0]) == "in000001.jpg" and os.path.basename(seq.gt_paths[-1]) == "gt000009.png"
    assert seq.getOutputName(0) == "bin000001" and seq.getOutputName(122) == "bin000123"        # CDnet.hpp:104-110
    assert set(np.unique(seq.roi)) == {0, 255} and seq.roi[:7].max() == 0                        # m_oInputROI = oROI>0
    f = seq.getInput(4)
    assert f.shape == (72, 96, 3) and f.dtype == np.uint8
    assert np.array_equal(f, cv2.imread(seq.input_paths[4], cv2.IMREAD_COLOR))                  # the reference's own decoder
    assert set(np.unique(seq.getGT(4))) <= {0, 85, 170, 255}
    th = D.CDnetSequence(D.write_synthetic_cdnet(str(tmp_path), "corridor", 64, 48, 3, category="thermal"))
    assert th.grayscale and th.channels == 1 and th.getInput(1).shape == (48, 64)                # CDnet.hpp:72
    # error cases of parseData
    os.remove(seq.gt_paths[-1])
    with pytest.raises(Exception, match="same amount of GT"):
        D.CDnetSequence(d)
    os.remove(os.path.join(d, "ROI.jpg"))
    with pytest.raises(Exception):
        D.CDnetSequence(d)


@pytest.mark.gpu
@pytest.mark.parametrize("precache", [True, False])
def test_analyze_matches_oracle_loop(lv, oracle, tmp_path, precache):
    """the whole loop (decode -> initialize with the sequence ROI -> apply with the sandbox's learning-rate protocol -> on-device
    BinClassif) against the same loop driven through the oracle"""
    D = _datasets()
    d = D.write_synthetic_cdnet(str(tmp_path), "highway", 320, 240, 24, seed=5)
    seq = D.CDnetSequence(d)
    out = D.analyze(seq, lv.BackgroundSubtractorSuBSENSE(seed=4), evaluate=True, output_dir=str(tmp_path / "results"), precache=precache,
                    init_frames=10, keep_masks=True)
    o = oracle.Oracle(oracle.ALGO_SUBSENSE, mode=oracle.MODE_SNAPSHOT, seed=4)
    o.initialize(seq.getInput(0), seq.roi)
    want = np.zeros(6, np.uint64)
    for i in range(len(seq)):
        m = o.apply(seq.getInput(i), 1.0 if i <= 10 else 0.0)
        assert np.array_equal(out["masks"][i], m), f"frame {i}"
        want = oracle.binclassif(m, seq.getGT(i), seq.roi, counters=want)
        saved = cv2.imread(str(tmp_path / "results" / (seq.getOutputName(i) + ".png")), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(saved, m)
    assert np.array_equal(out["counters"], want)
    mo = oracle.binclassif_metrics(want)
    assert all(abs(out["metrics"][k] - mo[k]) <= 1e-15 for k in mo)
    assert out["frames"] == 24 and out["hz"] > 0
    # throughput mode: no evaluation, no masks kept, two frames in flight
    fast = D.analyze(seq, lv.BackgroundSubtractorSuBSENSE(seed=4), evaluate=False, precache=precache)
    assert fast["frames"] == 24 and fast["metrics"] is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vibe", "pbas"])
def test_analyze_runs_vibe_and_pbas(lv, oracle, tmp_path, name):
    """the same loop for the two plain cv::BackgroundSubtractor classes (no ROI at initialize; masks scored from the host copy)"""
    D = _datasets()
    d = D.write_synthetic_cdnet(str(tmp_path), "highway", 320, 240, 16, seed=6)
    seq = D.CDnetSequence(d)
    if name == "vibe":
        g, o = lv.BackgroundSubtractorViBe_3ch(seed=4), oracle.ViBeOracle(3, mode=oracle.MODE_SNAPSHOT, seed=4)
    else:
        g, o = lv.BackgroundSubtractorPBAS_3ch(seed=4), oracle.PBASOracle(3, mode=oracle.MODE_SNAPSHOT, seed=4)
    out = D.analyze(seq, g, evaluate=True, init_frames=5, keep_masks=True)
    o.initialize(seq.getInput(0))
    want = np.zeros(6, np.uint64)
    for i in range(len(seq)):
        m = o.apply(seq.getInput(i), 1.0 if i <= 5 else g.getDefaultLearningRate())
        assert np.array_equal(out["masks"][i], m), f"frame {i}"
        want = oracle.binclassif(m, seq.getGT(i), seq.roi, counters=want)
    assert np.array_equal(out["counters"], want)


def test_mat_binary_archive_round_trip_and_reference_file(tmp_path):
    """lv::write / lv::read MatArchive_BINARY (modules/utils/src/opencv.cpp:514-531, 608-625): the round trip of
    modules/utils/test/opencv.cpp:659-669 for every cv depth, the byte layout of a known header, and -- in the build container, where the
    reference tree exists -- the reference's own archive modules/features2d/test/data/test_lbsp.bin against the committed fixture"""
    import struct
    from litiv_b200.datasets import read_mat_binary, write_mat_binary
    rng = np.random.default_rng(3)
    for dt in (np.uint8, np.int8, np.uint16, np.int16, np.int32, np.float32, np.float64):
        for shape in ((rng.integers(100, 200), rng.integers(100, 200)), (17, 9, 3), (5, 4, 1), (1, 1)):
            a = (rng.uniform(-200, 200, shape)).astype(dt)
            p = str(tmp_path / "m.bin")
            write_mat_binary(p, a)
            b = read_mat_binary(p)
            want = a[..., 0] if a.ndim == 3 and a.shape[2] == 1 else a
            assert b.dtype == a.dtype and np.array_equal(b, want)
    p = str(tmp_path / "h.bin")
    write_mat_binary(p, np.arange(65 * 65 * 3, dtype=np.uint16).reshape(65, 65, 3))
    raw = open(p, "rb").read()
    assert struct.unpack("<iQQiii", raw[:32]) == (18, 6, 65 * 65, 2, 65, 65) and len(raw) == 32 + 65 * 65 * 6   # CV_16UC3 == 18
    vol = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    write_mat_binary(p, vol, channels_last=False)                                                              # a 3-d single-channel Mat
    assert struct.unpack("<iQQi", open(p, "rb").read()[:24]) == (5, 4, 24, 3) and np.array_equal(read_mat_binary(p), vol)
    with pytest.raises(ValueError):
        open(p, "wb").write(b"\\x00" * 10)
        read_mat_binary(p)
    ref = "/root/reference/modules/features2d/test/data/test_lbsp.bin"
    if os.path.exists(ref):
        gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbsp_golden.npz"))["desc"]
        got = read_mat_binary(ref)
        assert got.dtype == np.uint16 and np.array_equal(got, gold)
        write_mat_binary(p, gold)
        assert open(p, "rb").read() == open(ref, "rb").read()      # byte-identical to what the reference wrote
