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
