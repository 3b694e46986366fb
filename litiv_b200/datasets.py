"""Frame ingest for the change-detection hot path: the CDnet sequence layout, a precaching
reader and the evaluation loop that feeds `apply()` and scores its masks.

Mirrors, in Python over the C ABI:
  * `DataProducer_<..., Dataset_CDnet>::parseData`  modules/datasets/include/litiv/datasets/impl/CDnet.hpp:66-99
    (`input/`, `groundtruth/`, `ROI.bmp` + `ROI.jpg`, grayscale for "thermal" / "turbulence", one GT file per frame)
  * the output naming `bin%06zu` (1-based)            CDnet.hpp:104-110
  * `DataPrecacher` (decode ahead of the consumer)    modules/datasets/include/litiv/datasets/utils.hpp:456-486
  * `Analyze()` of the reference's sandbox            apps/changedet/src/main.cpp:346-428 (learning rate 1 for the first 101
    frames, then the algorithm's default; every mask pushed to the evaluator)

Decoding stays on the host with OpenCV's `imread` (the codec the reference itself uses), so the frames `apply()` sees are
bit-identical to the reference's; a hardware JPEG decoder is not (different IDCT / chroma upsampling), which would void the
parity claim at the first kernel. Frames are decoded ahead into page-locked buffers and go through `lvb_apply_async`
(two frames in flight: upload of k+1 beside the kernels of k); masks are scored where they live with the on-device
`BinClassif` (no read-back needed unless the caller wants the masks).
"""
import os
import threading
import time

import numpy as np

from . import api


def _cv2():
    import cv2
    return cv2


class CDnetSequence:
    """one CDnet sequence directory (e.g. `dataset/baseline/highway/`)"""

    def __init__(self, path, name=None):
        cv2 = _cv2()
        self.path = os.path.join(path, "")
        self.name = name or os.path.basename(os.path.normpath(path))
        # CDnet.hpp:72: the category is part of the relative path
        self.grayscale = ("thermal" in self.path) or ("turbulence" in self.path)
        in_dir, gt_dir = os.path.join(self.path, "input"), os.path.join(self.path, "groundtruth")
        if not os.path.isdir(in_dir) or not os.path.isdir(gt_dir):
            raise api.LitivError(f"CDnet sequence '{self.name}' at '{self.path}' did not possess the required groundtruth and input directories")
        self.input_paths = sorted(os.path.join(in_dir, f) for f in os.listdir(in_dir) if os.path.isfile(os.path.join(in_dir, f)))
        self.gt_paths = sorted(os.path.join(gt_dir, f) for f in os.listdir(gt_dir) if os.path.isfile(os.path.join(gt_dir, f)))
        if not self.input_paths:
            raise api.LitivError("could not find any input frames")
        if len(self.gt_paths) != len(self.input_paths):
            raise api.LitivError(f"CDnet sequence '{self.name}' did not possess same amount of GT & input frames")
        roi = cv2.imread(os.path.join(self.path, "ROI.bmp"), cv2.IMREAD_GRAYSCALE)
        troi = cv2.imread(os.path.join(self.path, "ROI.jpg"))
        if roi is None or troi is None:
            raise api.LitivError(f"CDnet sequence '{self.name}' did not possess ROI.bmp/ROI.jpg files")
        if roi.shape[:2] != troi.shape[:2]:   # CDnet.hpp:87-90: keep the smallest overlap
            roi = roi[:min(roi.shape[0], troi.shape[0]), :min(roi.shape[1], troi.shape[1])].copy()
        self.roi = np.where(roi > 0, 255, 0).astype(np.uint8)   # m_oInputROI = oROI>0 ; the GT ROI is the same mask
        self.frame_size = self.roi.shape                       # (rows, cols)
        self.channels = 1 if self.grayscale else 3

    def __len__(self):
        return len(self.input_paths)

    def getInput(self, idx):
        cv2 = _cv2()
        img = cv2.imread(self.input_paths[idx], cv2.IMREAD_GRAYSCALE if self.grayscale else cv2.IMREAD_COLOR)
        if img is None:
            raise api.LitivError(f"could not read input frame {self.input_paths[idx]}")
        if img.shape[:2] != self.frame_size:   # frames larger than the ROI overlap are cropped like the ROI
            img = np.ascontiguousarray(img[:self.frame_size[0], :self.frame_size[1]])
        return img

    def getGT(self, idx):
        cv2 = _cv2()
        gt = cv2.imread(self.gt_paths[idx], cv2.IMREAD_GRAYSCALE)
        if gt is None:
            raise api.LitivError(f"could not read groundtruth frame {self.gt_paths[idx]}")
        if gt.shape[:2] != self.frame_size:
            gt = np.ascontiguousarray(gt[:self.frame_size[0], :self.frame_size[1]])
        return gt

    @staticmethod
    def getOutputName(idx):
        return "bin%06d" % (idx + 1)


class DataPrecacher:
    """decodes packets ahead of the consumer on worker threads (`cv2.imread` releases the GIL) into a ring of page-locked
    buffers; `get(idx)` returns (input, gt_or_None), `release(idx)` (in consumer order) hands the slot back"""

    def __init__(self, seq, with_gt=True, depth=12, workers=None):
        self.seq, self.with_gt, self.depth = seq, with_gt, max(2, depth)
        workers = workers or max(2, min(8, (os.cpu_count() or 4) // 2))
        h, w = seq.frame_size
        shape = (h, w) if seq.channels == 1 else (h, w, seq.channels)
        self._bufs = [api.pinned_empty(shape) for _ in range(self.depth)]
        self._gts = [None] * self.depth
        self._ready = set()       # packet indices decoded and not yet released
        self._released = 0        # packets [0, _released) have been consumed: packet idx may use its slot once idx < _released + depth
        self._next = 0
        self._cv = threading.Condition()
        self._stop = False
        self._err = None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()

    def _work(self):
        n = len(self.seq)
        while True:
            with self._cv:
                if self._stop or self._next >= n:
                    return
                idx = self._next
                self._next += 1
                while idx >= self._released + self.depth and not self._stop:
                    self._cv.wait(0.05)
                if self._stop:
                    return
            slot = idx % self.depth
            try:
                self._bufs[slot][...] = self.seq.getInput(idx)
                self._gts[slot] = self.seq.getGT(idx) if self.with_gt else None
            except Exception as e:  # surfaced by get()
                with self._cv:
                    self._err = e
                    self._cv.notify_all()
                return
            with self._cv:
                self._ready.add(idx)
                self._cv.notify_all()

    def get(self, idx):
        with self._cv:
            while idx not in self._ready:
                if self._err is not None:
                    raise self._err
                self._cv.wait(0.05)
        slot = idx % self.depth
        return self._bufs[slot], self._gts[slot]

    def release(self, idx):
        """the consumer is done with packet idx (its upload has completed): the slot may be reused"""
        with self._cv:
            self._ready.discard(idx)
            self._released = max(self._released, idx + 1)
            self._cv.notify_all()

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        for t in self._threads:
            t.join(timeout=2)


def analyze(seq, algo, evaluate=True, output_dir=None, precache=True, init_frames=100, keep_masks=False):
    """apps/changedet/src/main.cpp:346-428 for one sequence. `algo`: a constructed background subtractor of this package.
    Returns dict(frames, seconds, hz, counters (BinClassif), metrics, masks (if keep_masks))."""
    cv2 = _cv2()
    n = len(seq)
    if n <= 1:
        raise api.LitivError("a sequence needs more than one frame")
    default_lr = algo.getDefaultLearningRate()
    pre = DataPrecacher(seq, with_gt=evaluate) if precache else None
    get = (lambda i: pre.get(i)) if pre else (lambda i: (seq.getInput(i), seq.getGT(i) if evaluate else None))
    first, _ = get(0)
    # ViBe / PBAS are plain cv::BackgroundSubtractor classes in the reference: initialize(img) without a ROI, synchronous apply, and
    # their masks are scored from the host copy (with the sequence ROI, as the evaluator does for every algorithm)
    lbsp_family = isinstance(algo, api._BackgroundSubtractor)
    if lbsp_family:
        algo.initialize(np.array(first, copy=True), seq.roi)
    else:
        algo.initialize(np.array(first, copy=True))
    counters = api.BinClassif(device=getattr(algo, "device", 0))
    masks = []
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
    need_mask = bool(output_dir) or keep_masks
    t0 = time.perf_counter()
    pending = None  # (idx, gt): frame whose mask is in flight
    # two frames in flight when the masks need no host-side work between frames; scoring on the device reads the instance's
    # latest mask, so with evaluation on each frame is collected before the next one is submitted
    for idx in range(n):
        img, gt = get(idx)
        lr = 1.0 if idx <= init_frames else default_lr          # main.cpp:393
        if evaluate or need_mask or not lbsp_family:
            mask = algo.apply(img, lr)
            if evaluate:
                counters.accumulate(algo if lbsp_family else mask, gt, seq.roi)   # oBatch.push -> evaluator (BinClassif::accumulate with the GT ROI)
            if output_dir:
                cv2.imwrite(os.path.join(output_dir, seq.getOutputName(idx) + ".png"), mask)
            if keep_masks:
                masks.append(mask.copy())
            if pre:
                pre.release(idx)
        else:
            algo.apply_async(img, lr)
            if pending is not None:
                algo.sync_next()
                if pre:
                    pre.release(pending)
            pending = idx
    if pending is not None:
        algo.sync()
        if pre:
            pre.release(pending)
    dt = time.perf_counter() - t0
    if pre:
        pre.close()
    out = dict(frames=n, seconds=dt, hz=n / dt, counters=counters.counters.copy(), metrics=counters.metrics() if evaluate else None)
    if keep_masks:
        out["masks"] = masks
    return out


def write_synthetic_cdnet(root, name, width, height, nframes, seed=1, category="baseline", jpeg_quality=92):
    """a CDnet-shaped sequence directory generated from `synth.SynthSequence` (tests, demos, benchmarks: there is no network
    to fetch the real dataset). Ground truth uses the CDnet labels: 0, 255, 85 outside the ROI, 170 on object outlines."""
    cv2 = _cv2()
    from .synth import SynthSequence
    d = os.path.join(root, category, name)
    os.makedirs(os.path.join(d, "input"), exist_ok=True)
    os.makedirs(os.path.join(d, "groundtruth"), exist_ok=True)
    gray = category in ("thermal", "turbulence")
    seq = SynthSequence(width, height, 1 if gray else 3, seed=seed)
    roi = np.full((height, width), 255, np.uint8)
    roi[:height // 10] = 0
    cv2.imwrite(os.path.join(d, "ROI.bmp"), roi)
    cv2.imwrite(os.path.join(d, "ROI.jpg"), roi)
    with open(os.path.join(d, "temporalROI.txt"), "w") as f:
        f.write(f"1 {nframes}\n")
    k = np.ones((3, 3), np.uint8)
    for t in range(nframes):
        img, fg = seq.frame(t, with_gt=True)
        cv2.imwrite(os.path.join(d, "input", "in%06d.jpg" % (t + 1)), img, [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
        gt = np.where(fg, 255, 0).astype(np.uint8)
        edge = cv2.dilate(gt, k) != cv2.erode(gt, k)
        gt[edge] = 170
        gt[roi == 0] = 85
        cv2.imwrite(os.path.join(d, "groundtruth", "gt%06d.png" % (t + 1)), gt)
    return d


# --- lv::write / lv::read, MatArchive_BINARY (modules/utils/src/opencv.cpp:514-531, 608-625): the archive format of the reference's
# descriptor dumps (modules/features2d/test/data/test_lbsp.bin is one): int32 cv type, uint64 element size, uint64 element count,
# int32 dims, int32 size[dims], raw row-major data; little-endian, no padding. Host-side I/O next to the path.
_CV_DEPTHS = [np.uint8, np.int8, np.uint16, np.int16, np.int32, np.float32, np.float64]   # CV_8U .. CV_64F = 0 .. 6


def write_mat_binary(path, arr, channels_last=None):
    """writes `arr` as the reference's lv::write(path, mat, lv::MatArchive_BINARY) would write the equivalent cv::Mat. A 3-d array whose
    last axis has at most 4 entries is taken as rows x cols x channels (what cv2 hands out); pass channels_last=False to store a
    genuinely 3-dimensional single-channel Mat instead."""
    import struct
    arr = np.ascontiguousarray(arr)
    try:
        depth = [np.dtype(d) for d in _CV_DEPTHS].index(arr.dtype)
    except ValueError:
        raise ValueError(f"dtype {arr.dtype} has no cv::Mat depth") from None
    if channels_last is None:
        channels_last = arr.ndim == 3 and arr.shape[2] <= 4
    if arr.ndim < 2:
        arr = arr.reshape(-1, 1) if arr.ndim == 1 else arr.reshape(1, 1)     # cv::Mat has at least two dimensions
    cn = arr.shape[-1] if channels_last else 1
    sizes = arr.shape[:-1] if channels_last else arr.shape
    if not 1 <= cn <= 512:
        raise ValueError("cv::Mat supports 1..512 channels")
    total = int(np.prod(sizes, dtype=np.int64))
    with open(path, "wb") as f:
        f.write(struct.pack("<iQQi", depth + ((cn - 1) << 3), arr.dtype.itemsize * cn, total, len(sizes)))
        f.write(struct.pack("<%di" % len(sizes), *sizes))
        f.write(arr.tobytes())


def read_mat_binary(path):
    """lv::read(path, lv::MatArchive_BINARY): returns the array (rows x cols [x channels] for 2-d Mats, dims... [x channels] otherwise)"""
    import struct
    raw = open(path, "rb").read()
    if len(raw) < 24:
        raise ValueError("binary archive read failed")
    mtype, esz, total, dims = struct.unpack("<iQQi", raw[:24])
    if dims < 1 or dims > 32 or len(raw) < 24 + 4 * dims:
        raise ValueError("binary archive read failed")
    sizes = struct.unpack("<%di" % dims, raw[24:24 + 4 * dims])
    depth, cn = mtype & 7, (mtype >> 3) + 1
    if depth >= len(_CV_DEPTHS):
        raise ValueError("unsupported cv::Mat depth in archive")
    dt = np.dtype(_CV_DEPTHS[depth])
    if esz != dt.itemsize * cn or total != int(np.prod(sizes, dtype=np.int64)) or len(raw) < 24 + 4 * dims + esz * total:
        raise ValueError("binary archive read failed")
    data = np.frombuffer(raw, dt, count=total * cn, offset=24 + 4 * dims)
    return data.reshape(tuple(sizes) + ((cn,) if cn > 1 else ())).copy()
