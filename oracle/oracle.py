"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liblvo_oracle.so (the CPU restatement of the reference's hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALGO_LOBSTER, ALGO_SUBSENSE, ALGO_PAWCS = 0, 1, 2
MODE_REFERENCE, MODE_SNAPSHOT = 0, 1


class Params(C.Structure):
    _fields_ = [("rel_lbsp_threshold", C.c_float), ("lbsp_threshold_offset", C.c_int), ("desc_dist_threshold", C.c_int),
                ("color_dist_threshold", C.c_int), ("n_samples", C.c_int), ("n_required", C.c_int),
                ("n_samples_for_moving_avgs", C.c_int), ("n_global_words", C.c_int), ("median_blur_kernel_size", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liblvo_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.lvo_last_error.restype = C.c_char_p
        _LIB.lvo_apply_sequence.restype = C.c_double
        _LIB.lvo_cdist3.restype = C.c_uint64
        _LIB.lvo_cdist4.restype = C.c_uint64
        _LIB.lvo_cdist2.restype = C.c_uint64
    return _LIB


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise OracleError(lib().lvo_last_error().decode())


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


STATE_DTYPES = {
    "roi": np.uint8, "lastfg": np.uint8, "lastcolor": np.uint8, "lastdesc": np.uint16, "lut": np.uint8,
    "bg_color": np.uint8, "bg_desc": np.uint16, "T": np.float32, "R": np.float32, "v": np.float32,
    "Dlast": np.float32, "DminLT": np.float32, "DminST": np.float32, "rawLT": np.float32, "rawST": np.float32,
    "finLT": np.float32, "finST": np.float32, "dsLT": np.float32, "dsST": np.float32, "unstable": np.uint8,
    "blinks": np.uint8, "lastraw": np.uint8, "lastrawblink": np.uint8, "dilinv": np.uint8, "rawmask": np.uint8,
    "scalars": np.float64,
    # PAWCS
    "illum": np.uint8, "dil": np.uint8, "lw_first": np.uint32, "lw_last": np.uint32, "lw_occ": np.uint32, "lw_color": np.uint8,
    "lw_desc": np.uint16, "gw_weight": np.float32, "gw_map": np.float32, "gw_bits": np.uint8, "gw_color": np.uint8,
    "gw_desc": np.uint16, "gdict": np.int32, "glut": np.uint8,
}
SUBSENSE_STATE = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "T", "R", "v", "Dlast", "DminLT",
                  "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST", "unstable", "blinks", "lastraw",
                  "lastrawblink", "dilinv", "scalars"]
PAWCS_STATE = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "T", "R", "v", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST",
               "dsLT", "dsST", "unstable", "illum", "blinks", "lastraw", "lastrawblink", "dil", "dilinv", "lw_first", "lw_last", "lw_occ",
               "lw_color", "lw_desc", "gw_weight", "gw_map", "gw_bits", "gw_color", "gw_desc", "gdict", "glut", "scalars"]
LOBSTER_STATE = ["roi", "lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "scalars"]


class Oracle:
    """Mirrors IBackgroundSubtractor (initialize / apply / getBackgroundImage) over the CPU restatement."""

    def __init__(self, algo, mode=MODE_SNAPSHOT, seed=0, params=None):
        self._h = C.c_void_p()
        self.algo = algo
        _chk(lib().lvo_create(algo, C.byref(params) if params is not None else None, mode, C.c_uint64(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().lvo_destroy(self._h)
            self._h = None

    def initialize(self, img, roi=None):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        rp = None
        if roi is not None:
            roi, rp = _u8(roi)
        _chk(lib().lvo_initialize(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, c, rp))
        self.shape = (h, w, c)

    def apply(self, img, lr=0.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = self.shape
        assert img.size == h * w * c
        mask = np.empty((h, w), np.uint8)
        _chk(lib().lvo_apply(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), mask.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_double(lr)))
        return mask

    def apply_sequence(self, frames, lrs):
        """frames: [n,H,W,(C)] uint8; returns (seconds inside apply, last mask)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n = frames.shape[0]
        h, w, c = self.shape
        lrs = np.ascontiguousarray(lrs, dtype=np.float64)
        mask = np.empty((h, w), np.uint8)
        t = lib().lvo_apply_sequence(self._h, frames.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.c_size_t(h * w * c),
                                     mask.ctypes.data_as(C.POINTER(C.c_uint8)), lrs.ctypes.data_as(C.POINTER(C.c_double)))
        if t < 0:
            raise OracleError(lib().lvo_last_error().decode())
        return t, mask

    def refresh_model(self, frac, force_fg=False):
        _chk(lib().lvo_refresh_model(self._h, C.c_float(frac), int(force_fg)))

    def pawcs_refresh_model(self, base_occ, decr_frac, force_fg=False):
        _chk(lib().lvo_pawcs_refresh_model(self._h, C.c_uint64(base_occ), C.c_float(decr_frac), int(force_fg)))

    def set_auto_model_reset(self, v):
        _chk(lib().lvo_set_auto_model_reset(self._h, int(v)))

    def get_background_image(self):
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint8)
        _chk(lib().lvo_get_background_image(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out[..., 0] if c == 1 else out

    def get_background_descriptors_image(self):
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint16)
        _chk(lib().lvo_get_background_descriptors_image(self._h, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out[..., 0] if c == 1 else out

    def state_get(self, name):
        n = C.c_size_t()
        _chk(lib().lvo_state_size(self._h, name.encode(), C.byref(n)))
        out = np.empty(n.value // np.dtype(STATE_DTYPES[name]).itemsize, STATE_DTYPES[name])
        _chk(lib().lvo_state_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n))
        return out

    def state_set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=STATE_DTYPES[name])
        _chk(lib().lvo_state_set(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)))

    def stats(self):
        out = (C.c_uint64 * 5)()
        lib().lvo_get_stats(self._h, out)
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])

    def scan_hist(self):
        """histogram of the per-pixel scan depth accumulated since initialize (bin 63: >= 63)"""
        out = (C.c_uint64 * 64)()
        lib().lvo_get_scan_hist(self._h, out)
        return np.array(list(out), dtype=np.uint64)


class ViBeOracle:
    """BackgroundSubtractorViBe_1ch / _3ch (video/src/BackgroundSubtractorViBe.cpp) over the CPU restatement (oracle/lvo_vibe.hpp)"""

    def __init__(self, model_channels=3, color_dist_threshold=20, n_samples=20, n_required=2, mode=MODE_SNAPSHOT, seed=0):
        self._h = C.c_void_p()
        self.C, self.N = model_channels, n_samples
        _chk(lib().lvo_vibe_create(model_channels, color_dist_threshold, n_samples, n_required, mode, C.c_uint64(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.lvo_vibe_destroy(self._h)
            self._h = None

    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        return img, (1 if img.ndim == 2 else img.shape[2])

    def initialize(self, img):
        img, c = self._img(img)
        h, w = img.shape[:2]
        _chk(lib().lvo_vibe_initialize(self._h, img.ctypes.data_as(C.c_void_p), w, h, c))
        self.shape = (h, w)

    def apply(self, img, lr=16.0):
        img, c = self._img(img)
        assert img.shape[:2] == self.shape
        mask = np.empty(self.shape, np.uint8)
        _chk(lib().lvo_vibe_apply(self._h, img.ctypes.data_as(C.c_void_p), c, mask.ctypes.data_as(C.c_void_p), C.c_double(lr)))
        return mask

    def get_background_image(self):
        out = np.empty(self.shape + (self.C,), np.uint8)
        _chk(lib().lvo_vibe_get_background_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if self.C == 1 else out

    def model(self):
        """samples [N][H][W][C]"""
        out = np.empty((self.N,) + self.shape + (self.C,), np.uint8)
        _chk(lib().lvo_vibe_model(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes), 0))
        return out

    def set_model(self, arr, frame_idx=None):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        _chk(lib().lvo_vibe_model(self._h, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes), 1))
        if frame_idx is not None:
            lib().lvo_vibe_set_frame(self._h, C.c_uint64(frame_idx))

    def stats(self):
        out = (C.c_uint64 * 5)()
        lib().lvo_vibe_get_stats(self._h, out)
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])


PBAS_STATE = {"bg_color": np.uint8, "bg_grad": np.uint8, "R": np.float32, "T": np.float32, "meanmin": np.float32, "rawmask": np.uint8,
              "lastgrad": np.uint8, "scalars": np.float64}


class PBASOracle:
    """BackgroundSubtractorPBAS_1ch / _3ch (video/src/BackgroundSubtractorPBAS.cpp) over the CPU restatement (oracle/lvo_pbas.hpp)"""

    def __init__(self, model_channels=3, color_dist_threshold=30, update_rate=16.0, n_samples=35, n_required=2, mode=MODE_SNAPSHOT, seed=0):
        self._h = C.c_void_p()
        self.C, self.N = model_channels, n_samples
        _chk(lib().lvo_pbas_create(model_channels, color_dist_threshold, C.c_float(update_rate), n_samples, n_required, mode,
                                   C.c_uint64(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.lvo_pbas_destroy(self._h)
            self._h = None

    def initialize(self, img):
        img, c = ViBeOracle._img(img)
        h, w = img.shape[:2]
        _chk(lib().lvo_pbas_initialize(self._h, img.ctypes.data_as(C.c_void_p), w, h, c))
        self.shape = (h, w)

    def apply(self, img, lr=-1.0):
        img, c = ViBeOracle._img(img)
        assert img.shape[:2] == self.shape
        mask = np.empty(self.shape, np.uint8)
        _chk(lib().lvo_pbas_apply(self._h, img.ctypes.data_as(C.c_void_p), c, mask.ctypes.data_as(C.c_void_p), C.c_double(lr)))
        return mask

    def get_background_image(self):
        out = np.empty(self.shape + (self.C,), np.uint8)
        _chk(lib().lvo_pbas_get_background_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if self.C == 1 else out

    def _shape_of(self, name):
        if name in ("bg_color", "bg_grad"):
            return (self.N,) + self.shape + (self.C,)
        if name == "lastgrad":
            return self.shape + (self.C,)
        return (2,) if name == "scalars" else self.shape

    def state_get(self, name):
        out = np.empty(self._shape_of(name), PBAS_STATE[name])
        _chk(lib().lvo_pbas_state(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes), 0))
        return out

    def state_set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=PBAS_STATE[name])
        _chk(lib().lvo_pbas_state(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes), 1))

    def stats(self):
        out = (C.c_uint64 * 5)()
        lib().lvo_pbas_get_stats(self._h, out)
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])


def pbas_gradient_image(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    _chk(lib().lvo_pbas_gradient_image(img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p)))
    return out


class EdgeDetectorLBSPOracle:
    """EdgeDetectorLBSP (imgproc/src/EdgeDetectorLBSP.cpp) over the CPU restatement (oracle/lvo_edge_lbsp.hpp); like the reference
    object it keeps its gradient / mask buffers between calls"""

    def __init__(self, levels=3, hyst_low_factor=0.5, normalize_output=False):
        self._h = C.c_void_p()
        _chk(lib().lvo_edge_create(levels, C.c_double(hyst_low_factor), C.byref(self._h)))
        if normalize_output:
            lib().lvo_edge_set_normalize(self._h, 1)

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.lvo_edge_destroy(self._h)
            self._h = None

    def apply_threshold(self, img, thr=0.5):
        img, c = ViBeOracle._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().lvo_edge_apply_threshold(self._h, img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p), C.c_double(thr)))
        return out

    def apply(self, img):
        img, c = ViBeOracle._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().lvo_edge_apply(self._h, img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p)))
        return out

    def gradient_map(self, shape):
        h, w = shape[:2]
        out = np.empty((h, w, 4), np.uint8)
        _chk(lib().lvo_edge_gradient_map(self._h, w, h, out.ctypes.data_as(C.c_void_p)))
        return out

    def raw(self, which):
        """the persistent buffers as they are: 0 = gradient map (4 bytes per cell, padded by 2 on every side), 1 = edge mask (padded)"""
        n = C.c_size_t(0)
        _chk(lib().lvo_edge_raw(self._h, which, None, C.byref(n)))
        out = np.empty(n.value, np.uint8)
        _chk(lib().lvo_edge_raw(self._h, which, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out


def normalize_minmax_u8(a):
    """cv::normalize(a, a, 0, 255, NORM_MINMAX) for an 8-bit array (the restatement EdgeDetectorLBSP's normalised output uses)"""
    a = np.ascontiguousarray(a, np.uint8).copy()
    lib().lvo_normalize_minmax_u8(a.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
    return a


def vibe_match(model_channels, thr, a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return bool(lib().lvo_vibe_match(model_channels, thr, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))


def lbsp_compute(img, ref=None, rel=None, thr=0):
    """LBSP::compute2 (dense). rel=None -> absolute threshold `thr`; else relative `rel` with offset `thr`."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.zeros((h, w, c), np.uint16)
    rp = None
    if ref is not None:
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        assert ref.shape == img.shape
        rp = ref.ctypes.data_as(C.POINTER(C.c_uint8))
    _chk(lib().lvo_lbsp_compute(img.ctypes.data_as(C.POINTER(C.c_uint8)), rp, w, h, c, int(rel is not None),
                                C.c_float(rel if rel is not None else 0.0), int(thr), out.ctypes.data_as(C.POINTER(C.c_uint16))))
    return out[..., 0] if c == 1 else out


def lbsp_gradient(img):
    """dense LBSP::computeDescriptor_gradient map: [H][W][4] u8 = gradX (int8), gradY (int8), magnitude, 0"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.zeros((h, w, 4), np.uint8)
    _chk(lib().lvo_lbsp_gradient(img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p)))
    return out


def binclassif(classif, gt=None, roi=None, counters=None):
    """lv::BinClassif::accumulate (datasets/src/metrics.cpp:21-61); returns the six counters TP,TN,FP,FN,SE,DC (added to `counters`)"""
    classif = np.ascontiguousarray(classif, dtype=np.uint8)
    h, w = classif.shape
    out = np.zeros(6, np.uint64) if counters is None else np.ascontiguousarray(counters, dtype=np.uint64).copy()
    gp = rp = None
    if gt is not None:
        gt = np.ascontiguousarray(gt, dtype=np.uint8); gp = gt.ctypes.data_as(C.POINTER(C.c_uint8))
    if roi is not None:
        roi = np.ascontiguousarray(roi, dtype=np.uint8); rp = roi.ctypes.data_as(C.POINTER(C.c_uint8))
    _chk(lib().lvo_binclassif(classif.ctypes.data_as(C.POINTER(C.c_uint8)), gp, rp, w, h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def binclassif_metrics(c):
    """BinClassifMetrics (datasets/include/litiv/datasets/metrics.hpp:213-257), restated in Python floats (IEEE double)"""
    TP, TN, FP, FN = (int(v) for v in c[:4])
    rec = TP / (TP + FN) if TP + FN > 0 else 0.0
    pre = TP / (TP + FP) if TP + FP > 0 else 0.0
    ok = TP + FP > 0 and TP + FN > 0 and TN + FP > 0 and TN + FN > 0
    return dict(dRecall=rec, dSpecificity=TN / (TN + FP) if TN + FP > 0 else 0.0, dFPR=FP / (FP + TN) if FP + TN > 0 else 0.0,
                dFNR=FN / (TP + FN) if TP + FN > 0 else 0.0, dPBC=100.0 * (FN + FP) / (TP + TN + FP + FN) if TP + TN + FP + FN > 0 else 0.0,
                dPrecision=pre, dFMeasure=2.0 * (rec * pre) / (rec + pre) if rec + pre > 0 else 0.0,
                dMCC=((float(TP) * TN) - float(FP * FN)) / np.sqrt((float(TP) + FP) * (TP + FN) * (TN + FP) * (TN + FN)) if ok else 0.0)
