"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/_ref/liblitiv_ref.so: the reference's OWN hot-path sources (modules/video/src/BackgroundSubtractor*.cpp,
modules/features2d/src/LBSP.cpp and the litiv/utils headers), compiled unmodified against oracle/cvcompat by `make -C oracle _ref`.
Used to pin the oracle's reference-order mode bit-for-bit (tests/test_ref_pin_cpu.py) and as bench.py's CPU arm.
The reference draws from the process-global libc rand(): every Reference() seeds it (srand) at construction, so only one instance
may be driven at a time within a process.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import Params, STATE_DTYPES

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liblitiv_ref.so")
_LIB = None
REFERENCE_ROOT = os.environ.get("LITIV_REFERENCE_ROOT", "/root/reference")


def available():
    """the prebuilt library exists, or the reference tree is present so that it can be built"""
    return os.path.exists(_SO) or os.path.isdir(REFERENCE_ROOT)


def build():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref", f"REF={REFERENCE_ROOT}"])
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/liblitiv_ref.so is missing and the reference tree is not present to build it")
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ref_last_error.restype = C.c_char_p
        _LIB.ref_version.restype = C.c_char_p
        _LIB.ref_apply_sequence.restype = C.c_double
        _LIB.ref_cdist3.restype = C.c_uint64
    return _LIB


class ReferenceError_(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise ReferenceError_(lib().ref_last_error().decode())


class Reference:
    """IBackgroundSubtractor of the reference itself (algo ids as in the oracle: 0 LOBSTER, 1 SuBSENSE, 2 PAWCS)"""

    def __init__(self, algo, seed=0, params=None):
        self._h = C.c_void_p()
        self.algo = algo
        _chk(lib().ref_create(algo, C.byref(params) if params is not None else None, C.c_uint(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.ref_destroy(self._h)
            self._h = None

    def initialize(self, img, roi=None):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        rp = None
        if roi is not None:
            roi = np.ascontiguousarray(roi, dtype=np.uint8)
            rp = roi.ctypes.data_as(C.c_void_p)
        _chk(lib().ref_initialize(self._h, img.ctypes.data_as(C.c_void_p), w, h, c, rp))
        self.shape = (h, w, c)

    def apply(self, img, lr=0.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = self.shape
        assert img.size == h * w * c
        mask = np.empty((h, w), np.uint8)
        _chk(lib().ref_apply(self._h, img.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), C.c_double(lr)))
        return mask

    def apply_sequence(self, frames, lrs):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        h, w, c = self.shape
        lrs = np.ascontiguousarray(lrs, dtype=np.float64)
        mask = np.empty((h, w), np.uint8)
        t = lib().ref_apply_sequence(self._h, frames.ctypes.data_as(C.c_void_p), frames.shape[0], C.c_size_t(h * w * c),
                                     mask.ctypes.data_as(C.c_void_p), lrs.ctypes.data_as(C.c_void_p))
        if t < 0:
            raise ReferenceError_(lib().ref_last_error().decode())
        return t, mask

    def refresh_model(self, frac, force_fg=False):
        _chk(lib().ref_refresh_model(self._h, C.c_float(frac), int(force_fg)))

    def pawcs_refresh_model(self, base_occ, decr_frac, force_fg=False):
        _chk(lib().ref_pawcs_refresh_model(self._h, C.c_uint64(base_occ), C.c_float(decr_frac), int(force_fg)))

    def set_auto_model_reset(self, v):
        _chk(lib().ref_set_auto_model_reset(self._h, int(v)))

    def get_background_image(self):
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint8)
        _chk(lib().ref_get_background_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if c == 1 else out

    def get_background_descriptors_image(self):
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint16)
        _chk(lib().ref_get_background_descriptors_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if c == 1 else out

    def get_roi(self):
        h, w, _ = self.shape
        out = np.empty((h, w), np.uint8)
        _chk(lib().ref_get_roi(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def state_get(self, name):
        n = C.c_size_t()
        _chk(lib().ref_state_get(self._h, name.encode(), None, C.byref(n)))
        dt = np.uint8 if name == "lw_valid" else STATE_DTYPES[name]
        out = np.empty(n.value // np.dtype(dt).itemsize, dt)
        _chk(lib().ref_state_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out


def lbsp_compute(img, ref=None, rel=None, thr=0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    out = np.zeros((h, w, c), np.uint16)
    rp = None
    if ref is not None:
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        rp = ref.ctypes.data_as(C.c_void_p)
    _chk(lib().ref_lbsp_compute(img.ctypes.data_as(C.c_void_p), rp, w, h, c, int(rel is not None), C.c_float(rel if rel is not None else 0.0),
                                int(thr), out.ctypes.data_as(C.c_void_p)))
    return out[..., 0] if c == 1 else out


class ReferenceViBe:
    """BackgroundSubtractorViBe_1ch / _3ch of the reference itself (video/src/BackgroundSubtractorViBe.cpp, compiled unmodified)"""

    def __init__(self, model_channels=3, color_dist_threshold=20, n_samples=20, n_required=2, seed=0):
        self._h = C.c_void_p()
        self.C, self.N = model_channels, n_samples
        _chk(lib().ref_vibe_create(model_channels, color_dist_threshold, n_samples, n_required, C.c_uint(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.ref_vibe_destroy(self._h)
            self._h = None

    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        return img, (1 if img.ndim == 2 else img.shape[2])

    def initialize(self, img):
        img, c = self._img(img)
        h, w = img.shape[:2]
        _chk(lib().ref_vibe_initialize(self._h, img.ctypes.data_as(C.c_void_p), w, h, c))
        self.shape = (h, w)

    def apply(self, img, lr=16.0):
        img, c = self._img(img)
        mask = np.empty(self.shape, np.uint8)
        _chk(lib().ref_vibe_apply(self._h, img.ctypes.data_as(C.c_void_p), c, mask.ctypes.data_as(C.c_void_p), C.c_double(lr)))
        return mask

    def model(self):
        out = np.empty((self.N,) + self.shape + (self.C,), np.uint8)
        _chk(lib().ref_vibe_model(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out

    def get_background_image(self):
        out = np.empty(self.shape + (self.C,), np.uint8)
        _chk(lib().ref_vibe_get_background_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if self.C == 1 else out


class ReferencePBAS:
    """BackgroundSubtractorPBAS_1ch / _3ch of the reference itself (video/src/BackgroundSubtractorPBAS.cpp, compiled unmodified)"""
    STATE = {"bg_color": np.uint8, "bg_grad": np.uint8, "R": np.float32, "T": np.float32, "meanmin": np.float32, "scalars": np.float64}

    def __init__(self, model_channels=3, color_dist_threshold=30, update_rate=16.0, n_samples=35, n_required=2, seed=0):
        self._h = C.c_void_p()
        self.C, self.N = model_channels, n_samples
        _chk(lib().ref_pbas_create(model_channels, color_dist_threshold, C.c_float(update_rate), n_samples, n_required, C.c_uint(seed), C.byref(self._h)))
        self.shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.ref_pbas_destroy(self._h)
            self._h = None

    def initialize(self, img):
        img, c = ReferenceViBe._img(img)
        h, w = img.shape[:2]
        _chk(lib().ref_pbas_initialize(self._h, img.ctypes.data_as(C.c_void_p), w, h, c))
        self.shape = (h, w)

    def apply(self, img, lr=-1.0):
        img, c = ReferenceViBe._img(img)
        mask = np.empty(self.shape, np.uint8)
        _chk(lib().ref_pbas_apply(self._h, img.ctypes.data_as(C.c_void_p), c, mask.ctypes.data_as(C.c_void_p), C.c_double(lr)))
        return mask

    def state_get(self, name):
        shape = (self.N,) + self.shape + (self.C,) if name in ("bg_color", "bg_grad") else (2,) if name == "scalars" else self.shape
        out = np.empty(shape, self.STATE[name])
        _chk(lib().ref_pbas_state_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out

    def get_background_image(self):
        out = np.empty(self.shape + (self.C,), np.uint8)
        _chk(lib().ref_pbas_get_background_image(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[..., 0] if self.C == 1 else out


class ReferenceEdgeDetectorLBSP:
    """EdgeDetectorLBSP of the reference itself (imgproc/src/EdgeDetectorLBSP.cpp, compiled unmodified)"""

    def __init__(self, levels=3, hyst_low_factor=0.5, normalize_output=False):
        self._h = C.c_void_p()
        _chk(lib().ref_edge_create(levels, C.c_double(hyst_low_factor), int(bool(normalize_output)), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.ref_edge_destroy(self._h)
            self._h = None

    def apply_threshold(self, img, thr=0.5):
        img, c = ReferenceViBe._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().ref_edge_apply_threshold(self._h, img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p), C.c_double(thr)))
        return out

    def apply(self, img):
        img, c = ReferenceViBe._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().ref_edge_apply(self._h, img.ctypes.data_as(C.c_void_p), w, h, c, out.ctypes.data_as(C.c_void_p)))
        return out

    def raw(self, which):
        """the detector's persistent buffers as they are: 0 = gradient map (4 bytes per cell, padded), 1 = edge mask (padded)"""
        n = C.c_size_t(0)
        _chk(lib().ref_edge_raw(self._h, which, None, C.byref(n)))
        out = np.empty(n.value, np.uint8)
        _chk(lib().ref_edge_raw(self._h, which, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

