"""Python mirror of the reference's IBackgroundSubtractor surface over the C ABI (include/litiv_b200.h).

Method names follow the reference (modules/video/include/litiv/video/BackgroundSubtractionUtils.hpp:24-47):
initialize(img, roi) / apply(img, learningRate) -> fgmask / getBackgroundImage() / refreshModel(...).
The library has no CPU fallback: if the CUDA extension is missing or no device is present, calls raise.
"""
import os as _os
# every instance drives several CUDA streams (upload, kernels, mask chain, auxiliary, read-back): with the default of 8 hardware
# queues, streams of concurrent instances alias and serialise. Must be set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import ctypes as C
import os
import numpy as np

from . import build as _build

ALGO_LOBSTER, ALGO_SUBSENSE, ALGO_PAWCS = 0, 1, 2
_LIB = None

EXPORTS = [
    "lvb_last_error", "lvb_device_count", "lvb_default_params", "lvb_create", "lvb_destroy", "lvb_initialize", "lvb_apply",
    "lvb_apply_async", "lvb_sync", "lvb_apply_batch", "lvb_apply_device", "lvb_get_background_image",
    "lvb_get_background_descriptors_image", "lvb_refresh_model", "lvb_set_auto_model_reset", "lvb_get_roi", "lvb_set_roi",
    "lvb_default_learning_rate", "lvb_lbsp_compute", "lvb_state_size", "lvb_state_get", "lvb_state_set",
    "lvb_set_collect_stats", "lvb_get_stats", "lvb_kernel_launch_count", "lvb_stream", "lvb_set_profile", "lvb_get_profile",
    "lvb_host_alloc", "lvb_host_free", "lvb_mask_op", "lvb_pawcs_refresh_model", "lvb_sync_next", "lvb_flush", "lvb_get_profile_feedback", "lvb_get_profile_tail", "lvb_apply_stream", "lvb_get_background_image_device", "lvb_validate_roi",
    "lvb_binclassif_accumulate", "lvb_binclassif", "lvb_binclassif_metrics", "lvb_apply_batch_device",
    "lvb_vibe_create", "lvb_vibe_destroy", "lvb_vibe_initialize", "lvb_vibe_apply", "lvb_vibe_apply_device", "lvb_vibe_sync",
    "lvb_vibe_get_background_image", "lvb_vibe_model", "lvb_vibe_set_collect_stats", "lvb_vibe_get_stats", "lvb_vibe_set_profile",
    "lvb_vibe_get_profile", "lvb_vibe_stream",
    "lvb_pbas_create", "lvb_pbas_destroy", "lvb_pbas_initialize", "lvb_pbas_apply", "lvb_pbas_apply_device", "lvb_pbas_sync",
    "lvb_pbas_get_background_image", "lvb_pbas_state", "lvb_pbas_set_collect_stats", "lvb_pbas_get_stats", "lvb_pbas_set_profile",
    "lvb_pbas_get_profile", "lvb_pbas_stream", "lvb_lbsp_gradient",
    "lvb_edge_create", "lvb_edge_destroy", "lvb_edge_set_normalize", "lvb_edge_default_threshold", "lvb_edge_apply_threshold", "lvb_edge_apply",
    "lvb_edge_get_gradient_map", "lvb_edge_flood_sweeps", "lvb_edge_apply_threshold_device", "lvb_edge_stream",
]


class LitivError(RuntimeError):
    """Raised where the reference would throw lv::Exception (utils/defines.hpp.in:109-113)."""


class Params(C.Structure):
    _fields_ = [("rel_lbsp_threshold", C.c_float), ("lbsp_threshold_offset", C.c_int32), ("desc_dist_threshold", C.c_int32),
                ("color_dist_threshold", C.c_int32), ("n_samples", C.c_int32), ("n_required", C.c_int32),
                ("n_samples_for_moving_avgs", C.c_int32), ("n_global_words", C.c_int32), ("median_blur_kernel_size", C.c_int32)]


def lib_path():
    return _build.SO


def lib():
    """Loads the in-tree CUDA extension; fails loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_build.SO):
            raise LitivError("litiv_b200/liblitiv_b200.so is missing: run `python -m litiv_b200.build` (needs nvcc); there is no CPU fallback")
        L = C.CDLL(_build.SO)
        L.lvb_last_error.restype = C.c_char_p
        L.lvb_default_learning_rate.restype = C.c_double
        L.lvb_kernel_launch_count.restype = C.c_uint64
        L.lvb_stream.restype = C.c_void_p
        L.lvb_stream.argtypes = [C.c_void_p]
        L.lvb_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        L.lvb_destroy.argtypes = [C.c_void_p]
        L.lvb_initialize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.lvb_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.lvb_apply_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.lvb_apply_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.lvb_get_background_image_device.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_validate_roi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.lvb_sync.argtypes = [C.c_void_p]
        L.lvb_sync_next.argtypes = [C.c_void_p]
        L.lvb_flush.argtypes = [C.c_void_p]
        L.lvb_binclassif_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_binclassif.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.lvb_binclassif_metrics.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_apply_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.lvb_apply_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double]
        L.lvb_apply_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double]
        L.lvb_get_background_image.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_get_background_descriptors_image.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_refresh_model.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.lvb_pawcs_refresh_model.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_int]
        L.lvb_set_auto_model_reset.argtypes = [C.c_void_p, C.c_int]
        L.lvb_get_roi.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_set_roi.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_state_size.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.lvb_state_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.lvb_state_set.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.lvb_set_collect_stats.argtypes = [C.c_void_p, C.c_int]
        L.lvb_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_lbsp_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
        L.lvb_default_params.argtypes = [C.c_int, C.c_void_p]
        L.lvb_lbsp_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.lvb_edge_create.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.lvb_edge_destroy.argtypes = [C.c_void_p]
        L.lvb_edge_set_normalize.argtypes = [C.c_void_p, C.c_int]
        L.lvb_edge_default_threshold.restype = C.c_double
        L.lvb_edge_apply_threshold.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double]
        L.lvb_edge_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.lvb_edge_get_gradient_map.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_edge_flood_sweeps.argtypes = [C.c_void_p]
        L.lvb_edge_flood_sweeps.restype = C.c_uint64
        L.lvb_edge_apply_threshold_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_double]
        L.lvb_edge_stream.argtypes = [C.c_void_p]
        L.lvb_edge_stream.restype = C.c_void_p
        L.lvb_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.lvb_get_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_get_profile_feedback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_get_profile_tail.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_host_alloc.argtypes = [C.c_void_p, C.c_size_t]
        L.lvb_host_free.argtypes = [C.c_void_p]
        L.lvb_mask_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.lvb_vibe_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.lvb_vibe_destroy.argtypes = [C.c_void_p]
        L.lvb_vibe_initialize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t]
        L.lvb_vibe_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double]
        L.lvb_vibe_apply_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_double]
        L.lvb_vibe_sync.argtypes = [C.c_void_p]
        L.lvb_vibe_get_background_image.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_vibe_model.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32]
        L.lvb_vibe_set_collect_stats.argtypes = [C.c_void_p, C.c_int]
        L.lvb_vibe_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_vibe_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.lvb_vibe_get_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_vibe_stream.restype = C.c_void_p
        L.lvb_vibe_stream.argtypes = [C.c_void_p]
        L.lvb_pbas_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.lvb_pbas_destroy.argtypes = [C.c_void_p]
        L.lvb_pbas_initialize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t]
        L.lvb_pbas_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double]
        L.lvb_pbas_apply_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_double]
        L.lvb_pbas_sync.argtypes = [C.c_void_p]
        L.lvb_pbas_get_background_image.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_pbas_state.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
        L.lvb_pbas_set_collect_stats.argtypes = [C.c_void_p, C.c_int]
        L.lvb_pbas_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.lvb_pbas_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.lvb_pbas_get_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvb_pbas_stream.restype = C.c_void_p
        L.lvb_pbas_stream.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise LitivError(lib().lvb_last_error().decode())


def device_count():
    return lib().lvb_device_count()


def default_params(algo):
    p = Params()
    _chk(lib().lvb_default_params(algo, C.byref(p)))
    return p


def kernel_launch_count():
    return int(lib().lvb_kernel_launch_count())


STATE_DTYPES = {
    "roi": np.uint8, "lastfg": np.uint8, "lastcolor": np.uint8, "lastdesc": np.uint16, "lut": np.uint8,
    "bg_color": np.uint8, "bg_desc": np.uint16, "T": np.float32, "R": np.float32, "v": np.float32,
    "Dlast": np.float32, "DminLT": np.float32, "DminST": np.float32, "rawLT": np.float32, "rawST": np.float32,
    "finLT": np.float32, "finST": np.float32, "dsLT": np.float32, "dsST": np.float32, "unstable": np.uint8,
    "blinks": np.uint8, "lastraw": np.uint8, "lastrawblink": np.uint8, "dilinv": np.uint8, "rawmask": np.uint8,
    "ghost": np.uint8, "scalars": np.float64,
    # PAWCS
    "illum": np.uint8, "dil": np.uint8, "lw_first": np.uint32, "lw_last": np.uint32, "lw_occ": np.uint32, "lw_color": np.uint8,
    "lw_desc": np.uint16, "gw_weight": np.float32, "gw_map": np.float32, "gw_bits": np.uint8, "gw_color": np.uint8,
    "gw_desc": np.uint16, "gdict": np.int32, "glut": np.uint8,
}


class _BackgroundSubtractor:
    """Common drop-in surface (IIBackgroundSubtractor + IBackgroundSubtractorLBSP)."""
    ALGO = None

    def __init__(self, params=None, device=0, seed=0):
        self._h = C.c_void_p()
        self._params = params
        _chk(lib().lvb_create(self.ALGO, C.byref(params) if params is not None else None, device, seed, C.byref(self._h)))
        self.shape = None
        self._pending_mask = None
        self._inflight = []

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.lvb_destroy(h)
            self._h = None

    # -- reference API -------------------------------------------------------------------------
    def initialize(self, img, roi=None):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or img.size == 0:
            raise LitivError("provided image for initialization must be non-empty, continuous, and of type 8UC1/3/4")
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        rp = None
        if roi is not None:
            roi = np.ascontiguousarray(roi)
            if roi.dtype != np.uint8 or roi.shape != (h, w):
                raise LitivError("provided ROI mat size must be equal to the init frame size, and its type must be 8UC1")
            rp = roi.ctypes.data
        _chk(lib().lvb_initialize(self._h, img.ctypes.data, w, h, c, w * c, rp))
        self.shape = (h, w, c)

    def _check_img(self, img):
        if self.shape is None:
            raise LitivError("algo & model must be initialized first")
        img = np.asarray(img)
        h, w, c = self.shape
        if img.dtype != np.uint8 or img.shape != ((h, w) if c == 1 and img.ndim == 2 else (h, w, c)):
            raise LitivError("input image type/size mismatch with initialization type/size")
        return np.ascontiguousarray(img)

    def _check_out(self, out):
        """caller-supplied mask buffer: the C ABI writes W*H bytes through the raw pointer, so it has to be exactly that"""
        if out is None:
            return np.empty(self.shape[:2], np.uint8)
        if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != tuple(self.shape[:2]) or not out.flags.c_contiguous or not out.flags.writeable:
            raise LitivError("output mask must be a writeable C-contiguous uint8 array of the frame size")
        return out

    def apply(self, img, learningRate=None, out=None):
        img = self._check_img(img)
        lr = self.getDefaultLearningRate() if learningRate is None else learningRate
        mask = self._check_out(out)
        _chk(lib().lvb_apply(self._h, img.ctypes.data, mask.ctypes.data, float(lr)))
        return mask

    def apply_stream(self, frames, learningRates, outs=None):
        """n consecutive frames of one stream through lvb_apply_stream (C loop over lvb_apply_async / lvb_sync_next, two frames in
        flight); returns the list of masks"""
        imgs = [self._check_img(f) for f in frames]
        masks = [self._check_out(None if outs is None else outs[i]) for i in range(len(imgs))]
        n = len(imgs)
        ip = (C.c_void_p * n)(*[f.ctypes.data for f in imgs])
        mp = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
        lrs = (C.c_double * n)(*[float(x) for x in learningRates])
        _chk(lib().lvb_apply_stream(self._h, ip, mp, n, lrs))
        return masks

    def apply_async(self, img, learningRate=None, out=None):
        """enqueue one frame (up to two may be in flight); collect the masks in order with sync_next() / sync()"""
        img = self._check_img(img)
        lr = self.getDefaultLearningRate() if learningRate is None else learningRate
        mask = self._check_out(out)
        _chk(lib().lvb_apply_async(self._h, img.ctypes.data, mask.ctypes.data, float(lr)))
        self._inflight.append((mask, img))   # keeps the buffers alive until collected

    def sync_next(self):
        """wait for the oldest frame in flight and return its mask"""
        _chk(lib().lvb_sync_next(self._h))
        return self._inflight.pop(0)[0]

    def sync(self):
        """collect every frame in flight; returns the newest mask"""
        _chk(lib().lvb_sync(self._h))
        m = self._inflight[-1][0] if self._inflight else None
        self._inflight = []
        return m

    def apply_device(self, d_img_ptr, d_step, d_mask_ptr=None, learningRate=None):
        """device-resident frame (raw CUDA pointers, e.g. torch tensor .data_ptr()); asynchronous on self.stream"""
        lr = self.getDefaultLearningRate() if learningRate is None else learningRate
        _chk(lib().lvb_apply_device(self._h, d_img_ptr, d_step, d_mask_ptr, float(lr)))

    def flush(self):
        """make self.stream wait for the side-stream work (mask chain) of every frame enqueued so far"""
        _chk(lib().lvb_flush(self._h))

    def getBackgroundImage(self):
        if self.shape is None:
            raise LitivError("algo must be initialized first")
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint8)
        _chk(lib().lvb_get_background_image(self._h, out.ctypes.data))
        return out[..., 0] if c == 1 else out

    def getBackgroundDescriptorsImage(self):
        if self.shape is None:
            raise LitivError("algo must be initialized first")
        h, w, c = self.shape
        out = np.empty((h, w, c), np.uint16)
        _chk(lib().lvb_get_background_descriptors_image(self._h, out.ctypes.data))
        return out[..., 0] if c == 1 else out

    def getDefaultLearningRate(self):
        return lib().lvb_default_learning_rate(self.ALGO)

    def setAutomaticModelReset(self, enabled):
        _chk(lib().lvb_set_auto_model_reset(self._h, int(bool(enabled))))

    def getROICopy(self):
        h, w, _ = self.shape
        out = np.empty((h, w), np.uint8)
        _chk(lib().lvb_get_roi(self._h, out.ctypes.data))
        return out

    def validateROI(self, roi):
        """IIBackgroundSubtractor::validateROI: clears the 2-px border of the (uint8, HxW) ROI in place and returns it"""
        if not isinstance(roi, np.ndarray) or roi.dtype != np.uint8 or roi.ndim != 2 or roi.size == 0 or not roi.flags.c_contiguous or not roi.flags.writeable:
            raise LitivError("provided ROI must be non-empty and of type 8UC1")
        _chk(lib().lvb_validate_roi(roi.ctypes.data, roi.shape[1], roi.shape[0], 2))
        return roi

    def getBackgroundImageDevice(self, d_out_ptr):
        """getBackgroundImage into device memory (raw CUDA pointer to W*H*C bytes), the cv::cuda::GpuMat form of the reference's display path"""
        _chk(lib().lvb_get_background_image_device(self._h, d_out_ptr))

    def setROI(self, roi):
        if self.shape is None:
            raise LitivError("algo & model must be initialized first")
        roi = np.ascontiguousarray(roi)
        if roi.dtype != np.uint8 or roi.shape != tuple(self.shape[:2]):
            raise LitivError("provided ROI mat size must be equal to the init frame size, and its type must be 8UC1")
        _chk(lib().lvb_set_roi(self._h, roi.ctypes.data))

    def refreshModel(self, fSamplesRefreshFrac, bForceFGUpdate=False):
        _chk(lib().lvb_refresh_model(self._h, float(fSamplesRefreshFrac), int(bool(bForceFGUpdate))))

    # -- parity / instrumentation --------------------------------------------------------------
    @property
    def stream(self):
        return lib().lvb_stream(self._h)

    def state_get(self, name):
        n = C.c_size_t()
        _chk(lib().lvb_state_size(self._h, name.encode(), C.byref(n)))
        out = np.empty(n.value // np.dtype(STATE_DTYPES[name]).itemsize, STATE_DTYPES[name])
        _chk(lib().lvb_state_get(self._h, name.encode(), out.ctypes.data, n.value))
        return out

    def state_set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=STATE_DTYPES[name])
        _chk(lib().lvb_state_set(self._h, name.encode(), arr.ctypes.data, arr.nbytes))

    def set_profile(self, enabled):
        _chk(lib().lvb_set_profile(self._h, int(bool(enabled))))

    def get_profile(self):
        ms, n = C.c_double(), C.c_uint64()
        _chk(lib().lvb_get_profile(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def get_profile_feedback(self):
        ms, n = C.c_double(), C.c_uint64()
        _chk(lib().lvb_get_profile_feedback(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def get_profile_tail(self):
        ms, n = C.c_double(), C.c_uint64()
        _chk(lib().lvb_get_profile_tail(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_collect_stats(self, enabled):
        _chk(lib().lvb_set_collect_stats(self._h, int(bool(enabled))))

    def stats(self):
        out = (C.c_uint64 * 5)()
        _chk(lib().lvb_get_stats(self._h, out))
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])


class BackgroundSubtractorSuBSENSE(_BackgroundSubtractor):
    """BackgroundSubtractorSuBSENSE_<lv::CUDA> (reference ctor: BackgroundSubtractorSuBSENSE.hpp:52-57)."""
    ALGO = ALGO_SUBSENSE

    def __init__(self, nDescDistThresholdOffset=3, nMinColorDistThreshold=30, nBGSamples=50, nRequiredBGSamples=2,
                 nSamplesForMovingAvgs=100, fRelLBSPThreshold=0.333, device=0, seed=0):
        p = default_params(ALGO_SUBSENSE)
        p.desc_dist_threshold, p.color_dist_threshold, p.n_samples = nDescDistThresholdOffset, nMinColorDistThreshold, nBGSamples
        p.n_required, p.n_samples_for_moving_avgs, p.rel_lbsp_threshold = nRequiredBGSamples, nSamplesForMovingAvgs, fRelLBSPThreshold
        super().__init__(p, device, seed)


class BackgroundSubtractorLOBSTER(_BackgroundSubtractor):
    """BackgroundSubtractorLOBSTER_<lv::CUDA> (reference ctor: BackgroundSubtractorLOBSTER.hpp:50-55)."""
    ALGO = ALGO_LOBSTER

    def __init__(self, nDescDistThreshold=4, nColorDistThreshold=30, nBGSamples=35, nRequiredBGSamples=2,
                 nLBSPThresholdOffset=0, fRelLBSPThreshold=0.333, device=0, seed=0):
        p = default_params(ALGO_LOBSTER)
        p.desc_dist_threshold, p.color_dist_threshold, p.n_samples = nDescDistThreshold, nColorDistThreshold, nBGSamples
        p.n_required, p.lbsp_threshold_offset, p.rel_lbsp_threshold = nRequiredBGSamples, nLBSPThresholdOffset, fRelLBSPThreshold
        super().__init__(p, device, seed)


class BackgroundSubtractorPAWCS(_BackgroundSubtractor):
    """BackgroundSubtractorPAWCS_<lv::CUDA> (reference ctor: BackgroundSubtractorPAWCS.hpp:51-55)."""
    ALGO = ALGO_PAWCS

    def __init__(self, nDescDistThresholdOffset=2, nMinColorDistThreshold=20, nMaxNbWords=50, nSamplesForMovingAvgs=100,
                 fRelLBSPThreshold=0.333, device=0, seed=0):
        p = default_params(ALGO_PAWCS)
        p.desc_dist_threshold, p.color_dist_threshold, p.n_samples = nDescDistThresholdOffset, nMinColorDistThreshold, nMaxNbWords
        p.n_samples_for_moving_avgs, p.rel_lbsp_threshold = nSamplesForMovingAvgs, fRelLBSPThreshold
        super().__init__(p, device, seed)

    def refreshModel(self, nBaseOccCount, fOccDecrFrac, bForceFGUpdate=False):
        """BackgroundSubtractorPAWCS::refreshModel(nBaseOccCount, fOccDecrFrac, bForceFGUpdate) (PAWCS.cpp:107-429)"""
        _chk(lib().lvb_pawcs_refresh_model(self._h, int(nBaseOccCount), float(fOccDecrFrac), int(bool(bForceFGUpdate))))


MASK_DILATE, MASK_ERODE, MASK_MEDIAN, MASK_HOLES = 0, 1, 2, 3


class _BackgroundSubtractorViBe:
    """BackgroundSubtractorViBe (video/include/litiv/video/BackgroundSubtractorViBe.hpp:50-77): initialize(img) / apply(img, lr=16) /
    getBackgroundImage(); a plain cv::BackgroundSubtractor in the reference (no ROI, no LBSP layer)."""
    MODEL_CHANNELS = None

    def __init__(self, nColorDistThreshold=20, nBGSamples=20, nRequiredBGSamples=2, device=0, seed=0):
        self._h = C.c_void_p()
        self.N = nBGSamples
        _chk(lib().lvb_vibe_create(self.MODEL_CHANNELS, nColorDistThreshold, nBGSamples, nRequiredBGSamples, device, seed, C.byref(self._h)))
        self.shape = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.lvb_vibe_destroy(h)
            self._h = None

    def _img(self, img, init=False):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or img.size == 0:
            raise LitivError("provided image must be non-empty, continuous, and of type 8UC1/8UC3")
        if not init:
            if self.shape is None:
                raise LitivError("algo must be initialized first")
            if img.shape[:2] != self.shape:
                raise LitivError("input image size mismatch with initialization size")
        return np.ascontiguousarray(img), (1 if img.ndim == 2 else img.shape[2])

    def initialize(self, img):
        img, c = self._img(img, init=True)
        h, w = img.shape[:2]
        _chk(lib().lvb_vibe_initialize(self._h, img.ctypes.data, w, h, c, w * c))
        self.shape = (h, w)

    def apply(self, img, learningRate=16.0, out=None):
        img, c = self._img(img)
        mask = np.empty(self.shape, np.uint8) if out is None else out
        _chk(lib().lvb_vibe_apply(self._h, img.ctypes.data, c, mask.ctypes.data, float(learningRate)))
        return mask

    def apply_device(self, d_img_ptr, channels, d_step, d_mask_ptr=None, learningRate=16.0):
        """device-resident frame (raw CUDA pointers); asynchronous on self.stream"""
        _chk(lib().lvb_vibe_apply_device(self._h, d_img_ptr, channels, d_step, d_mask_ptr, float(learningRate)))

    def sync(self):
        _chk(lib().lvb_vibe_sync(self._h))

    def getBackgroundImage(self):
        if self.shape is None:
            raise LitivError("algo must be initialized first")
        out = np.empty(self.shape + (self.MODEL_CHANNELS,), np.uint8)
        _chk(lib().lvb_vibe_get_background_image(self._h, out.ctypes.data))
        return out[..., 0] if self.MODEL_CHANNELS == 1 else out

    def getDefaultLearningRate(self):
        return 16.0

    @property
    def stream(self):
        return lib().lvb_vibe_stream(self._h)

    def model(self):
        """samples in the reference's layout [N][H][W][C] (m_voBGImg)"""
        out = np.empty((self.N,) + self.shape + (self.MODEL_CHANNELS,), np.uint8)
        _chk(lib().lvb_vibe_model(self._h, out.ctypes.data, out.nbytes, 0, 0))
        return out

    def set_model(self, arr, frame_idx):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        _chk(lib().lvb_vibe_model(self._h, arr.ctypes.data, arr.nbytes, 1, int(frame_idx)))

    def set_collect_stats(self, enabled):
        _chk(lib().lvb_vibe_set_collect_stats(self._h, int(enabled)))

    def stats(self):
        out = (C.c_uint64 * 5)()
        _chk(lib().lvb_vibe_get_stats(self._h, out))
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])

    def set_profile(self, enabled):
        _chk(lib().lvb_vibe_set_profile(self._h, int(enabled)))

    def get_profile(self):
        ms, n = C.c_double(), C.c_uint64()
        _chk(lib().lvb_vibe_get_profile(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class BackgroundSubtractorViBe_1ch(_BackgroundSubtractorViBe):
    """video/include/litiv/video/BackgroundSubtractorViBe.hpp:80-90"""
    MODEL_CHANNELS = 1


class BackgroundSubtractorViBe_3ch(_BackgroundSubtractorViBe):
    """video/include/litiv/video/BackgroundSubtractorViBe.hpp:93-103 (8UC3 frames, or 8UC1 frames expanded to BGR)"""
    MODEL_CHANNELS = 3


PBAS_STATE = {"bg_color": np.uint8, "bg_grad": np.uint8, "R": np.float32, "T": np.float32, "meanmin": np.float32, "rawmask": np.uint8,
              "lastgrad": np.uint8, "scalars": np.float64}


class _BackgroundSubtractorPBAS(_BackgroundSubtractorViBe):
    """BackgroundSubtractorPBAS (video/include/litiv/video/BackgroundSubtractorPBAS.hpp:88-121): initialize(img) /
    apply(img, learningRateOverride=-1) / getBackgroundImage()."""

    def __init__(self, nInitColorDistThreshold=30, fInitUpdateRate=16.0, nBGSamples=35, nRequiredBGSamples=2, device=0, seed=0):
        self._h = C.c_void_p()
        self.N = nBGSamples
        _chk(lib().lvb_pbas_create(self.MODEL_CHANNELS, nInitColorDistThreshold, fInitUpdateRate, nBGSamples, nRequiredBGSamples, device,
                                   seed, C.byref(self._h)))
        self.shape = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.lvb_pbas_destroy(h)
            self._h = None

    def initialize(self, img):
        img, c = self._img(img, init=True)
        h, w = img.shape[:2]
        _chk(lib().lvb_pbas_initialize(self._h, img.ctypes.data, w, h, c, w * c))
        self.shape = (h, w)

    def apply(self, img, learningRateOverride=-1.0, out=None):
        img, c = self._img(img)
        mask = np.empty(self.shape, np.uint8) if out is None else out
        _chk(lib().lvb_pbas_apply(self._h, img.ctypes.data, c, mask.ctypes.data, float(learningRateOverride)))
        return mask

    def apply_device(self, d_img_ptr, channels, d_step, d_mask_ptr=None, learningRateOverride=-1.0):
        _chk(lib().lvb_pbas_apply_device(self._h, d_img_ptr, channels, d_step, d_mask_ptr, float(learningRateOverride)))

    def sync(self):
        _chk(lib().lvb_pbas_sync(self._h))

    def getBackgroundImage(self):
        if self.shape is None:
            raise LitivError("algo must be initialized first")
        out = np.empty(self.shape + (self.MODEL_CHANNELS,), np.uint8)
        _chk(lib().lvb_pbas_get_background_image(self._h, out.ctypes.data))
        return out[..., 0] if self.MODEL_CHANNELS == 1 else out

    def getDefaultLearningRate(self):
        return -1.0

    @property
    def stream(self):
        return lib().lvb_pbas_stream(self._h)

    def _shape_of(self, name):
        if name in ("bg_color", "bg_grad"):
            return (self.N,) + self.shape + (self.MODEL_CHANNELS,)
        if name == "lastgrad":
            return self.shape + (self.MODEL_CHANNELS,)
        return (2,) if name == "scalars" else self.shape

    def state_get(self, name):
        out = np.empty(self._shape_of(name), PBAS_STATE[name])
        _chk(lib().lvb_pbas_state(self._h, name.encode(), out.ctypes.data, out.nbytes, 0))
        return out

    def state_set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=PBAS_STATE[name])
        _chk(lib().lvb_pbas_state(self._h, name.encode(), arr.ctypes.data, arr.nbytes, 1))

    model = set_model = None

    def set_collect_stats(self, enabled):
        _chk(lib().lvb_pbas_set_collect_stats(self._h, int(enabled)))

    def stats(self):
        out = (C.c_uint64 * 5)()
        _chk(lib().lvb_pbas_get_stats(self._h, out))
        return dict(roi_px=out[0], samples_scanned=out[1], sample_writes=out[2], fg_px=out[3], frames=out[4])

    def set_profile(self, enabled):
        _chk(lib().lvb_pbas_set_profile(self._h, int(enabled)))

    def get_profile(self):
        ms, n = C.c_double(), C.c_uint64()
        _chk(lib().lvb_pbas_get_profile(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class BackgroundSubtractorPBAS_1ch(_BackgroundSubtractorPBAS):
    """video/include/litiv/video/BackgroundSubtractorPBAS.hpp:124-134"""
    MODEL_CHANNELS = 1


class BackgroundSubtractorPBAS_3ch(_BackgroundSubtractorPBAS):
    """video/include/litiv/video/BackgroundSubtractorPBAS.hpp:137-147 (8UC3 frames, or 8UC1 frames expanded to BGR)"""
    MODEL_CHANNELS = 3


def mask_op(op, mask, param=0, device=0):
    """bit-packed GPU mask operator on a byte mask (see lvb_mask_op)"""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(mask)
    _chk(lib().lvb_mask_op(op, mask.ctypes.data, out.ctypes.data, mask.shape[1], mask.shape[0], param, device))
    return out


def pinned_empty(shape, dtype=np.uint8):
    """numpy array over page-locked host memory (lvb_host_alloc); frames/masks in it skip the staging copy"""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _chk(lib().lvb_host_alloc(C.byref(p), nbytes))
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


def apply_batch(subtractors, imgs, learningRate):
    """n independent streams, one frame each, overlapped on the GPU (lvb_apply_batch)."""
    n = len(subtractors)
    imgs = [s._check_img(i) for s, i in zip(subtractors, imgs)]
    masks = [np.empty(s.shape[:2], np.uint8) for s in subtractors]
    hs = (C.c_void_p * n)(*[s._h for s in subtractors])
    ip = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    mp = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
    _chk(lib().lvb_apply_batch(hs, ip, mp, n, float(learningRate)))
    return masks


class DeviceBatch:
    """n independent streams fed with device-resident frames (lvb_apply_batch_device): the handle / pointer arrays are built once"""

    def __init__(self, subtractors):
        self.subs = list(subtractors)
        self.n = len(self.subs)
        self._hs = (C.c_void_p * self.n)(*[s._h for s in self.subs])
        self._ip = (C.c_void_p * self.n)()
        self._mp = (C.c_void_p * self.n)()

    def apply(self, d_img_ptrs, d_step, d_mask_ptrs, learningRate):
        for i in range(self.n):
            self._ip[i] = d_img_ptrs[i]
            self._mp[i] = d_mask_ptrs[i] if d_mask_ptrs is not None else None
        _chk(lib().lvb_apply_batch_device(self._hs, self._ip, d_step, self._mp, self.n, float(learningRate)))


def lbsp_gradient(img, device=0):
    """dense LBSP::computeDescriptor_gradient (features2d LBSP.hpp:235-256): [H][W][4] u8 = gradX (int8), gradY (int8), magnitude, 0"""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim not in (2, 3) or img.size == 0:
        raise LitivError("input image must be non-empty, continuous, and of type 8UC1/8UC3")
    h, w = img.shape[:2]
    out = np.empty((h, w, 4), np.uint8)
    _chk(lib().lvb_lbsp_gradient(img.ctypes.data, w, h, 1 if img.ndim == 2 else img.shape[2], out.ctypes.data, device))
    return out


class EdgeDetectorLBSP:
    """EdgeDetectorLBSP (imgproc/include/litiv/imgproc/EdgeDetectorLBSP.hpp:33-83): apply_threshold(img, thr) -> 0 / 255 edge mask,
    apply(img) -> confidence map (16 per threshold that marks the pixel). Keeps its maps between calls like the reference object."""

    def __init__(self, nLevels=3, dHystLowThrshFactor=0.5, bNormalizeOutput=False, device=0):
        self._h = C.c_void_p()
        _chk(lib().lvb_edge_create(int(nLevels), float(dHystLowThrshFactor), device, C.byref(self._h)))
        if bNormalizeOutput:   # EdgeDetectorLBSP.cpp:431-432: cv::normalize(NORM_MINMAX) of the confidence map; the reference's default is false
            _chk(lib().lvb_edge_set_normalize(self._h, 1))
        self._shape = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.lvb_edge_destroy(self._h)
            self._h = None

    def getDefaultThreshold(self):
        return lib().lvb_edge_default_threshold()

    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or img.size == 0 or (img.ndim == 3 and img.shape[2] not in (1, 2, 3, 4)):
            raise LitivError("input image must be non-empty and continuous, 8UC1 .. 8UC4")
        return img, (1 if img.ndim == 2 else img.shape[2])

    def apply_threshold(self, img, dDetThreshold=0.5):
        img, c = self._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().lvb_edge_apply_threshold(self._h, img.ctypes.data, w, h, c, out.ctypes.data, float(dDetThreshold)))
        self._shape = (h, w)
        return out

    def apply(self, img):
        img, c = self._img(img)
        h, w = img.shape[:2]
        out = np.empty((h, w), np.uint8)
        _chk(lib().lvb_edge_apply(self._h, img.ctypes.data, w, h, c, out.ctypes.data))
        self._shape = (h, w)
        return out

    def gradient_map(self):
        if self._shape is None:
            raise LitivError("no pass has run yet")
        out = np.empty(self._shape + (4,), np.uint8)
        _chk(lib().lvb_edge_get_gradient_map(self._h, out.ctypes.data))
        return out

    def flood_sweeps(self):
        return int(lib().lvb_edge_flood_sweeps(self._h))

    def apply_threshold_device(self, d_img_ptr, w, h, channels, d_step, d_edges_ptr=None, dDetThreshold=0.5):
        """frame already in device memory (raw pointers, row pitch d_step bytes); returns when the mask is complete"""
        _chk(lib().lvb_edge_apply_threshold_device(self._h, d_img_ptr, w, h, channels, d_step, d_edges_ptr, float(dDetThreshold)))
        self._shape = (h, w)

    @property
    def stream(self):
        return lib().lvb_edge_stream(self._h)


class LBSP:
    """Dense LBSP extractor (features2d LBSP::compute2). LBSP(20) -> absolute threshold; LBSP(0.333, 0) -> relative."""

    def __init__(self, threshold, nThresholdOffset=None, device=0):
        if isinstance(threshold, float) or nThresholdOffset is not None:
            if threshold < 0:
                raise LitivError("relative LBSP threshold must be non-negative")
            self.rel, self.thr = float(threshold), int(nThresholdOffset or 0)
        else:
            self.rel, self.thr = None, int(threshold)
        self.ref = None
        self.device = device

    def setReference(self, img):
        self.ref = None if img is None else np.ascontiguousarray(img, dtype=np.uint8)

    def windowSize(self):
        return (5, 5)

    def borderSize(self, nDim=0):
        """LBSP::borderSize (features2d/src/LBSP.cpp: only dimensions 0 and 1 exist; features2d/test/lbsp.cpp:13-14)"""
        if nDim not in (0, 1):
            raise LitivError("border size is only defined for 2 dimensions")
        return 2

    def descriptorSize(self):
        return 2

    def descriptorType(self):
        """CV_16U == CV_16UC1 (features2d/test/lbsp.cpp:16-17)"""
        return 2

    def defaultNorm(self):
        """cv::NORM_HAMMING (features2d/test/lbsp.cpp:18)"""
        return 6

    def compute2(self, img):
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8 or img.size == 0:
            raise LitivError("input image must be non-empty, continuous, and of type 8UC1/8UC3")
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        out = np.zeros((h, w, c), np.uint16)
        rp = None
        if self.ref is not None:
            if self.ref.shape != img.shape:
                raise LitivError("ref image must be empty, or of the same size/type as the input image")
            rp = self.ref.ctypes.data
        _chk(lib().lvb_lbsp_compute(img.ctypes.data, rp, w, h, c, int(self.rel is not None), self.rel or 0.0, self.thr, out.ctypes.data, self.device))
        return out[..., 0] if c == 1 else out


class BinClassif:
    """lv::BinClassif (modules/datasets/include/litiv/datasets/metrics.hpp:32-67) with accumulate() on the device."""
    NAMES = ("nTP", "nTN", "nFP", "nFN", "nSE", "nDC")   # BinClassif::CountersList order

    def __init__(self, device=0):
        self.counters = np.zeros(6, np.uint64)
        self.device = device

    def __getattr__(self, name):
        if name in BinClassif.NAMES:
            return int(self.counters[BinClassif.NAMES.index(name)])
        raise AttributeError(name)

    def total(self, bWithDontCare=False):
        return int(self.counters[:4].sum()) + (int(self.counters[5]) if bWithDontCare else 0)

    @staticmethod
    def _mat(a, shape=None):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.uint8)
        if a.ndim != 2 or (shape is not None and a.shape != shape):
            raise LitivError("all input mat sizes must match")
        return a

    def accumulate(self, oClassif, oGT=None, oROI=None):
        """oClassif: a host 8UC1 mask, or a background subtractor instance (its latest foreground mask is scored in HBM)"""
        if isinstance(oClassif, _BackgroundSubtractor):
            h, w, _ = oClassif.shape
            gt, roi = self._mat(oGT, (h, w)), self._mat(oROI, (h, w))
            _chk(lib().lvb_binclassif_accumulate(oClassif._h, gt.ctypes.data if gt is not None else None,
                                                 roi.ctypes.data if roi is not None else None, self.counters.ctypes.data))
            return self
        m = self._mat(oClassif)
        if m is None or m.size == 0:
            raise LitivError("binary classifier results must be non-empty and of type 8UC1")
        gt, roi = self._mat(oGT, m.shape), self._mat(oROI, m.shape)
        _chk(lib().lvb_binclassif(m.ctypes.data, gt.ctypes.data if gt is not None else None, roi.ctypes.data if roi is not None else None,
                                  m.shape[1], m.shape[0], self.counters.ctypes.data, self.device))
        return self

    def metrics(self):
        """BinClassifMetrics (metrics.hpp:213-257)"""
        out = np.zeros(8, np.float64)
        _chk(lib().lvb_binclassif_metrics(self.counters.ctypes.data, out.ctypes.data))
        return dict(zip(("dRecall", "dSpecificity", "dFPR", "dFNR", "dPBC", "dPrecision", "dFMeasure", "dMCC"), out.tolist()))
