"""litiv_b200 — B200-native (sm_100a) change-detection hot path of plstcharles/litiv:
LBSP descriptors + SuBSENSE / LOBSTER / PAWCS apply(), behind the reference's IBackgroundSubtractor surface."""
from .api import (ALGO_LOBSTER, ALGO_PAWCS, ALGO_SUBSENSE, LBSP, BinClassif, DeviceBatch, BackgroundSubtractorLOBSTER,  # noqa: F401
                  BackgroundSubtractorPAWCS, BackgroundSubtractorSuBSENSE, BackgroundSubtractorViBe_1ch, BackgroundSubtractorViBe_3ch,
                  BackgroundSubtractorPBAS_1ch, BackgroundSubtractorPBAS_3ch, LitivError, Params, apply_batch, default_params, device_count,
                  kernel_launch_count, lbsp_gradient, EdgeDetectorLBSP, lib, lib_path, mask_op, pinned_empty,
                  MASK_DILATE, MASK_ERODE, MASK_MEDIAN, MASK_HOLES)
