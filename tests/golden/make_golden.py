"""Generates the committed golden fixtures from the reference's own test data (run in the build container only;
/root/reference does not exist on the GPU box).

lbsp_golden.npz  <- modules/features2d/test/lbsp.cpp:21-79 (regression_compute): 65x65 crop around (371,371) of
                    samples/data/multispectral_stereo_ex/img2.png, LBSP(size_t(20)) dense descriptors stored in
                    modules/features2d/test/data/test_lbsp.bin (lv::write MatArchive_BINARY: int32 type, u64 elemSize,
                    u64 total, int32 dims, int32 size[dims], raw data).
"""
import struct
import sys
import numpy as np
import cv2

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
img = cv2.imread(f"{REF}/samples/data/multispectral_stereo_ex/img2.png")
assert img is not None and img.shape == (600, 800, 3)
crop = np.ascontiguousarray(img[371 - 32:371 + 33, 371 - 32:371 + 33])
raw = open(f"{REF}/modules/features2d/test/data/test_lbsp.bin", "rb").read()
mtype, esz, total, dims = struct.unpack("<iQQi", raw[:24])
sizes = struct.unpack("<%di" % dims, raw[24:24 + 4 * dims])
assert (esz, total, dims, sizes) == (6, 65 * 65, 2, (65, 65))
gold = np.frombuffer(raw[24 + 4 * dims:], np.uint16).reshape(65, 65, 3)
np.savez_compressed(__file__.replace("make_golden.py", "lbsp_golden.npz"), crop=crop, desc=gold, abs_threshold=20)
print("wrote lbsp_golden.npz", crop.shape, gold.shape)
