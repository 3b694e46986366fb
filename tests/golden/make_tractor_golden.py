"""Generates tests/golden/tractor_crop.npz from the reference's real 1080p sample sequence (samples/data/tractor.mp4, decoded with cv2 in this
container; the file itself is not copied): 20 consecutive frames (60..79, the tractor is moving through the crop) of a 320x240 crop, plus what the
REFERENCE'S OWN CODE (oracle/_ref) produces on them for SuBSENSE, LOBSTER and PAWCS with srand(0): SHA-256 of every final mask, the last mask, the
background image. The fixture travels to the GPU box, where /root/reference does not exist.
Run from the repo root: python tests/golden/make_tractor_golden.py [/root/reference]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
X0, Y0, W, H, F0, NF = 800, 420, 320, 240, 60, 20


def main():
    import cv2
    from oracle import ref as R
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    cap = cv2.VideoCapture(os.path.join(ref_root, "samples/data/tractor.mp4"))
    assert cap.isOpened()
    frames = []
    for i in range(F0 + NF):
        ok, f = cap.read()
        assert ok and f.shape == (1080, 1920, 3)
        if i >= F0:
            frames.append(np.ascontiguousarray(f[Y0:Y0 + H, X0:X0 + W]))
    frames = np.stack(frames)
    out = {"frames": frames, "crop": np.array([X0, Y0, W, H, F0, NF])}
    for name, algo, lr in (("lobster", 0, 16.0), ("subsense", 1, None), ("pawcs", 2, 0.0)):
        for gray in (False, True):
            fr = frames if not gray else np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
            a = R.Reference(algo, seed=0)
            a.initialize(fr[0])
            h = hashlib.sha256()
            last = None
            for t in range(1, NF):
                rate = lr if lr is not None else (1.0 if t <= 8 else 0.0)   # samples/changedet protocol, shortened
                last = a.apply(fr[t], rate)
                h.update(last.tobytes())
            key = f"{name}_{'gray' if gray else 'rgb'}"
            out[key + "_masks_sha256"] = np.frombuffer(h.digest(), np.uint8)
            out[key + "_last_mask"] = last
            out[key + "_bg"] = a.get_background_image()
            print(key, h.hexdigest()[:16], "fg px in last mask:", int((last > 0).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tractor_crop.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(ROOT, "tests", "golden", "tractor_crop.npz")), "bytes")


if __name__ == "__main__":
    main()
