#!/usr/bin/env python
"""The reference's CDnet sandbox (apps/changedet/src/main.cpp:346-428) on this package: for every sequence directory found
under --root (`<category>/<sequence>/{input,groundtruth,ROI.bmp,ROI.jpg}`), initialize / apply / score, print the speed and the
CDnet metrics. `--synthetic WxH:N` first writes a CDnet-shaped synthetic sequence (there is no network to fetch the dataset).
  python tools/changedet.py --synthetic 320x240:300 --algo subsense
  python tools/changedet.py --root /data/CDnet2014/dataset --algo pawcs --save results/"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import litiv_b200 as lv
from litiv_b200 import datasets as D

ap = argparse.ArgumentParser()
ap.add_argument("--root")
ap.add_argument("--synthetic", help="WxH:frames")
ap.add_argument("--algo", default="subsense", choices=["subsense", "lobster", "pawcs", "vibe", "pbas"])
ap.add_argument("--no-eval", action="store_true", help="throughput mode: no scoring, two frames in flight")
ap.add_argument("--no-precache", action="store_true")
ap.add_argument("--save", help="directory for bin%%06d.png masks")
args = ap.parse_args()
root = args.root
if args.synthetic:
    size, n = args.synthetic.split(":")
    w, h = (int(v) for v in size.split("x"))
    root = tempfile.mkdtemp(prefix="cdnet_synth_")
    D.write_synthetic_cdnet(root, "synthetic", w, h, int(n), seed=7)
cls = {"subsense": lv.BackgroundSubtractorSuBSENSE, "lobster": lv.BackgroundSubtractorLOBSTER, "pawcs": lv.BackgroundSubtractorPAWCS,
       "vibe": lv.BackgroundSubtractorViBe_3ch, "pbas": lv.BackgroundSubtractorPBAS_3ch}[args.algo]
for cat in sorted(os.listdir(root)):
    cdir = os.path.join(root, cat)
    if not os.path.isdir(cdir):
        continue
    for name in sorted(os.listdir(cdir)):
        sdir = os.path.join(cdir, name)
        if not os.path.isdir(os.path.join(sdir, "input")):
            continue
        seq = D.CDnetSequence(sdir)
        out = D.analyze(seq, cls(seed=0), evaluate=not args.no_eval, output_dir=os.path.join(args.save, cat, name) if args.save else None,
                        precache=not args.no_precache)
        line = {"category": cat, "sequence": name, "algo": args.algo, "frame": [seq.frame_size[1], seq.frame_size[0], seq.channels],
                "frames": out["frames"], "seconds": round(out["seconds"], 4), "hz": round(out["hz"], 1),
                "mpx_per_s": round(out["hz"] * seq.frame_size[0] * seq.frame_size[1] / 1e6, 1)}
        if out["metrics"]:
            line.update({k: round(v, 5) for k, v in out["metrics"].items()})
        print(json.dumps(line))
