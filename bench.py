#!/usr/bin/env python
"""bench.py — SuBSENSE Mpx/s at 1080p per B200 (BASELINE.json metric), weak-scaled one stream per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one apply() of the hot path on one 1920x1080 RGB frame of a synthetic CDnet-shaped sequence
(BASELINE.json configs[3]).  `value` is measured with the frames already resident in HBM (lvb_apply_device), `e2e`
through the reference-facing call (apply(img, fgmask, lr) with HOST buffers in pinned memory: H2D copy of the frame
and D2H copy of the mask inside the timed region).  The timed region starts after the reference's own warm-up
protocol (samples/changedet/src/main.cpp:56: learning-rate override 1 for the first frames, then T(x)).
`--impl reference` times the CPU restatement of the reference (oracle/, reference-order mode: the reference itself
needs OpenCV C++ and cannot be built in this image) on the host cores, one stream per thread.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA initialises (see litiv_b200/api.py)

W, H, C = 1920, 1080, 3
BOOT_FRAMES = 60          # protocol frames before anything is timed (lr=1 for the first 50)
N_UNIQUE = 24             # distinct synthetic frames kept resident and played ping-pong (continuous motion)
METRIC = "subsense_1080p_mpx_per_s"
WORKLOAD = "SuBSENSE 1920x1080 RGB single stream per GPU (BASELINE.json configs[3])"
SCAN_DRAM_BYTES_NCU = 436.6e6  # per subsense_scan launch at this workload (profiles/r01h_scan_feedback_ncu.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region: NVML (a few hundred samples per second) when nvidia_ml_py is
    importable, else the nvidia-smi query line of B200_PROFILING.md (a handful of samples)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")) and index < len(vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
            while not self._stop_evt.is_set():
                try:
                    sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    try:
                        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if r & b else "Not Active" for b, _ in bits])
                except Exception:
                    pass
                self._stop_evt.wait(0.004)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_frames(seed, n):
    from litiv_b200.synth import SynthSequence
    seq = SynthSequence(W, H, C, seed=seed)
    return seq, [seq.frame(t) for t in range(n)]


def pingpong(i, n):
    """0,1,..,n-1,n-2,..,1,0,1.. : continuous motion over a finite set of frames"""
    period = 2 * (n - 1)
    k = i % period
    return k if k < n else period - k


def lr_for(t):
    return 1.0 if t <= 50 else 0.0


def cpu_baseline(frames, seconds_budget=20.0):
    """bounded sample of the same workload on ONE host core (the reference is single-threaded per stream)"""
    from oracle import oracle as O
    o = O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_REFERENCE, seed=0)
    o.initialize(frames[0])
    n_done, t_in = 0, 0.0
    t0 = time.time()
    while time.time() - t0 < seconds_budget and n_done < 16:
        k = n_done + 1
        t, _ = o.apply_sequence(frames[pingpong(k, len(frames))][None], [lr_for(k)])
        t_in += t
        n_done += 1
    return {"value": W * H * n_done / t_in / 1e6, "unit": "Mpx/s", "cores": 1, "kind": "port",
            "sample": f"first {n_done} frames (lr=1 bootstrap phase) of the same 1080p sequence, oracle reference-order mode, 1 thread"}


def run_reference(args):
    """--impl reference: the CPU restatement on the host cores, one 1080p stream per thread (lv::WorkerPool model)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.lib()
    ncores = os.cpu_count() or 1
    nthreads = max(1, min(ncores, 16))
    seq, frames = make_frames(4, 8)
    oracles = []
    for i in range(nthreads):
        o = O.Oracle(O.ALGO_SUBSENSE, mode=O.MODE_REFERENCE, seed=i)
        o.initialize(frames[0])
        oracles.append(o)
    counter = [0]

    def step():
        counter[0] += 1
        k = counter[0]
        f = frames[pingpong(k, len(frames))][None]
        ths = [threading.Thread(target=o.apply_sequence, args=(f, [lr_for(k)])) for o in oracles]
        [t.start() for t in ths]
        [t.join() for t in ths]

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    val = W * H * nthreads * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpx/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H, C], "cpu_streams": nthreads,
                       "note": "same 1080p workload, one independent stream per host thread (the reference is single-threaded per stream)"},
            "cpu_baseline": {"value": val, "unit": "Mpx/s", "cores": nthreads, "kind": "port",
                             "sample": f"{args.steps} frames x {nthreads} independent 1080p streams, oracle reference-order mode (reference needs OpenCV C++: unbuildable here)"},
            "e2e": {"value": val, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 12)
        args.warmup = min(args.warmup, 2)
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import litiv_b200 as lv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or lv.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: litiv_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seq, frames = make_frames(4 + 1000 * rank, N_UNIQUE)
    dev = torch.device("cuda", local)
    pitch = (W * C + 127) // 128 * 128
    d_frames = torch.zeros((N_UNIQUE, H, pitch), dtype=torch.uint8, device=dev)
    for i, f in enumerate(frames):
        d_frames[i, :, :W * C] = torch.from_numpy(f.reshape(H, W * C)).to(dev)
    d_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)

    sub = lv.BackgroundSubtractorSuBSENSE(device=local, seed=rank)
    sub.initialize(frames[0])
    stream = torch.cuda.ExternalStream(sub.stream, device=dev)
    frame_no = [0]

    def step_device():
        frame_no[0] += 1
        k = frame_no[0]
        i = pingpong(k, N_UNIQUE)
        sub.apply_device(d_frames[i].data_ptr(), pitch, d_mask.data_ptr(), lr_for(k))

    for _ in range(BOOT_FRAMES):
        step_device()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lv.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    sub.flush()          # the mask chain of the last frames runs on a side stream: order e1 behind it
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lv.kernel_launch_count() - l0
    clocks = sampler.stop()

    # dominant kernels: per-launch CUDA-event timing on the instance stream + the algorithmic bytes they moved. The reference's
    # per-pixel loop is two kernels here (scan: LBSP + sample consensus; feedback: maps + stochastic updates), so SURVEY.md §8(d)'s
    # per-pixel figure B_alg = 131 + 9(s+u) is split between them: B_scan = 27 + 9 s + 9 u_nb, B_fb = 104 + 9 u_own.
    sub.set_profile(True)
    sub.set_collect_stats(True)
    nprof = max(20, min(args.steps, 100))
    t_prof0 = torch.cuda.Event(enable_timing=True); t_prof1 = torch.cuda.Event(enable_timing=True)
    t_prof0.record(stream)
    for _ in range(nprof):
        step_device()
    sub.flush(); t_prof1.record(stream)
    torch.cuda.synchronize()
    pa_ms, pa_n = sub.get_profile()
    fb_ms, fb_n = sub.get_profile_feedback()
    tp_ms, tp_n = sub.get_profile_tail()
    st = sub.stats()
    sub.set_profile(False)
    sub.set_collect_stats(False)
    roi_px = st["roi_px"] / max(st["frames"], 1)
    sbar = st["samples_scanned"] / max(st["roi_px"], 1)
    u = st["sample_writes"] / max(st["roi_px"], 1)
    u_nb = u / 2.0                               # own-slot and neighbour writes are drawn with the same rate (1/LR each)
    b_alg = 131.0 + 9.0 * (sbar + u)            # B_alg = B_fixed(110+7C) + 3C*(s + u), C=3
    b_scan = 27.0 + 9.0 * sbar + 9.0 * u_nb     # input 3 + raw 1 + lastColor RW 6 + lastDesc RW 12 + R 4 + unstable 1 ; samples read ; nb writes
    b_fb = b_alg - b_scan
    hbm_peak, peak_src = peaks()
    pa_avg_ms = pa_ms / max(pa_n, 1)
    fb_avg_ms = fb_ms / max(fb_n, 1)
    achieved = roi_px * b_scan / (pa_avg_ms * 1e-3) / 1e9 if pa_avg_ms > 0 else 0.0
    fb_achieved = roi_px * b_fb / (fb_avg_ms * 1e-3) / 1e9 if fb_avg_ms > 0 else 0.0

    # end to end through the reference-facing C-ABI call with HOST buffers (pinned): every step uploads its frame and reads its
    # mask back inside the timed region. Headline: the asynchronous form (lvb_apply_async / lvb_sync_next: two frames in flight,
    # the upload of frame k+1 overlaps the kernels of frame k - the `apply_cuda` async mode the reference's apps/changedet expects);
    # the strictly synchronous apply(img, fgmask, lr) is timed too and reported next to it.
    h_frames = [lv.pinned_empty((H, W, C)) for _ in range(4)]
    for hf, f in zip(h_frames, frames[:4]):
        hf[...] = f
    h_masks = [lv.pinned_empty((H, W)) for _ in range(2)]
    e2e_steps = max(10, min(args.steps, 100))
    for j in range(3):
        sub.apply(h_frames[j % 4], 0.0, out=h_masks[0])
    barrier()
    t0 = time.perf_counter()
    for j in range(e2e_steps):
        sub.apply(h_frames[pingpong(j, 4)], 0.0, out=h_masks[0])
    torch.cuda.synchronize()
    sync_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    sub.apply_async(h_frames[0], 0.0, out=h_masks[0])
    for j in range(1, e2e_steps):
        sub.apply_async(h_frames[pingpong(j, 4)], 0.0, out=h_masks[j % 2])
        sub.sync_next()
    sub.sync_next()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    t_max = torch.tensor([ms, e2e_s * 1e3, sync_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all, sync_ms_all = float(t_max[0]), float(t_max[1]), float(t_max[2])

    if rank == 0:
        value = W * H * args.steps * world / (ms_all * 1e-3) / 1e6
        e2e_val = W * H * e2e_steps * world / (e2e_ms_all * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H, C],
                       "streams_per_gpu": 1, "fps_per_stream": args.steps / (ms_all * 1e-3), "boot_frames": BOOT_FRAMES,
                       "l2": "per-frame working set (sample model 1.66 GB + maps) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "Mpx/s", "h2d_bytes_per_step": W * H * C, "d2h_bytes_per_step": W * H, "steps": e2e_steps,
                    "api": "lvb_apply_async(host frame, host mask, lr) + lvb_sync_next: two frames in flight, pinned host buffers",
                    "synchronous_apply_value": W * H * e2e_steps * world / (sync_ms_all * 1e-3) / 1e6},
            "roofline": {"bound": "hbm", "kernel": "subsense_scan<3>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": SCAN_DRAM_BYTES_NCU,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/r01h_scan_feedback_ncu.md)", "peak_source": peak_src, "avg_launch_ms": pa_avg_ms,
                         "launches_timed": int(pa_n), "alg_bytes_per_px": b_scan, "scan_depth": sbar, "sample_writes_per_px": u,
                         "roi_px": roi_px, "kernel_share_of_step": pa_avg_ms / (ms_all / args.steps),
                         "tail_passes_avg_ms": tp_ms / max(tp_n, 1),
                         "second_kernel": {"kernel": "subsense_feedback<3>", "avg_launch_ms": fb_avg_ms, "alg_bytes_per_px": b_fb,
                                           "achieved": fb_achieved, "frac": fb_achieved / hbm_peak, "kernel_share_of_step": fb_avg_ms / (ms_all / args.steps)},
                         "frame": {"alg_bytes_per_px": b_alg, "achieved": roi_px * b_alg / (ms_all / args.steps * 1e-3) / 1e9,
                                   "frac": roi_px * b_alg / (ms_all / args.steps * 1e-3) / 1e9 / hbm_peak,
                                   "note": "whole frame: SURVEY 8(d) B_alg x ROI px / ms_per_step (scan + feedback are on the critical path, the mask chain overlaps them)"}},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(frames)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
