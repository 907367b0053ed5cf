#!/usr/bin/env python
"""Long-horizon / unknown-map runs of BASELINE configs 3 and 5 on one B200, host out of the frame loop.

    python tools/long_run.py c5 [--frames 10000] [--particles 1048576] [--capacity 1024] [--report 1000]
    python tools/long_run.py c3 [--frames 200]   [--particles 4194304] [--capacity 256]  [--report 50]

Spawn mode (``FastSLAM(spawn=True)``: the reference's new-landmark path with the three documented patches, SURVEY.md
A.6) on an initially EMPTY map; the corridor world holds ``capacity`` true landmarks, the robot drives along it
(v = 0.2 m/s, 11 Hz) and the map of every particle grows as landmarks come into view.  Scans come from the
device-side ``BearingSimulator`` (the K = 8 landmarks nearest the true pose), so nothing but kernel launches and
one scalar control per frame leaves the host.  Every ``--report`` frames one JSON line is printed: throughput over
the window (CUDA events), the fused kernel's and the resampler's average time (sampled frames), mean / max live
landmarks per particle, matched fraction, landmarks spawned / readings stored, flags, pose error of the estimate.
The last line is a summary.  fp32 landmark records, fp32 landmark algebra (the bench's instantiation).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c5"])
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--particles", type=int, default=None)
    ap.add_argument("--capacity", type=int, default=None)
    ap.add_argument("--report", type=int, default=None)
    ap.add_argument("--orphans", type=int, default=32)
    ap.add_argument("--true-landmarks", type=int, default=None,
                    help="landmarks in the world (default: 45 %% of the capacity for c5 -- pairing rays creates about as many "
                         "spurious potential landmarks as true ones, and a full map cannot take in what comes into view -- the capacity for c3)")
    ap.add_argument("--arith", default="f32", choices=["f32", "f64"])
    ap.add_argument("--sample-every", type=int, default=20, help="frames between kernel-time samples")
    args = ap.parse_args(argv)
    preset = {"c3": dict(frames=200, particles=1 << 22, capacity=256, report=50),
              "c5": dict(frames=10000, particles=1 << 20, capacity=1024, report=1000)}[args.config]
    T = args.frames or preset["frames"]
    M = args.particles or preset["particles"]
    N = args.capacity or preset["capacity"]
    report = args.report or preset["report"]
    K = 8

    import numpy as np
    import torch
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT, DT_NSEC, make_world
    from parakeet_slam_b200.simulator import BearingSimulator

    torch.cuda.set_device(0)
    free, total = torch.cuda.mem_get_info()
    n_true = args.true_landmarks or (int(0.45 * N) if args.config == "c5" else N)
    world = make_world(n_true, "corridor", T, 0.2, 0.1, seed=2024)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    urng = random.Random(12345)
    fs = FastSLAM([], num_particles=M, capacity=N, dtype="f32", arithmetic=args.arith, noise="philox", seed=2024,
                  uniform=urng.random, clock=clk, spawn=True, orphan_capacity=args.orphans)
    sim = BearingSimulator(world, obs_per_frame=K, seed=7)
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = 0.2, 0.0
    fs.last_control = tw
    ev = lambda: torch.cuda.Event(enable_timing=True)
    x = y = th = 0.0
    header = dict(config=args.config, particles=M, capacity=N, blobs=K, frames=T, orphan_slots=args.orphans,
                  landmark_storage="f32", arithmetic=args.arith, block_bytes=fs.block_bytes,
                  pool_gb=M * fs.block_bytes / 1e9, hbm_free_gb_before=free / 1e9,
                  true_landmarks=n_true,
                  scans="BearingSimulator on device (K nearest landmarks of %d, sigma_bearing 0.02, sigma_colour 0.3)" % n_true)
    print(json.dumps(header), flush=True)
    lines = []
    win_start = ev()
    win_start.record()
    samples = []
    tot = dict(matched=0, unmatched=0, spawned=0, orphaned=0, promoted=0, flags=0, copied=0)
    t_wall0 = time.perf_counter()
    for t in range(T):
        clk.ns += DT_NSEC
        x += 0.2 * DT * math.cos(th)
        y += 0.2 * DT * math.sin(th)
        scan = sim.scan((x, y, th))
        sampled = (t % args.sample_every) == args.sample_every - 1
        if sampled:
            e = [ev() for _ in range(4)]
            e[0].record()
        fs.motion_update(tw)
        if sampled:
            fs.wait_blocks()   # the previous frame's block copies (copy stream) stay out of K2's interval
            e[1].record()
        fs.measurement_update(scan)                    # K2 + K2b, scan resident on the device
        if sampled:
            e[2].record()
        fs.low_variance_resample()
        if sampled:
            fs.wait_blocks()   # sampled frames time the whole resampling chain, copies included
            e[3].record()
            samples.append(e)
            st = fs.stats()                            # (synchronises; only on sampled frames)
            for k in ("matched", "unmatched", "spawned", "orphaned", "promoted"):
                tot[k] += st[k]
            tot["flags"] |= st["flags"]
            tot["copied"] += st["blocks_copied"]
        if (t + 1) % report == 0 or t + 1 == T:
            win_stop = ev()
            win_stop.record()
            torch.cuda.synchronize()
            frames_in_win = (t % report) + 1
            ms = win_start.elapsed_time(win_stop)
            k2 = [s[1].elapsed_time(s[2]) for s in samples]
            rs = [s[2].elapsed_time(s[3]) for s in samples]
            nl = fs.aux[:, 0].to(torch.float64)
            est = fs.summary()
            n_s = max(1, len(samples))
            line = dict(frame=t + 1, frames_in_window=frames_in_win, window_ms=ms, ms_per_frame=ms / frames_in_win,
                        updates_per_s=M * K * frames_in_win / (ms * 1e-3),
                        measure_ms_avg=sum(k2) / n_s, resample_ms_avg=sum(rs) / n_s, sampled_frames=len(samples),
                        n_live_mean=float(nl.mean().item()), n_live_max=int(nl.max().item()),
                        matched_fraction=tot["matched"] / float(max(1, tot["matched"] + tot["unmatched"])),
                        spawned_per_particle_frame=tot["spawned"] / float(M * n_s),
                        orphaned_per_particle_frame=tot["orphaned"] / float(M * n_s),
                        promoted_per_particle_frame=tot["promoted"] / float(M * n_s),
                        f_dup_mean=tot["copied"] / float(M * n_s), flags=tot["flags"],
                        pose_error_m=math.hypot(est[0] - x, est[1] - y), true_x=x)
            print(json.dumps(line), flush=True)
            lines.append(line)
            samples = []
            tot = dict(matched=0, unmatched=0, spawned=0, orphaned=0, promoted=0, flags=0, copied=0)
            win_start = ev()
            win_start.record()
    wall = time.perf_counter() - t_wall0
    total_ms = sum(l["window_ms"] for l in lines)
    print(json.dumps(dict(summary=True, config=args.config, frames=T, particles=M, capacity=N,
                          updates_per_s=M * K * T / (total_ms * 1e-3), device_s=total_ms * 1e-3, wall_s=wall,
                          n_live_mean_final=lines[-1]["n_live_mean"], n_live_max_final=lines[-1]["n_live_max"],
                          flags_all=int(np.bitwise_or.reduce([l["flags"] for l in lines])))), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
