#!/usr/bin/env python
"""Benchmark of the FastSLAM hot path (BASELINE.json metric: particle x observation updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one filter frame over one batch of synthetic ``synth360`` input: motion update,
fused association + EKF + weight kernel, low-variance resample (scan, ancestors, copy-on-resample).
N = 1 runs BASELINE config 2 (2^20 particles x 64 landmarks, 8 bearings/frame, fp32 landmark
storage, on-device Philox motion noise).  N > 1 (under torchrun) shards particles over ranks with
the per-GPU work fixed ("weak" scaling): N * 2^20 particles in one filter.

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput, ``e2e`` the same metric
through the drop-in Python API (``FastSLAM.cam_cb`` with host ``VizScan`` messages plus
``summary()`` read back every frame).  ``roofline`` describes the dominant kernel (the fused
measurement update) timed with CUDA events inside the timed steps.
"""
from __future__ import annotations

import argparse
import json
import math
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle_observation_updates_per_sec"
UNIT = "updates/s"
PARTICLES_PER_GPU = 1 << 20
LANDMARKS = 64
BLOBS = 8
KERNELS_PER_STEP = 11  # motion, measure, weight_scan, thresholds, ancestors, fill_runs,
#                        dead_scan, block_offsets, free_list, assign, copy_blocks
KERNELS_PER_STEP_PEER = 18   # + 2 peer barriers, exchange plan, push headers, push blocks, offspring window,
#                              unpack blocks (sharded filter, peer exchange; no NCCL kernel in a frame)
KERNELS_PER_STEP_NCCL = 16   # + 2 x (pack headers, pack blocks), offspring window, unpack (plus 2 NCCL collectives)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the NumPy restatement of the reference (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, particles, landmarks, blobs, frames = args
    import numpy as np
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c2", num_particles=particles, num_landmarks=landmarks, obs_per_frame=blobs,
                        frames=frames, motion_seed=seed)
    st = onp.OracleState(particles, scn.landmarks, preset_covar=scn.preset_covar)
    rs = np.random.RandomState(seed)
    t0 = time.perf_counter()
    for t in range(frames):
        onp.frame(st, scn.observations[t], rs.standard_normal((particles, 3)), scn.v, scn.w, scn.dt,
                  float(scn.u01[t]), sequential_resample=False)
    return time.perf_counter() - t0


def cpu_baseline(particles_per_proc=2048, frames=20, procs=None, landmarks=LANDMARKS, blobs=BLOBS):
    """Throughput of the oracle port with one replica per host core (the reference itself is
    single-threaded Python; independent replicas are how it would use a whole host)."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    jobs = [(1000 + i, particles_per_proc, landmarks, blobs, frames) for i in range(procs)]
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(1, 64, landmarks, blobs, 1)] * procs)  # import + warm-up
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    updates = procs * particles_per_proc * blobs * frames
    return {"value": updates / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "%d replicas x %d particles x %d landmarks x %d blobs x %d frames of the "
                      "config-2 scenario, NumPy oracle port (oracle/fastslam_np.py), %.1f s wall"
                      % (procs, particles_per_proc, landmarks, blobs, frames, wall)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port; the Python-2/ROS reference cannot
    travel to the GPU box) on all host cores, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    procs = os.cpu_count() or 1
    per_proc = 1024
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(1, 64, LANDMARKS, BLOBS, 1)] * procs)
        for _ in range(min(warm, 1)):
            pool.map(_cpu_worker, [(7 + i, per_proc, LANDMARKS, BLOBS, 1) for i in range(procs)])
        # bound the whole run to a few minutes: at most 12 timed steps, one frame each
        timed = min(steps, 12)
        t0 = time.perf_counter()
        for s in range(timed):
            pool.map(_cpu_worker, [(100 * s + i, per_proc, LANDMARKS, BLOBS, 1) for i in range(procs)])
        wall = time.perf_counter() - t0
    updates = procs * per_proc * BLOBS * timed
    value = updates / wall
    sample = ("each step = %d replicas x %d particles x %d landmarks x %d blobs, 1 frame "
              "(bounded sample of config 2); %d timed steps" % (procs, per_proc, LANDMARKS, BLOBS, timed))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": timed, "warmup": min(warm, 1), "ms_per_step": 1e3 * wall / timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "config 2 (2^20 particles x 64 landmarks x 8 blobs), bounded CPU sample",
                   "particles": procs * per_proc, "landmarks": LANDMARKS, "blobs_per_frame": BLOBS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class _Clock(object):
    """Injectable ROS-like clock advanced by the bench (dt comes from the clock, prkt_core_v2.py:158)."""

    def __init__(self):
        from parakeet_slam_b200.rosless import Time
        self._Time = Time
        self.ns = 0

    def __call__(self):
        return self._Time(0, self.ns)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from parakeet_slam_b200 import _lib
    from parakeet_slam_b200.core import FastSLAM, Feature
    from parakeet_slam_b200.rosless import messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario, scan_from_observations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device()

    steps, warm = max(1, args.steps), max(3, args.warmup)
    M_local = args.particles_per_gpu
    M_total = M_local * world
    N, K = args.landmarks, BLOBS
    total_frames = 2 * (warm + steps) + 12
    scn = make_scenario("c2", num_particles=M_total, num_landmarks=N, obs_per_frame=K, frames=total_frames)
    feats = []
    for row in scn.landmarks:
        feats.append(Feature(mean=np.array(row), covar=np.identity(5) * scn.preset_covar))
    clk = _Clock()
    import random
    urng = random.Random(12345)
    exchange = None
    if world > 1:
        from parakeet_slam_b200.sharded import ShardedFastSLAM
        exchange = args.exchange
        try:
            fs = ShardedFastSLAM(feats, num_particles=M_total, dtype=args.dtype, noise="philox", seed=2024,
                                 uniform=urng.random, clock=clk, exchange=exchange, arithmetic=args.arith)
        except _lib.ParakeetLibraryError as exc:
            if exchange != "peer":
                raise
            # CUDA IPC unavailable on this box: same filter over NCCL (both are GPU paths); say so in the line
            exchange = "nccl (peer memory unavailable: %s)" % str(exc)[:120]
            fs = ShardedFastSLAM(feats, num_particles=M_total, dtype=args.dtype, noise="philox", seed=2024,
                                 uniform=urng.random, clock=clk, exchange="nccl", arithmetic=args.arith)
    else:
        fs = FastSLAM(feats, num_particles=M_total, dtype=args.dtype, noise="philox", seed=2024,
                      uniform=urng.random, clock=clk, arithmetic=args.arith)
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    frame_idx = [0]

    def step(events=None):
        t = frame_idx[0]
        frame_idx[0] += 1
        clk.ns += DT_NSEC
        if events:
            events[0].record()
        fs.motion_update(tw)
        if events:
            events[1].record()
        fs.measurement_update(scn.observations[t])
        if events:
            events[2].record()
        fs.low_variance_resample()
        if events:
            events[3].record()

    # ---- device-resident throughput ("value") -------------------------------------------------
    if world > 1:
        for _ in range(8):   # NCCL connects lazily: keep its first collectives out of the W warm-up steps
            step()
    for _ in range(warm):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_step_events = [[ev() for _ in range(4)] for _ in range(steps)]
    start, stop = ev(), ev()
    matched = evaluated = copied = 0
    start.record()
    for s in range(steps):
        step(per_step_events[s])
    stop.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = start.elapsed_time(stop)
    st = fs.stats()
    matched_frac = st["matched"] / float(max(1, st["matched"] + st["unmatched"]))
    eval_per_particle = st["evaluated"] / float(M_total if world > 1 else M_local)
    f_dup = st["blocks_copied"] / float(M_total if world > 1 else M_local)
    k_ms = [[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in per_step_events]
    ms_motion = sum(k[0] for k in k_ms) / steps
    ms_measure = sum(k[1] for k in k_ms) / steps
    ms_resample = sum(k[2] for k in k_ms) / steps

    # ---- end to end through the drop-in API ("e2e") ----------------------------------------------
    class View(object):
        last_sensor_reading = None
    view = View()
    scans = [scan_from_observations(scn.observations[frame_idx[0] + i]) for i in range(warm + steps)]

    def e2e_step(i):
        clk.ns += DT_NSEC
        view.last_sensor_reading = scans[i]
        fs.cam_cb(view)            # host VizScan -> kernel arguments (H2D), all kernels
        return fs.summary()        # D2H read of the frame's result (synchronises)
    for i in range(warm):
        e2e_step(i)
    sync_all()
    t0 = time.perf_counter()
    e_start, e_stop = ev(), ev()
    e_start.record()
    for i in range(steps):
        est = e2e_step(warm + i)
    e_stop.record()
    sync_all()
    e2e_ms = max(e_start.elapsed_time(e_stop), 1e3 * (time.perf_counter() - t0))

    # ---- the same steps with fp64 landmark algebra (the 100 %-index-parity instantiation), for the record ------
    alt_ms_total = alt_ms_measure = 0.0
    alt = None
    if args.arith == "f32" and not args.no_f64_block:
        urng_alt = random.Random(12345)
        clk_alt = _Clock()
        if world > 1:
            fs.close()
            fs_alt = ShardedFastSLAM(feats, num_particles=M_total, dtype=args.dtype, noise="philox", seed=2024,
                                     uniform=urng_alt.random, clock=clk_alt, arithmetic="f64",
                                     exchange=exchange if exchange in ("peer", "nccl") else "nccl")
        else:
            fs_alt = FastSLAM(feats, num_particles=M_total, dtype=args.dtype, noise="philox", seed=2024,
                              uniform=urng_alt.random, clock=clk_alt, arithmetic="f64")
        fs_alt.last_control = tw

        def alt_step(t, events=None):
            clk_alt.ns += DT_NSEC
            fs_alt.motion_update(tw)
            if events:
                events[0].record()
            fs_alt.measurement_update(scn.observations[t])
            if events:
                events[1].record()
            fs_alt.low_variance_resample()
        for t in range(warm + (8 if world > 1 else 0)):
            alt_step(t)
        sync_all()
        a0, a1 = ev(), ev()
        alt_events = [[ev(), ev()] for _ in range(steps)]
        a0.record()
        for s_ in range(steps):
            alt_step(warm + s_, alt_events[s_])
        a1.record()
        sync_all()
        alt_ms_total = a0.elapsed_time(a1)
        alt_ms_measure = sum(e[0].elapsed_time(e[1]) for e in alt_events) / steps
        alt = True

    # ---- reduce over ranks: max time -------------------------------------------------------------
    times = torch.tensor([ms_total, e2e_ms, ms_measure, ms_motion, ms_resample, alt_ms_total, alt_ms_measure],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms_measure, ms_motion, ms_resample, alt_ms_total, alt_ms_measure = [float(x) for x in times.cpu()]

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        updates = float(M_total) * K * steps
        value = updates / (ms_total * 1e-3)
        # algorithmic bytes of the fused measurement kernel per particle:
        hot_b = 4                                        # colour key per landmark
        rec_b = (64 if args.dtype == "f32" else 160) + 4  # cold record + its key
        m_matched = matched_frac * K
        bytes_particle = 32 + 8 + hot_b * N + rec_b * eval_per_particle + rec_b * m_matched + 8 + 4 * K
        achieved = bytes_particle * M_local / (ms_measure * 1e-3) / 1e9
        survey_bytes = 24 + (84 if args.dtype == "f32" else 164) * (N + m_matched) + 8 + 4 * K
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as fh:
                tr = json.load(fh)["measure_kernel<float,f32>" if args.arith == "f32" else "measure_kernel<float>"]
            if (args.dtype, M_local, N, K) == (tr["dtype"], tr["particles_per_gpu"], tr["landmarks"], tr["blobs"]):
                traffic = tr["dram_bytes"]
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.arith, "data": "synthetic",
            "config": {
                "workload": "BASELINE config 2: 2^20 particles x 64 landmarks, 8 bearings/frame, single B200"
                            if world == 1 and M_local == PARTICLES_PER_GPU and N == LANDMARKS else
                            "%d particles x %d landmarks, 8 bearings/frame over %d GPU(s) (config-2 shard per GPU)"
                            % (M_total, N, world),
                "particles": M_total, "particles_per_gpu": M_local, "landmarks": N, "blobs_per_frame": K,
                "landmark_storage": args.dtype,
                "arithmetic": "fp64" if args.arith == "f64" else
                              "fp32 landmark algebra in K2 (gates, Mahalanobis forms, EKF); fp64 poses, weights, resampling", "motion_noise": "philox4x32-10 on device",
                "resample": "systematic every frame, copy-on-resample (duplicates only)",
                "l2": "inputs larger than L2 (%.1f GB landmark pool per GPU vs 126 MB)"
                      % (M_local * N * rec_b / 1e9),
                "matched_fraction": matched_frac, "exact_evaluations_per_particle": eval_per_particle,
                "f_dup_last_frame": f_dup,
                "parallelism": "particle-sharded x%d" % world,
                "exchange": exchange,
            },
            "kernel_ms": {"motion": ms_motion, "measure": ms_measure, "resample_total": ms_resample},
            "roofline": {
                "bound": "hbm", "kernel": "measure_kernel<%s%s>" % ("float" if args.dtype == "f32" else "double",
                                                                   ", fp32 algebra" if args.arith == "f32" else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_kind, "traffic": traffic,
                "limiter": ("fp32 landmark algebra: instruction issue (~65 % busy) with 16 warps per SM; DRAM traffic per "
                            "launch / K2 time is the HBM utilisation -- see profiles/r1_summary.md") if args.arith == "f32"
                           else ("not HBM: fp64 dependent-latency / issue (12 warps per SM at 168 registers; DRAM "
                                 "throughput ~27 % of peak under ncu) -- see profiles/r1_summary.md"),
                "traffic_note": "DRAM bytes per launch from ncu --set full (profiles/r1_kernels_ncu.csv); captured one build "
                                "before the full-sector record stores, see profiles/r1_summary.md",
                "algorithmic_bytes_per_launch": bytes_particle * M_local,
                "algorithmic_bytes_per_particle": bytes_particle,
                "survey_aos_bytes_per_particle": survey_bytes,
                "frac_vs_survey_aos_bytes": survey_bytes * M_local / (ms_measure * 1e-3) / 1e9 / peak,
            },
            "e2e": {"value": updates / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": K * 4 * 8 + 3 * 8, "d2h_bytes_per_step": 5 * 8,
                    "ms_per_step": e2e_ms / steps, "api": "FastSLAM.cam_cb(view) + FastSLAM.summary()"},
            "gpu_launches": (KERNELS_PER_STEP if world == 1 else KERNELS_PER_STEP_PEER if exchange == "peer"
                             else KERNELS_PER_STEP_NCCL) * steps,
            "clocks": clocks,
            "summary_last": list(est),
        }
        if alt:
            line["f64_arithmetic"] = {
                "note": "same workload and steps with fp64 landmark algebra (FastSLAM(arithmetic='f64')): the instantiation "
                        "whose indices are bit-exact on every reference fixture with fp64 storage",
                "value": updates / (alt_ms_total * 1e-3), "unit": UNIT, "ms_per_step": alt_ms_total / steps,
                "measure_ms": alt_ms_measure,
                "roofline_frac": bytes_particle * M_local / (alt_ms_measure * 1e-3) / 1e9 / peak,
            }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="landmark storage type")
    ap.add_argument("--particles-per-gpu", type=int, default=PARTICLES_PER_GPU)
    ap.add_argument("--landmarks", type=int, default=LANDMARKS)
    ap.add_argument("--arith", default="f32", choices=["f64", "f32"],
                    help="arithmetic of the landmark algebra in K2 (f32 needs --dtype f32); poses, weights and "
                         "resampling are fp64 either way")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64-block", action="store_true", help="skip the secondary fp64-arithmetic measurement")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="cross-shard exchange engine of the sharded filter (N > 1)")
    args = ap.parse_args(argv)
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
