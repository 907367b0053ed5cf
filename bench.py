#!/usr/bin/env python
"""Benchmark of the FastSLAM hot path (BASELINE.json metric: particle x observation updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one filter frame over one batch of synthetic ``synth360`` input: motion update, fused association +
EKF + weight kernel, low-variance resample (scan, ancestors, copy-on-resample).  N = 1 runs BASELINE config 2
(2^20 particles x 64 landmarks, 8 bearings/frame, fp32 landmark storage, on-device Philox motion noise).  N > 1
(under torchrun) shards particles over ranks with the per-GPU work fixed ("weak" scaling): N * 2^20 particles in one
filter; before anything is timed every rank checks that the sharded filter reproduces the single-GPU filter bit for
bit (``sharded_identical``), and at N = 8 (or with ``--config4``) BASELINE config 4's shard (2^21 x 256 per GPU) is
measured as well.

Timing.  A BLOCK is exactly ``--steps`` frames bracketed by a barrier + ``torch.cuda.synchronize()`` on both sides
and timed with CUDA events on the launching stream (max over ranks).  One block of a sub-millisecond frame is a few
milliseconds, far too short to be repeatable, so blocks are repeated until they add up to >= 1 s of device time and
the MEDIAN block is reported (``ms_per_step`` = median block / steps; ``blocks`` says how many -- whole
re-initialisation cycles of the scenario, because frames get dearer as duplicates grow within a cycle --, ``block_ms``
their spread).  The throughput passes (``value``, ``e2e``) carry nothing but the frames between their two events;
``kernel_ms`` comes from a second, shorter pass with four events per step.  On one GPU the block copies of a
resampling run on the filter's copy stream beside the next motion update (``kernel_ms_note``).  Clocks are sampled over
the whole repeated window.

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput, ``e2e`` the same metric through the drop-in
Python API (``FastSLAM.cam_cb`` with host ``VizScan`` messages plus ``summary()`` read back every frame).
``roofline`` describes the dominant kernel (the fused measurement update) timed with CUDA events inside the timed
steps; ``roofline.pattern_ceiling`` puts the measured cost of that kernel's memory accesses alone beside it
(``tools/k2_mem_probe.cu``).  ``cpu_baseline`` times the UNMODIFIED reference (byte-compiled into ``oracle/_ref`` by ``oracle/build_ref.py``)
on the host cores beside it.
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
# launches per frame: motion, measure, weight_scan, thresholds, resample_plan, free_list_fused, assign,
# copy_blocks
KERNELS_PER_STEP = 8
KERNELS_PER_STEP_PEER = 13   # weight_scan(+all-gather), thresholds(+barrier+plan), plan, push headers, push blocks, flag post, free list, assign (local), copy, assign (arrivals, +barrier wait), copy
KERNELS_PER_STEP_NCCL = 16   # + 2 x (pack headers, pack blocks), offspring window, unpack (plus 2 NCCL collectives)
REF_SAMPLE_PARTICLES = 16    # particles per replica of the bounded reference sample (config-2 map, 8 blobs)


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
                 "--format=csv,noheader,nounits", "-lms", "50"],
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
        sm, smax, power, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(smax),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (oracle/_ref or /root/reference) and the NumPy port of it
# --------------------------------------------------------------------------------------------------
def _port_worker(args):
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


def port_baseline(particles_per_proc=2048, frames=8, procs=None, landmarks=LANDMARKS, blobs=BLOBS):
    """Throughput of the vectorised NumPy restatement (oracle/fastslam_np.py), one replica per host core."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    jobs = [(1000 + i, particles_per_proc, landmarks, blobs, frames) for i in range(procs)]
    with ctx.Pool(procs) as pool:
        pool.map(_port_worker, [(1, 64, landmarks, blobs, 1)] * procs)  # import + warm-up
        t0 = time.perf_counter()
        pool.map(_port_worker, jobs)
        wall = time.perf_counter() - t0
    updates = procs * particles_per_proc * blobs * frames
    return {"value": updates / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "%d replicas x %d particles x %d landmarks x %d blobs x %d frames of the config-2 scenario, "
                      "NumPy port (oracle/fastslam_np.py), %.1f s wall" % (procs, particles_per_proc, landmarks, blobs,
                                                                            frames, wall)}


def _ref_scenario_kwargs(frames):
    return dict(name="c2", num_particles=REF_SAMPLE_PARTICLES, num_landmarks=LANDMARKS, obs_per_frame=BLOBS,
                frames=max(frames, 4), trajectory="corridor")


def reference_config1(frames=3):
    """The unmodified reference on BASELINE config 1 (100 particles x 20 landmarks x 8 blobs), ONE process (it is
    single-threaded by construction): T-circle, where unmatched blobs make ``hypothesis_set`` grow, and T-corridor,
    fully matched (SURVEY.md 8(d) "CPU reference timing").  First frame excluded."""
    from oracle import ref_driver
    from parakeet_slam_b200.scenario import make_scenario
    out = {}
    for traj in ("circle", "corridor"):
        st = ref_driver.ReferenceStepper(make_scenario("c1", frames=frames + 1, trajectory=traj))
        secs = [st.step() for _ in range(frames + 1)][1:]
        out[traj] = {"updates_per_s_per_core": st.updates_per_frame / statistics.mean(secs),
                     "frame_s": statistics.mean(secs), "frames": frames}
    return out


def reference_baseline(steps=12, warm=1, procs=None, with_config1=True):
    """The UNMODIFIED reference (oracle/ref_shim executing its code from /root/reference or from the compiled
    oracle/_ref) on a bounded sample of config 2, one replica per host core in lock-step frames."""
    from oracle import ref_driver, ref_shim
    procs = procs or os.cpu_count() or 1
    pool = ref_driver.ReplicaPool(procs, _ref_scenario_kwargs(steps + warm))
    try:
        for _ in range(warm):
            pool.step()
        secs = [pool.step() for _ in range(steps)]
    finally:
        pool.close()
    wall = sum(secs)
    value = pool.updates_per_frame * steps / wall
    out = {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "origin": ref_shim.origin(),
           "per_core": value / procs, "step_s": wall / steps,
           "sample": "%d replicas (one per host core) x %d particles x %d landmarks x %d blobs, %d lock-step frames of the "
                     "config-2 scenario after %d warm-up, unmodified prkt_core_v2.FastSLAM.cam_cb, %.1f s wall"
                     % (procs, REF_SAMPLE_PARTICLES, LANDMARKS, BLOBS, steps, warm, wall)}
    if with_config1:
        out["config1_single_process"] = reference_config1()
    return out, secs


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same metric and
    config (a bounded sample of it per step).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref_shim
    steps = max(1, min(args.steps, 60))      # ~0.5 s per step: the whole run stays within a few minutes
    warm = max(0, min(args.warmup, 3))
    procs = os.cpu_count() or 1
    workload = ("BASELINE config 2 (2^20 particles x 64 landmarks, 8 bearings/frame), bounded CPU sample: each step = "
                "one frame of %d replicas x %d particles" % (procs, REF_SAMPLE_PARTICLES))
    if ref_shim.available():
        base, secs = reference_baseline(steps=steps, warm=warm, procs=procs)
        ms_per_step = 1e3 * sum(secs) / steps
        extra = {"port": port_baseline(frames=4, procs=procs)}
    else:
        # neither /root/reference nor oracle/_ref: time the NumPy port and say so
        base = port_baseline(particles_per_proc=1024, frames=steps, procs=procs)
        base["reason"] = "reference bytecode (oracle/_ref) absent on this box; NumPy port timed instead"
        ms_per_step = 1e3 * procs * 1024 * BLOBS / base["value"]
        workload = workload.replace("%d particles" % REF_SAMPLE_PARTICLES, "1024 particles")
        extra = {}
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "particles": procs * REF_SAMPLE_PARTICLES, "landmarks": LANDMARKS,
                   "blobs_per_frame": BLOBS},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line.update(extra)
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


class Runner(object):
    """One filter + its synthetic input + block timing."""

    EPOCH = 240   # frames before the filter is re-initialised (outside the timed blocks), see maybe_reset()

    def __init__(self, args, world, rank, M_local, N, dtype, arith, exchange, num_colors=None, scenario="c2",
                 frames=EPOCH):
        import random
        import numpy as np
        import torch
        from parakeet_slam_b200 import _lib
        from parakeet_slam_b200.core import FastSLAM, Feature
        from parakeet_slam_b200.rosless import messages
        from parakeet_slam_b200.scenario import make_scenario
        self.torch, self.world, self.rank = torch, world, rank
        self.M_local, self.M_total, self.N, self.K = M_local, M_local * world, N, BLOBS
        self.dtype, self.arith = dtype, arith
        # `frames` consecutive frames of the scenario; the filter is re-initialised before it would run past them
        self.scn = make_scenario(scenario, num_particles=self.M_total, num_landmarks=N, obs_per_frame=self.K,
                                 frames=frames, num_colors=num_colors,
                                 layout="polar" if scenario == "c2" else None)
        feats = [Feature(mean=np.array(row), covar=np.identity(5) * self.scn.preset_covar) for row in self.scn.landmarks]
        self.feats = feats
        self.clk = _Clock()
        urng = random.Random(12345)
        self.exchange = None
        kw = dict(num_particles=self.M_total, dtype=dtype, noise="philox", seed=2024, uniform=urng.random, clock=self.clk,
                  arithmetic=arith)
        if world > 1:
            from parakeet_slam_b200.sharded import ShardedFastSLAM
            self.exchange = exchange
            try:
                self.fs = ShardedFastSLAM(feats, exchange=exchange, **kw)
            except _lib.ParakeetLibraryError as exc:
                if exchange != "peer":
                    raise
                # CUDA IPC unavailable on this box: same filter over NCCL (both are GPU paths); say so in the line
                self.exchange = "nccl (peer memory unavailable: %s)" % str(exc)[:120]
                self.fs = ShardedFastSLAM(feats, exchange="nccl", **kw)
        else:
            self.fs = FastSLAM(feats, **kw)
        self.tw = messages.Twist()
        self.tw.linear.x, self.tw.angular.z = self.scn.v, self.scn.w
        self.fs.last_control = self.tw
        self.frame = 0
        self.skew = None

    def close(self):
        if hasattr(self.fs, "close"):
            self.fs.close()
        self.fs = None
        self.torch.cuda.empty_cache()

    def sync_all(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def maybe_reset(self, steps, settle=4):
        """Called between timed blocks: when the next block would run past the scenario's last frame, put every particle
        back at the origin with the preset map (FastSLAM.__init__'s state) and let `settle` untimed frames pass.  A
        long run would otherwise drift out of the regime BASELINE describes (the matched fraction of
        mutable landmarks decays after ~250 frames, which makes frames cheaper)."""
        if self.frame + steps <= self.scn.frames:
            return
        self.fs._load_presets(self.feats)
        self.fs._frame = 0
        self.frame = 0
        for _ in range(settle):
            self.step()

    def step(self, events=None):
        from parakeet_slam_b200.scenario import DT_NSEC
        fs = self.fs
        t = self.frame % self.scn.frames
        self.frame += 1
        self.clk.ns += DT_NSEC
        if events:
            events[0].record()
        fs.motion_update(self.tw)
        if events:
            # single GPU: the previous frame's block copies run on the filter's copy stream beside this motion update;
            # K2 waits for them anyway (FastSLAM._pool), waiting here keeps them out of K2's interval
            fs.wait_blocks()
            events[1].record()
        fs.measurement_update(self.scn.observations[t])
        if self.skew is not None:
            fs.pose[:, 3] *= self.skew[self.frame % 2]
        if events:
            events[2].record()
        fs.low_variance_resample()
        if events:
            events[3].record()

    def measure(self, step_fn, steps, warm, min_total_ms=1000.0, max_blocks=200, kernel_events=True):
        """Warm up, then repeat blocks of exactly `steps` steps until they total >= min_total_ms of device time.
        Returns dict(block_ms=[...max over ranks...], kernel_ms={...median over blocks...})."""
        torch = self.torch
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(warm):
            step_fn(None)
        self.sync_all()
        blocks, kms = [], []
        nblocks = None
        while True:
            per_step = [[ev() for _ in range(4)] for _ in range(steps)] if kernel_events else None
            start, stop = ev(), ev()
            self.maybe_reset(steps)
            self.sync_all()
            start.record()
            for s in range(steps):
                step_fn(per_step[s] if kernel_events else None)
            self.fs.wait_blocks()   # the last frame's block copies belong to the block
            stop.record()
            self.sync_all()
            t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device="cuda")
            if kernel_events:
                k = [[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in per_step]
                t = torch.cat([t, torch.tensor([sum(x[i] for x in k) / steps for i in range(3)], dtype=torch.float64,
                                               device="cuda")])
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t = [float(v) for v in t.cpu()]
            blocks.append(t[0])
            if kernel_events:
                kms.append(t[1:])
            if nblocks is None:   # every rank sees the same (max-reduced) first block: same decision everywhere
                nblocks = int(min(max_blocks, max(3, math.ceil(min_total_ms / max(t[0], 1e-3)))))
                # frames get dearer within a re-initialisation cycle (duplicates grow): time whole cycles, so that the
                # medians of two passes (device-resident, e2e) are taken over the same mix of blocks
                cycle = (self.scn.frames - 4) // steps   # maybe_reset(): 4 settling frames, then whole blocks
                if cycle > 1 and nblocks > cycle:
                    nblocks = int(min(max_blocks, cycle * math.ceil(nblocks / cycle)))
            if len(blocks) >= nblocks:
                break
        out = {"block_ms": blocks, "median_ms": statistics.median(blocks)}
        if kernel_events:
            out["kernel_ms"] = {n: statistics.median(k[i] for k in kms) for i, n in
                                enumerate(("motion", "measure", "resample_total"))}
        return out

    def stats(self):
        st = self.fs.stats()
        tot = float(self.M_total)
        matched = st["matched"] / float(max(1, st["matched"] + st["unmatched"]))
        return st, matched, st["evaluated"] / tot, st["blocks_copied"] / tot


def k2_bytes_per_particle(dtype, N, K, evals_pp, matched_frac):
    """Algorithmic bytes of the fused measurement kernel per particle."""
    rec_b = (64 if dtype == "f32" else 160) + 4      # cold record + its key
    return 32 + 8 + 4 * N + rec_b * evals_pp + rec_b * matched_frac * K + 8 + 4 * K


def traffic_from_profiles(arith, dtype, M_local, N, K):
    """DRAM bytes per launch of K2 from the committed ncu capture of this build, if it describes this workload."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                tr = json.load(fh)["measure_kernel<double>" if dtype == "f64" else
                                   "measure_kernel<float,f32>" if arith == "f32" else "measure_kernel<float>"]
            if (dtype, M_local, N, K) == (tr["dtype"], tr["particles_per_gpu"], tr["landmarks"], tr["blobs"]):
                return tr["dram_bytes"], name
        except Exception:
            continue
    return None, None


def pattern_ceiling(args, M_local, N, K, ms_measure):
    """What K2's access pattern costs with NO arithmetic (tools/k2_mem_probe.cu, committed result), beside this run's K2."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_k2_mem_probe.json")) as fh:
            pr = json.load(fh)
        if (args.dtype, args.arith, M_local, N, K) != ("f32", "f32", pr["particles"], pr["landmarks"], pr["blobs"]):
            return None
        return {"ms": pr["all_ms"], "gbs": pr["all_gbs"], "k2_over_ceiling": ms_measure / pr["all_ms"],
                "source": "profiles/r2_k2_mem_probe.json (tools/k2_mem_probe.cu: keys + 8 record reads + 8 write-backs + "
                          "pose + ids per particle through a 4-stage cp.async ring, 16 warps/SM, permuted slots)"}
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from parakeet_slam_b200 import _lib
    from parakeet_slam_b200.scenario import DT_NSEC, scan_from_observations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device()

    steps, warm = max(1, args.steps), max(3, args.warmup)
    M_local, N, K = args.particles_per_gpu, args.landmarks, BLOBS
    peaks, peak_kind = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- N > 1: the sharded filter must equal the single-GPU filter before its throughput means anything -------
    selfcheck = None
    if world > 1 and not args.no_selfcheck:
        from parakeet_slam_b200.selfcheck import sharded_equals_single
        try:
            selfcheck = sharded_equals_single(exchange=args.exchange)
        except _lib.ParakeetLibraryError as exc:
            selfcheck = sharded_equals_single(exchange="nccl")
            selfcheck["note"] = "peer memory unavailable (%s): checked over NCCL" % str(exc)[:80]
        if not selfcheck["identical"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "sharded filter differs from the single-GPU filter",
                                  "sharded_identical": False, "selfcheck": selfcheck}))
            dist.destroy_process_group()
            return 3

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident throughput ("value") ------------------------------------------------------------------
    R = Runner(args, world, rank, M_local, N, args.dtype, args.arith, args.exchange)
    if world > 1:
        for _ in range(8):   # NCCL connects lazily: keep its first collectives out of the W warm-up steps
            R.step()
    # two passes: the throughput from blocks with NOTHING but the frames between the two block events (as in the e2e
    # pass), then the per-kernel split from shorter blocks that carry four events per step (they cost ~2 %)
    main = R.measure(R.step, steps, warm, kernel_events=False)
    main["kernel_ms"] = R.measure(R.step, steps, warm, min_total_ms=300.0)["kernel_ms"]
    st, matched_frac, eval_pp, f_dup = R.stats()
    ms_step = main["median_ms"] / steps
    exchange = R.exchange

    # ---- end to end through the drop-in API ("e2e") --------------------------------------------------------------
    class View(object):
        last_sensor_reading = None
    view = View()
    scans = [scan_from_observations(o) for o in R.scn.observations]

    def e2e_step(_events):
        R.clk.ns += DT_NSEC
        view.last_sensor_reading = scans[R.frame % len(scans)]
        R.frame += 1
        R.fs.cam_cb(view)            # host VizScan -> kernel arguments (H2D), all kernels
        e2e_step.last = R.fs.summary()   # D2H read of the frame's result (synchronises)
    e2e = R.measure(e2e_step, steps, warm, kernel_events=False)
    e2e_ms_step = e2e["median_ms"] / steps

    updates_step = float(R.M_total) * K
    bytes_pp = k2_bytes_per_particle(args.dtype, N, K, eval_pp, matched_frac)
    ms_measure = main["kernel_ms"]["measure"]
    achieved = bytes_pp * M_local / (ms_measure * 1e-3) / 1e9
    traffic, traffic_src = traffic_from_profiles(args.arith, args.dtype, M_local, N, K)

    # ---- resampler: copy-on-resample at stated duplicate fractions (single GPU) ----------------------------------
    sweep = None
    if world == 1 and not args.quick:
        sweep = []
        fs = R.fs
        idx = torch.arange(R.M_total, device="cuda")
        live_block = 4 * N + (64 if args.dtype == "f32" else 160) * N
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for label, dead in (("1%", idx % 100 == 0), ("10%", idx % 10 == 0), ("50%", idx % 2 == 1), ("94% (full gather: "
                            "every 16th particle survives)", idx % 16 != 0)):
            ts, copied = [], 0
            for rep in range(5):
                fs.pose[:, 3] = torch.where(dead, 0.0, 1.0).to(torch.float64)
                a, b = ev(), ev()
                a.record()
                fs.low_variance_resample()
                fs.wait_blocks()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
                copied = int(fs._n_copied.item())
            t_ms = statistics.median(ts[1:])
            sweep.append({"target": label, "f_dup": copied / float(R.M_total), "resample_ms": t_ms,
                          "blocks_copied": copied,
                          "copy_gbs_lower_bound": 2.0 * live_block * copied / (t_ms * 1e-3) / 1e9})

    line = {
        "metric": METRIC, "value": updates_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.arith, "data": "synthetic",
        "blocks": len(main["block_ms"]),
        "block_ms": {"min": min(main["block_ms"]), "median": main["median_ms"], "max": max(main["block_ms"])},
        "config": {
            "workload": "BASELINE config 2: 2^20 particles x 64 landmarks, 8 bearings/frame, single B200"
                        if world == 1 and M_local == PARTICLES_PER_GPU and N == LANDMARKS else
                        "%d particles x %d landmarks, 8 bearings/frame over %d GPU(s) (config-2 shard per GPU)"
                        % (R.M_total, N, world),
            "particles": R.M_total, "particles_per_gpu": M_local, "landmarks": N, "blobs_per_frame": K,
            "landmark_storage": args.dtype,
            "arithmetic": "fp64" if args.arith == "f64" else
                          "fp32 landmark algebra in K2 (gates, Mahalanobis forms, EKF); fp64 poses, weights, resampling",
            "motion_noise": "philox4x32-10 on device",
            "resample": "systematic every frame, copy-on-resample (duplicates only)",
            "timing": "median of %d blocks of %d steps (>= 1 s of device time in total, whole re-initialisation cycles), "
                      "CUDA events, max over ranks; per-kernel split from a second pass with four events per step"
                      % (len(main["block_ms"]), steps),
            "l2": "inputs larger than L2 (%.1f GB landmark pool per GPU vs 126 MB)"
                  % (M_local * N * ((64 if args.dtype == "f32" else 160) + 4) / 1e9),
            "matched_fraction": matched_frac, "exact_evaluations_per_particle": eval_pp,
            "f_dup_last_frame": f_dup,
            "parallelism": "particle-sharded x%d" % world,
            "exchange": exchange,
        },
        "kernel_ms": main["kernel_ms"],
        "roofline": {
            "bound": "hbm", "kernel": "measure_kernel<%s%s>" % ("float" if args.dtype == "f32" else "double",
                                                               ", fp32 algebra" if args.arith == "f32" else ""),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_kind, "traffic": traffic, "traffic_source": traffic_src,
            "limiter": "DRAM row activations of the scattered 64-byte record reads and write-backs: a probe that issues K2's "
                       "accesses with no arithmetic (tools/k2_mem_probe.cu) needs 0.39 ms per 2^20 particles, whatever the "
                       "store policy; on the SM side 16 warps per SM at ~65 % of the issue slots.  DRAM traffic is 1.06x the "
                       "algorithmic bytes -- see profiles/r2_summary.md"
                       if args.arith == "f32" else
                       "fp64 dependent latency (12 warps per SM at 168 registers) -- see profiles/r2_summary.md",
            "algorithmic_bytes_per_launch": bytes_pp * M_local,
            "algorithmic_bytes_per_particle": bytes_pp,
            "pattern_ceiling": pattern_ceiling(args, M_local, N, K, ms_measure),
        },
        "kernel_ms_note": "motion = the motion update beside the previous frame's block copies (copy stream, single GPU); "
                          "resample_total = weight scan ... permutation, block copies not included",
        "e2e": {"value": updates_step / (e2e_ms_step * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": K * 4 * 8 + 3 * 8, "d2h_bytes_per_step": 5 * 8,
                "ms_per_step": e2e_ms_step, "blocks": len(e2e["block_ms"]),
                "api": "FastSLAM.cam_cb(view) + FastSLAM.summary()"},
        "gpu_launches": (KERNELS_PER_STEP if world == 1 else KERNELS_PER_STEP_PEER if exchange == "peer"
                         else KERNELS_PER_STEP_NCCL) * steps,
        "summary_last": list(e2e_step.last),
    }
    if world > 1:
        mig = torch.tensor([st.get("migrated_in", 0)], dtype=torch.int64, device="cuda")
        dist.all_reduce(mig)
        line["config"]["migrations_last_frame_all_ranks"] = int(mig.item())
        if getattr(R.fs, "_timing_on", False):
            line["peer_chain_ms"] = R.fs.timing_report()
        if exchange == "peer":
            # time every rank spent inside the two flag barriers of a frame (accumulated by the kernels themselves):
            # what lock-step costs on top of the kernels
            bw = R.fs.barrier_wait_report()
            t = torch.tensor([bw["totals_barrier_ms"], bw["pushes_barrier_ms"]], dtype=torch.float64, device="cuda")
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            line["barrier_wait_ms_per_rank"] = {"totals": [round(float(a[0]), 4) for a in allt],
                                                "pushes": [round(float(a[1]), 4) for a in allt],
                                                "frames": bw["frames"]}
    if selfcheck is not None:
        line["sharded_identical"] = selfcheck["identical"]
        line["selfcheck"] = selfcheck
    if sweep is not None:
        line["resample_f_dup_sweep"] = sweep
    R.close()
    del R

    def variant(M_loc, Nv, dtype, arith, note, num_colors=None, scenario="c2", skew=None, min_total_ms=400.0):
        """The same steps on another instantiation / workload (shorter: >= 0.4 s of device time)."""
        V = Runner(args, world, rank, M_loc, Nv, dtype, arith, args.exchange, num_colors=num_colors, scenario=scenario)
        if skew is not None:
            gidx = V.fs.particle_offset + torch.arange(V.fs.num_particles, device="cuda")
            odd = (gidx // M_loc) % 2 == 1
            V.skew = [torch.where(odd, skew, 1.0).to(torch.float64), torch.where(odd, 1.0, skew).to(torch.float64)]
        if world > 1:
            for _ in range(4):
                V.step()
        m = V.measure(V.step, steps, warm, min_total_ms=min_total_ms, kernel_events=False)
        m["kernel_ms"] = V.measure(V.step, steps, warm, min_total_ms=min_total_ms / 2)["kernel_ms"]
        stv, mf, ev_pp, fd = V.stats()
        b = k2_bytes_per_particle(dtype, Nv, K, ev_pp, mf)
        ms = m["median_ms"] / steps
        out = {"note": note, "value": float(V.M_total) * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
               "particles_per_gpu": M_loc, "landmarks": Nv, "landmark_storage": dtype, "arithmetic": arith,
               "kernel_ms": m["kernel_ms"], "blocks": len(m["block_ms"]), "matched_fraction": mf,
               "exact_evaluations_per_particle": ev_pp, "f_dup_last_frame": fd,
               "roofline_frac": b * M_loc / (m["kernel_ms"]["measure"] * 1e-3) / 1e9 / peak,
               "algorithmic_bytes_per_particle": b}
        if world > 1:
            out["migrations_last_frame"] = stv.get("migrated_in")
            mig = torch.tensor([stv.get("migrated_in", 0)], dtype=torch.int64, device="cuda")
            dist.all_reduce(mig)
            out["migrations_last_frame_all_ranks"] = int(mig.item())
        V.close()
        return out

    variants = {}
    if not args.quick:
        if args.arith == "f32":
            variants["f64_arithmetic"] = variant(
                M_local, N, args.dtype, "f64",
                "same workload with fp64 landmark algebra on the fp32 records (FastSLAM(arithmetic='f64'))")
        if world == 1:
            variants["f64_storage_f64_arithmetic"] = variant(
                M_local, N, "f64", "f64",
                "same workload, fp64 records AND fp64 algebra: the instantiation whose indices are bit-exact against the "
                "reference on every fixture (FastSLAM(dtype='f64'))")
            variants["ambiguous_colours"] = variant(
                M_local, N, args.dtype, args.arith,
                "6 colours shared by the 64 landmarks (~11 colour-compatible landmarks per blob, all evaluated exactly; the "
                "association is decided by the position likelihood)", num_colors=6)
    if (world == 8 and not args.no_config4) or args.config4:
        c4 = {}
        c4["balanced"] = variant(
            args.config4_particles, 256, args.dtype, args.arith,
            "BASELINE config 4 shard: 2^21 particles x 256 landmarks per GPU (16 M over 8 GPUs), known map, corridor",
            scenario="c4")
        if world > 1:
            c4["skewed"] = variant(
                args.config4_particles, 256, args.dtype, args.arith,
                "same, weights of alternating ranks' particles scaled by 0.8 every frame: ~11 % of every other shard's "
                "offspring (17 KB records) cross NVLink per frame", scenario="c4", skew=0.8)
        variants["config4"] = c4
    line.update(variants)

    clocks = sampler.stop() if rank == 0 else None
    line["clocks"] = clocks
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_shim
            if ref_shim.available():
                line["cpu_baseline"], _ = reference_baseline(steps=12, warm=1)
                line["cpu_baseline"]["port"] = port_baseline()
            else:
                line["cpu_baseline"] = port_baseline()
                line["cpu_baseline"]["reason"] = "reference bytecode (oracle/_ref) absent on this box"
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="landmark storage type")
    ap.add_argument("--particles-per-gpu", type=int, default=PARTICLES_PER_GPU)
    ap.add_argument("--landmarks", type=int, default=LANDMARKS)
    ap.add_argument("--arith", default="f32", choices=["f64", "f32"],
                    help="arithmetic of the landmark algebra in K2 (f32 needs --dtype f32); poses, weights and "
                         "resampling are fp64 either way")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="main measurement and e2e only (no variants, no sweep)")
    ap.add_argument("--no-f64-block", action="store_true", help="alias of --quick (kept for older command lines)")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the sharded == single-GPU check at N > 1")
    ap.add_argument("--config4", action="store_true", help="also measure BASELINE config 4's shard (default at N = 8)")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--config4-particles", type=int, default=1 << 21, help="particles per GPU of the config-4 block")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="cross-shard exchange engine of the sharded filter (N > 1)")
    args = ap.parse_args(argv)
    args.quick = args.quick or args.no_f64_block
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
