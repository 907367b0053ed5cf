"""TEST INFRASTRUCTURE ONLY -- drive the loaded reference (``oracle.ref_shim``) over a
``parakeet_slam_b200.scenario.Scenario`` and record everything parity needs.

The frame loop is ``simple_driver``-style: the control is constant, the clock
advances ``dt`` per frame and every frame is one ``FastSLAM.cam_cb`` call
(reference ``prkt_core_v2.py:59-137``), which performs the motion update
(``:75-77``), association + EKF + weighting (``:84-124``) and the low-variance
resample (``:137``).  ``match_features_to_scan`` and ``low_variance_resample`` are
wrapped (not replaced) to record association ids, pre-resample weights and
ancestor indices.
"""
from __future__ import annotations

import random as _pyrandom
import warnings

import numpy as np

from parakeet_slam_b200.rosless import clock
from parakeet_slam_b200.scenario import Scenario, scan_from_observations

from . import ref_shim


class _View(object):
    """Stands in for ``CamSlam360``: ``cam_cb`` only reads ``last_sensor_reading``
    (reference ``prkt_core_v2.py:82``)."""

    def __init__(self):
        self.last_sensor_reading = None


def heading_of(ref, particle) -> float:
    return float(ref.utils.quaternion_to_heading(particle.state.pose.pose.orientation))


def pose_of(ref, particle):
    p = particle.state.pose.pose.position
    return (float(np.asarray(p.x).reshape(-1)[0]), float(np.asarray(p.y).reshape(-1)[0]),
            heading_of(ref, particle))


def build_reference_filter(ref, scn: Scenario, num_particles=None, known_map=True, potential_slots=()):
    """``FastSLAM`` with M particles (the reference hard-codes 50, ``:41``) preloaded with the
    scenario's true landmarks, covar ``preset_covar * I5`` (``prkt_ros.py:33-37``)."""
    core = ref.core
    M = scn.num_particles if num_particles is None else num_particles
    feats = []
    if known_map:
        for row in scn.landmarks:
            f = core.Feature(mean=ref.matrix.Matrix([float(v) for v in row]),
                             covar=ref.matrix.Matrix(np.identity(5) * scn.preset_covar))
            f.__immutable__ = bool(scn.immutable)
            feats.append(f)
    clock.set(0.0)
    fs = core.FastSLAM(feats)
    if M != fs.num_particles:
        fs.num_particles = M
        fs.particles = [core.FilterParticle() for _ in range(M)]
        for particle in fs.particles:
            particle.load_feature_list(feats)
    # some landmarks as POTENTIAL features: potential_features[-id] (:287, :679); a state the API allows
    # although the reference's own flow never reaches it (finding F5)
    for particle in fs.particles:
        for j in potential_slots:
            particle.potential_features[-(j + 1)] = particle.feature_set.pop(j + 1)
    return fs


def landmark_state(fs, n_slots):
    """[M, n, 5] means, [M, n, 5, 5] covars, [M, n] update counts for ids 1..n_slots."""
    M = len(fs.particles)
    mean = np.zeros((M, n_slots, 5))
    cov = np.zeros((M, n_slots, 5, 5))
    cnt = np.zeros((M, n_slots), dtype=np.int64)
    for i, p in enumerate(fs.particles):
        for id_, f in list(p.feature_set.items()) + list(p.potential_features.items()):
            mean[i, abs(id_) - 1] = np.asarray(f.mean, dtype=np.float64).reshape(5)
            cov[i, abs(id_) - 1] = np.asarray(f.covar, dtype=np.float64)
            cnt[i, abs(id_) - 1] = f.update_count
    return mean, cov, cnt


def potential_state(fs, n_slots):
    pot = np.zeros((len(fs.particles), n_slots), dtype=bool)
    for i, p in enumerate(fs.particles):
        for id_ in p.potential_features:
            pot[i, -id_ - 1] = True
    return pot


def spawn_state(ref, fs):
    """Spawn-mode state of every particle: landmarks keyed by SIGNED id (``feature_set`` ids > 0,
    ``potential_features`` ids < 0) as (mean[5], cov[5,5], update_count), and the orphaned readings
    ``hypothesis_set`` in insertion order as (id, x, y, heading + bearing, r, g, b)."""
    out = []
    for p in fs.particles:
        lms = {}
        for id_, f in list(p.feature_set.items()) + list(p.potential_features.items()):
            lms[int(id_)] = (np.asarray(f.mean, dtype=np.float64).reshape(5).copy(),
                             np.asarray(f.covar, dtype=np.float64).reshape(5, 5).copy(), int(f.update_count))
        orphans = []
        for id_, (state, blob) in p.hypothesis_set.items():
            pos = state.pose.pose.position
            h = float(ref.utils.quaternion_to_heading(state.pose.pose.orientation))
            orphans.append((int(id_), float(np.asarray(pos.x).reshape(-1)[0]), float(np.asarray(pos.y).reshape(-1)[0]),
                            float(blob.bearing) + h, float(blob.color.r), float(blob.color.g), float(blob.color.b)))
        out.append(dict(landmarks=lms, orphans=orphans, next_id=int(p.next_id)))
    return out


def run_reference(scn: Scenario, frames=None, num_particles=None, record_landmarks_at=(),
                  ref=None, spawn=False, known_map=True, timing=None, potential_slots=()):
    """Run the reference over ``frames`` frames.  Returns a dict of numpy traces:

    ``pose_pre``  [T, M, 3] pose after motion, before resampling (the pose the
                  measurement update saw), ``pose_post`` [T, M, 3] after resampling,
    ``assoc`` [T, M, K] ids, ``weight`` [T, M] pre-resample weights, ``ancestors``
    [T, M], ``summary`` [T, 3], ``next_id`` [T, M] (after resampling),
    ``lm_mean``/``lm_cov``/``lm_count`` dicts keyed by frame.
    """
    import time

    if ref is None:
        ref = ref_shim.load_reference(with_ros_node=False)
        if spawn:
            ref_shim.apply_spawn_patches(ref)
    core = ref.core
    T = scn.frames if frames is None else frames
    M = scn.num_particles if num_particles is None else num_particles
    K = scn.obs_per_frame

    fs = build_reference_filter(ref, scn, num_particles=M, known_map=known_map, potential_slots=potential_slots)
    twist = ref.msgs.Twist()
    twist.linear.x = scn.v
    twist.angular.z = scn.w
    fs.last_control = twist

    np.random.seed(scn.motion_seed)
    _pyrandom.seed(scn.meta.get("resample_seed", 12345))
    # the reference does ``from random import random`` at import time (``:28``); that is the
    # module-level function of the global Random instance, so seeding the module works.

    trace = dict(pose_pre=np.zeros((T, M, 3)), pose_post=np.zeros((T, M, 3)),
                 assoc=np.zeros((T, M, K), dtype=np.int32), weight=np.zeros((T, M)),
                 ancestors=np.zeros((T, M), dtype=np.int32), summary=np.zeros((T, 3)),
                 next_id=np.zeros((T, M), dtype=np.int64), lm_mean={}, lm_cov={}, lm_count={},
                 frame_seconds=np.zeros(T))

    rec = {}
    orig_match = core.FilterParticle.match_features_to_scan

    def match_wrapper(self, scan):
        out = orig_match(self, scan)
        rec["assoc"].append([int(pair[0]) for pair in out])
        rec["pose"].append(pose_of(ref, self))
        return out

    orig_resample = core.FastSLAM.low_variance_resample

    def resample_wrapper(self):
        old = list(self.particles)
        rec["weight"] = [float(np.asarray(p.weight).reshape(-1)[0]) for p in old]
        for idx, p in enumerate(old):
            p._trace_index = idx
        orig_resample(self)
        rec["ancestors"] = [p._trace_index for p in self.particles]

    core.FilterParticle.match_features_to_scan = match_wrapper
    core.FastSLAM.low_variance_resample = resample_wrapper
    view = _View()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for t in range(T):
                rec["assoc"], rec["pose"] = [], []
                clock.advance_nsec(int(round(scn.dt * 1e9)))
                view.last_sensor_reading = scan_from_observations(scn.observations[t], ref.msgs)
                t0 = time.perf_counter()
                fs.cam_cb(view)
                trace["frame_seconds"][t] = time.perf_counter() - t0
                trace["assoc"][t] = np.asarray(rec["assoc"], dtype=np.int32).reshape(M, K)
                trace["pose_pre"][t] = np.asarray(rec["pose"])
                trace["weight"][t] = rec["weight"]
                anc = rec["ancestors"]
                if len(anc) != M:
                    raise RuntimeError("reference emitted %d != %d particles" % (len(anc), M))
                trace["ancestors"][t] = anc
                trace["pose_post"][t] = [pose_of(ref, p) for p in fs.particles]
                trace["summary"][t] = [float(v) for v in fs.summary()]
                trace["next_id"][t] = [p.next_id for p in fs.particles]
                if t in record_landmarks_at and not spawn:
                    mean, cov, cnt = landmark_state(fs, scn.num_landmarks)
                    trace["lm_mean"][t], trace["lm_cov"][t], trace["lm_count"][t] = mean, cov, cnt
                    trace.setdefault("lm_potential", {})[t] = potential_state(fs, scn.num_landmarks)
                if spawn and t in record_landmarks_at:
                    trace.setdefault("spawn_state", {})[t] = spawn_state(ref, fs)
                if timing is not None and timing(t, trace["frame_seconds"][: t + 1]):
                    trace["frames_run"] = t + 1
                    break
    finally:
        core.FilterParticle.match_features_to_scan = orig_match
        core.FastSLAM.low_variance_resample = orig_resample
    trace.setdefault("frames_run", T)
    trace["filter"] = fs
    return trace


class ReferenceStepper(object):
    """The unmodified reference advanced ONE ``cam_cb`` frame at a time (for timing: ``bench.py``'s reference arm
    and ``cpu_baseline`` leg).  Nothing is recorded; ``step()`` returns the wall-clock seconds of the frame."""

    def __init__(self, scn: Scenario, ref=None, seed=None):
        if ref is None:
            ref = ref_shim.load_reference(with_ros_node=False)
        self.ref = ref
        self.scn = scn
        self.fs = build_reference_filter(ref, scn)
        twist = ref.msgs.Twist()
        twist.linear.x = scn.v
        twist.angular.z = scn.w
        self.fs.last_control = twist
        np.random.seed(scn.motion_seed if seed is None else seed)
        _pyrandom.seed(scn.meta.get("resample_seed", 12345) if seed is None else seed)
        self.view = _View()
        self.t = 0

    def step(self) -> float:
        import time
        scn = self.scn
        clock.advance_nsec(int(round(scn.dt * 1e9)))
        self.view.last_sensor_reading = scan_from_observations(scn.observations[self.t % scn.frames], self.ref.msgs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            self.fs.cam_cb(self.view)
            dt = time.perf_counter() - t0
        self.t += 1
        return dt

    @property
    def updates_per_frame(self) -> int:
        return len(self.fs.particles) * self.scn.obs_per_frame


def _replica_main(conn, scenario_kwargs, seed):
    """Worker process of ``ReplicaPool``: one reference filter, stepped on request."""
    try:
        from parakeet_slam_b200.scenario import make_scenario
        stepper = ReferenceStepper(make_scenario(**scenario_kwargs), seed=seed)
        conn.send(("ready", stepper.updates_per_frame))
        while True:
            msg = conn.recv()
            if msg == "stop":
                break
            conn.send(("done", stepper.step()))
    except Exception as exc:  # pragma: no cover - reported to the parent
        conn.send(("error", repr(exc)))
    finally:
        conn.close()


class ReplicaPool(object):
    """``n`` independent replicas of the (single-threaded) reference, one process per host core, advanced in
    lock-step: how the reference would use a whole host (SURVEY.md 8(d) "CPU reference timing")."""

    def __init__(self, n, scenario_kwargs, first_seed=1000):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for i in range(n):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_replica_main, args=(child, scenario_kwargs, first_seed + i), daemon=True)
            p.start()
            self.conns.append(parent)
            self.procs.append(p)
        self.updates_per_frame = 0
        for c in self.conns:
            kind, val = c.recv()
            if kind != "ready":
                self.close()
                raise RuntimeError("reference replica failed to start: %s" % (val,))
            self.updates_per_frame += val

    def step(self) -> float:
        """One frame on every replica; returns the wall-clock seconds until the slowest has finished."""
        import time
        t0 = time.perf_counter()
        for c in self.conns:
            c.send("step")
        for c in self.conns:
            kind, val = c.recv()
            if kind != "done":
                raise RuntimeError("reference replica failed: %s" % (val,))
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send("stop")
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
