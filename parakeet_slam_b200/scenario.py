"""Synthetic 360-degree bearing+colour scenario "synth360" (SURVEY.md section 8(d)).

Replaces the un-vendored ``viz_feature_sim`` simulator: it produces, per frame,
the K x (bearing, r, g, b) observation block that a ``VizScan`` carries
(reference ``matrix.py:35-39``, ``prkt_core_v2.py:344``).  All randomness is
NumPy legacy ``RandomState`` / Python ``random`` so that the reference (run as
the oracle) and the device filter see identical inputs:

* world                seed 2024  (landmark positions and colours)
* observation noise    seed 7
* motion noise         seed 12345 (standard normals ``[frames, M, 3]``, consumed in
                       particle order, three per particle: drive, heading-1,
                       heading-2 -- reference ``prkt_core_v2.py:185,190,193``)
* resampling uniform   ``random.seed(12345)``, one ``random()`` per frame
                       (reference ``prkt_core_v2.py:226``)

The control input is ``simple_driver.py:15-20``'s: 11 Hz, v = 0.2, omega = v/2 = 0.1
("T-circle"); "T-corridor" is the same with omega = 0.
"""
from __future__ import annotations

import math
import random as _pyrandom
from dataclasses import dataclass, field

import numpy as np

# simple_driver.py:15 rospy.Rate(11); ROS time is integer nanoseconds, so the dt the filter
# sees (``Duration.to_sec()``, prkt_core_v2.py:158,174) is the ns-quantised period.
DT_NSEC = 90909091
DT = DT_NSEC / 1e9


@dataclass
class Scenario:
    name: str
    num_particles: int
    num_landmarks: int
    obs_per_frame: int
    frames: int
    v: float
    w: float
    dt: float
    landmarks: np.ndarray          # [N, 5]  x, y, r, g, b
    true_poses: np.ndarray         # [frames, 3]
    observations: np.ndarray       # [frames, K, 4]  bearing, r, g, b
    obs_landmark: np.ndarray       # [frames, K]  index of the true landmark of each blob
    u01: np.ndarray                # [frames]  resampling uniforms
    motion_seed: int = 12345
    preset_covar: float = 0.25
    immutable: bool = False
    meta: dict = field(default_factory=dict)

    def motion_noise(self, frame_lo: int = 0, frame_hi: int | None = None,
                     num_particles: int | None = None) -> np.ndarray:
        """Standard normals ``[frame_hi-frame_lo, M, 3]`` of the motion stream."""
        M = self.num_particles if num_particles is None else num_particles
        hi = self.frames if frame_hi is None else frame_hi
        rs = np.random.RandomState(self.motion_seed)
        if frame_lo:
            # skip whole frames in bounded chunks (legacy gauss caches in pairs: 3*M even or
            # odd both keep the stream aligned as long as we draw the same count)
            remaining = frame_lo * M * 3
            while remaining > 0:
                n = min(remaining, 1 << 22)
                rs.standard_normal(n)
                remaining -= n
        return rs.standard_normal((hi - frame_lo, M, 3))

    def motion_noise_stream(self, num_particles: int | None = None):
        """Generator yielding one ``[M, 3]`` block per frame from one stream."""
        M = self.num_particles if num_particles is None else num_particles
        rs = np.random.RandomState(self.motion_seed)
        for _ in range(self.frames):
            yield rs.standard_normal((M, 3))


def wrap_pi(a):
    return (a + math.pi) % (2.0 * math.pi) - math.pi


def make_world(num_landmarks: int, layout: str, frames: int, v: float, w: float,
               seed: int = 2024, num_colors: int | None = None) -> np.ndarray:
    rs = np.random.RandomState(seed)
    lm = np.empty((num_landmarks, 5), dtype=np.float64)
    if layout == "polar":
        cy = v / w if w != 0.0 else 2.0
        ang = rs.uniform(0.0, 2.0 * math.pi, num_landmarks)
        rad = rs.uniform(3.0, 9.0, num_landmarks)
        lm[:, 0] = 0.0 + rad * np.cos(ang)
        lm[:, 1] = cy + rad * np.sin(ang)
    elif layout == "corridor":
        lm[:, 0] = rs.uniform(-5.0, 0.02 * frames + 5.0, num_landmarks)
        lm[:, 1] = rs.uniform(-8.0, 8.0, num_landmarks)
    else:
        raise ValueError("unknown layout %r" % (layout,))
    lm[:, 2:5] = rs.uniform(0.0, 255.0, (num_landmarks, 3))
    if num_colors:
        # colour-ambiguous world: landmark j wears palette colour j % num_colors, so every blob is colour-compatible
        # with num_landmarks / num_colors landmarks and the association is decided by the position likelihood
        palette = np.random.RandomState(seed + 1).uniform(20.0, 235.0, (num_colors, 3))
        lm[:, 2:5] = palette[np.arange(num_landmarks) % num_colors]
    return lm


def make_scenario(name: str = "c1", *, num_particles: int | None = None,
                  num_landmarks: int | None = None, obs_per_frame: int = 8,
                  frames: int | None = None, trajectory: str | None = None,
                  layout: str | None = None, sigma_bearing: float = 0.02,
                  sigma_color: float = 0.3, immutable: bool = False,
                  world_seed: int = 2024, obs_seed: int = 7, motion_seed: int = 12345,
                  resample_seed: int = 12345, num_colors: int | None = None) -> Scenario:
    """Build one of the BASELINE.json configurations (or a scaled variant).

    ``name``: "c1" (100 x 20, 500 frames, T-circle), "c2" (2^20 x 64, T-corridor),
    "c3" (2^22 x 256), "c4" (2^24 x 256), "c5" (2^20 x 1024, 10k frames).
    Any field can be overridden for scaled-down parity cases.
    """
    presets = {
        "c1": dict(M=100, N=20, frames=500, trajectory="circle", layout="polar"),
        "c2": dict(M=1 << 20, N=64, frames=220, trajectory="corridor", layout="polar"),
        "c3": dict(M=1 << 22, N=256, frames=200, trajectory="corridor", layout="corridor"),
        "c4": dict(M=1 << 24, N=256, frames=200, trajectory="corridor", layout="corridor"),
        "c5": dict(M=1 << 20, N=1024, frames=10000, trajectory="corridor", layout="corridor"),
    }
    p = presets[name]
    M = p["M"] if num_particles is None else num_particles
    N = p["N"] if num_landmarks is None else num_landmarks
    T = p["frames"] if frames is None else frames
    traj = p["trajectory"] if trajectory is None else trajectory
    lay = p["layout"] if layout is None else layout
    K = obs_per_frame
    v = 0.2
    w = 0.1 if traj == "circle" else 0.0
    dt = DT

    lm = make_world(N, lay, T, v, 0.1, seed=world_seed, num_colors=num_colors)

    poses = np.empty((T, 3), dtype=np.float64)
    x = y = th = 0.0
    rs = np.random.RandomState(obs_seed)
    obs = np.empty((T, K, 4), dtype=np.float64)
    obs_lm = np.empty((T, K), dtype=np.int64)
    Keff = min(K, N)
    for t in range(T):
        th += w * dt
        x += v * dt * math.cos(th)
        y += v * dt * math.sin(th)
        poses[t] = (x, y, th)
        d2 = (lm[:, 0] - x) ** 2 + (lm[:, 1] - y) ** 2
        order = np.argsort(d2, kind="stable")[:Keff]
        bearing = wrap_pi(np.arctan2(lm[order, 1] - y, lm[order, 0] - x) - th)
        bearing = bearing + rs.normal(0.0, sigma_bearing, Keff)
        color = lm[order, 2:5] + rs.normal(0.0, sigma_color, (Keff, 3))
        obs[t, :Keff, 0] = bearing
        obs[t, :Keff, 1:4] = color
        obs_lm[t, :Keff] = order
        if Keff < K:  # fewer landmarks than blobs: repeat the last one
            obs[t, Keff:] = obs[t, Keff - 1]
            obs_lm[t, Keff:] = obs_lm[t, Keff - 1]

    pr = _pyrandom.Random(resample_seed)
    u01 = np.array([pr.random() for _ in range(T)], dtype=np.float64)

    return Scenario(name=name, num_particles=M, num_landmarks=N, obs_per_frame=K, frames=T,
                    v=v, w=w, dt=dt, landmarks=lm, true_poses=poses, observations=obs,
                    obs_landmark=obs_lm, u01=u01, motion_seed=motion_seed,
                    immutable=immutable,
                    meta=dict(trajectory=traj, layout=lay, sigma_bearing=sigma_bearing,
                              sigma_color=sigma_color, world_seed=world_seed, obs_seed=obs_seed,
                              resample_seed=resample_seed, num_colors=num_colors))


def scan_from_observations(obs_frame: np.ndarray, msgs=None):
    """Wrap one ``[K, 4]`` observation block as a ``VizScan`` of ``Blob`` messages."""
    if msgs is None:
        from .rosless import messages as msgs
    scan = msgs.VizScan()
    for row in np.asarray(obs_frame, dtype=np.float64):
        blob = msgs.Blob()
        blob.bearing = float(row[0])
        blob.color.r = float(row[1])
        blob.color.g = float(row[2])
        blob.color.b = float(row[3])
        scan.observes.append(blob)
    return scan
