"""On-device accuracy analysis: the working form of the reference's
``analyze_slam.py:1-36`` (squared x / y error between the estimate and the truth, accumulated per frame)
with the error helpers of ``utils.py:83-213`` (``heading_error``, ``minimize_angle``, ``dist``), plus what a
particle filter needs on top: the effective sample size and the per-landmark map error.  All reductions
run over the device-resident particles (``pk_accuracy``, ``pk_map_error``); only a few scalars come back."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib


def accuracy(fs, true_pose):
    """Spread of the particle cloud about the true pose and weight statistics of the last measurement update.

    Returns a dict: ``rms_x``, ``rms_y``, ``rms_heading`` (root mean square over particles),
    ``mean_x``, ``mean_y``, ``n_eff`` = (sum w)^2 / sum w^2, ``sum_w``, ``max_w``."""
    torch, lib = fs._torch, fs._lib
    M = fs.num_particles
    with fs._lock, fs._on_device():
        out = torch.zeros((8,), dtype=torch.float64, device=fs._device)
        ws = torch.zeros((8 * 512,), dtype=torch.float64, device=fs._device)
        _lib.check(lib.pk_accuracy(_lib.ptr(fs.pose), M, float(true_pose[0]), float(true_pose[1]), float(true_pose[2]),
                                   _lib.ptr(out), _lib.ptr(ws), fs._stream()), "pk_accuracy")
        o = out.cpu().numpy()
    sw, sw2 = float(o[0]), float(o[1])
    return dict(sum_w=sw, n_eff=(sw * sw / sw2) if sw2 > 0.0 else 0.0, max_w=float(o[7]),
                rms_x=math.sqrt(o[2] / M), rms_y=math.sqrt(o[3] / M), rms_heading=math.sqrt(o[4] / M),
                mean_x=float(o[5]) / M, mean_y=float(o[6]) / M)


def map_error(fs, true_landmarks):
    """Per true landmark j (reference id j+1): RMS position error of that landmark over the particles that hold it,
    and how many do.  Returns (rms [N], count [N])."""
    torch, lib = fs._torch, fs._lib
    truth = np.ascontiguousarray(true_landmarks, dtype=np.float64).reshape(-1, 5)
    N = truth.shape[0]
    with fs._lock, fs._on_device():
        t = torch.from_numpy(truth).to(fs._device)
        err2 = torch.zeros((N,), dtype=torch.float64, device=fs._device)
        cnt = torch.zeros((N,), dtype=torch.int64, device=fs._device)
        _lib.check(lib.pk_map_error(_lib.ptr(fs._pool), fs.capacity, fs._dt, _lib.ptr(fs.slot), _lib.ptr(fs.aux),
                                    fs.num_particles, _lib.ptr(t), N, _lib.ptr(err2), _lib.ptr(cnt), fs._stream()),
                   "pk_map_error")
        e, c = err2.cpu().numpy(), cnt.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        rms = np.sqrt(np.where(c > 0, e / np.maximum(c, 1), np.nan))
    return rms, c


class SlamAnalyzer(object):
    """``analyze_slam.py``'s loop: feed (truth, estimate) pairs, read the accumulated error.  ``estimate`` is
    ``FastSLAM.summary()``; ``report()`` returns what the reference logs (``:36``: square roots of the summed
    squared x and y errors) together with the per-frame RMS and the ``calc_errors`` triple of the last pair."""

    def __init__(self):
        self.x_squared = 0.0           # analyze_slam.py:24
        self.y_squared = 0.0           # :25
        self.count = 0                 # :26
        self.last = None
        self.n_eff = []

    def add(self, truth, estimate, n_eff=None):
        self.count += 1                                                   # :30
        self.x_squared += math.pow(truth[0] - estimate[0], 2)             # :31
        self.y_squared += math.pow(truth[1] - estimate[1], 2)             # :32
        self.last = (tuple(truth), tuple(estimate))
        if n_eff is not None:
            self.n_eff.append(float(n_eff))

    @staticmethod
    def calc_errors(location, goal):
        """``utils.calc_errors`` (``utils.py:83-101``) on (x, y, heading) triples: error along the goal heading,
        signed error normal to it, heading error."""
        rx, ry = location[0] - goal[0], location[1] - goal[1]
        gx, gy = math.cos(goal[2]), math.sin(goal[2])
        along = rx * gx + ry * gy                                         # along_axis_error :103-137
        nx, ny = rx - gx * along, ry - gy * along                         # off_axis_error :139-196
        sign = -1.0 if (gx * ry - gy * rx) < 0 else 1.0
        off = 0.0 if (abs(rx) < 1e-12 and abs(ry) < 1e-12) else sign * math.sqrt(nx * nx + ny * ny)
        return along, off, location[2] - goal[2]                          # heading_error :199-205

    def report(self):
        n = max(self.count, 1)
        out = dict(frames=self.count, avg_x=math.sqrt(self.x_squared), avg_y=math.sqrt(self.y_squared),   # :36
                   rms_x=math.sqrt(self.x_squared / n), rms_y=math.sqrt(self.y_squared / n))
        if self.last is not None:
            out["along"], out["off"], out["heading"] = self.calc_errors(self.last[1], self.last[0])
        if self.n_eff:
            out["n_eff_min"], out["n_eff_last"] = min(self.n_eff), self.n_eff[-1]
        return out
