"""Device-side 360-degree bearing + colour scan simulator.

Stands in for the un-vendored ``viz_feature_sim`` node, whose ``VizScan`` of ``Blob``s is what
``CamSlam360`` hands to ``FastSLAM.cam_cb`` (reference ``prkt_ros.py:17-127``, ``matrix.py:35-39``,
``prkt_core_v2.py:344``).  The rule is the ``synth360`` scenario's (``scenario.make_scenario``): the K
landmarks nearest the true pose in ascending distance order, bearing =
``wrap_pi(atan2(ly - y, lx - x) - theta) + N(0, sigma_bearing^2)``, colour = truth + ``N(0, sigma_color^2)``.

The scan is produced by ``pk_simulate_scan`` and stays on the device; ``FastSLAM.measurement_update``
accepts it as is (``pk_measurement_update_dev``), so a long-horizon run never touches the host.
``to_vizscan`` copies it back as a ``VizScan`` message for callers that want the reference's wire format.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class BearingSimulator(object):
    def __init__(self, landmarks, obs_per_frame=8, sigma_bearing=0.02, sigma_color=0.3, device=None, seed=7):
        import torch
        _lib.require_device()
        self._torch = torch
        self._lib = _lib.load()
        self._device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        lm = np.ascontiguousarray(landmarks, dtype=np.float64).reshape(-1, 5)
        if not 1 <= obs_per_frame <= _lib.PK_MAX_OBS:
            raise ValueError("obs_per_frame must be in [1, %d]" % _lib.PK_MAX_OBS)
        self.num_landmarks = lm.shape[0]
        self.obs_per_frame = int(obs_per_frame)
        self.sigma_bearing = float(sigma_bearing)
        self.sigma_color = float(sigma_color)
        self.seed = int(seed)
        self.frame = 0
        self._lm = torch.from_numpy(lm).to(self._device)
        self._ws = torch.zeros((self.num_landmarks,), dtype=torch.float64, device=self._device)
        self.last_landmarks = torch.zeros((self.obs_per_frame,), dtype=torch.int32, device=self._device)

    def scan(self, pose, noise=None, out=None):
        """One frame: ``pose`` = true (x, y, theta); ``noise`` = ``[K, 4]`` standard normals (host array or
        device tensor; column 0 perturbs the bearing, 1..3 the colour) or None for on-device noise.
        Returns the device tensor ``[K, 4]`` (bearing, r, g, b)."""
        torch, lib = self._torch, self._lib
        K = self.obs_per_frame
        if out is None:
            out = torch.empty((K, 4), dtype=torch.float64, device=self._device)
        nptr = 0
        if noise is not None:
            if not hasattr(noise, "data_ptr"):
                noise = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64).reshape(K, 4)).to(self._device)
            nptr = _lib.ptr(noise)
        st = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        with torch.cuda.device(self._device):
            _lib.check(lib.pk_simulate_scan(_lib.ptr(self._lm), self.num_landmarks, float(pose[0]), float(pose[1]),
                                            float(pose[2]), K, nptr, self.seed, self.frame, self.sigma_bearing,
                                            self.sigma_color, _lib.ptr(self._ws), _lib.ptr(out),
                                            _lib.ptr(self.last_landmarks), st), "pk_simulate_scan")
        self.frame += 1
        return out

    @staticmethod
    def to_vizscan(obs):
        """Device scan -> ``VizScan`` of ``Blob`` messages (one D2H copy)."""
        from .scenario import scan_from_observations
        return scan_from_observations(obs.cpu().numpy() if hasattr(obs, "cpu") else obs)
