"""Scalar helper methods of the reference's ``FilterParticle`` evaluated on the device through
the probe entry points of the C ABI (``pk_probe_likelihood``, ``pk_probe_ekf``): the same
``__device__`` functions the fused measurement kernel calls, one triple per thread."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib


def obs_direction(bearing):
    """``unit((cos b, sin b, 0.0))`` as ``closest_point`` builds it (reference
    ``prkt_core_v2.py:510``, ``utils.py:69-76``), in host floats."""
    c = math.cos(bearing)
    s = math.sin(bearing)
    length = math.sqrt(c * c + s * s + 0.0 * 0.0)
    inv = 1.0 / length
    return c * inv, s * inv


def _dev(a, dtype=np.float64):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def likelihood_batch(pose3, blob4, mean5, covp, covc, params=None):
    """``probability_of_match`` for n independent (pose, blob, landmark) triples.
    pose3 [n,3], blob4 [n,4], mean5 [n,5], covp [n,2,2], covc [n,3,3] -> [n]."""
    import torch
    _lib.require_device()
    lib = _lib.load()
    n = len(pose3)
    params = params if params is not None else _lib.default_params()
    dirs = np.array([obs_direction(float(b)) for b in np.asarray(blob4)[:, 0]]).reshape(n, 2)
    t = [_dev(pose3), _dev(blob4), _dev(dirs), _dev(mean5), _dev(np.reshape(covp, (n, 4))),
         _dev(np.reshape(covc, (n, 9)))]
    out = torch.zeros((n,), dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pk_probe_likelihood(*[_lib.ptr(x) for x in t], n, ctypes.byref(params), _lib.ptr(out), st),
               "pk_probe_likelihood")
    return out.cpu().numpy()


def ekf_batch(pose2, blob4, mean5, covp, covc, meta=None, params=None):
    """One EKF update per triple -> (mean5', covp', covc', factor)."""
    import torch
    _lib.require_device()
    lib = _lib.load()
    n = len(pose2)
    params = params if params is not None else _lib.default_params()
    t = [_dev(pose2), _dev(blob4), _dev(mean5), _dev(np.reshape(covp, (n, 4))), _dev(np.reshape(covc, (n, 9)))]
    m = None if meta is None else _dev(meta, np.int32)
    o = [torch.zeros((n, k), dtype=torch.float64, device="cuda") for k in (5, 4, 9, 1)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pk_probe_ekf(*[_lib.ptr(x) for x in t], _lib.ptr(m), n, ctypes.byref(params),
                                *[_lib.ptr(x) for x in o], st), "pk_probe_ekf")
    r = [x.cpu().numpy() for x in o]
    return r[0], r[1].reshape(n, 2, 2), r[2].reshape(n, 3, 3), r[3].reshape(n)


def _state_tuple(state):
    from .core import FilterParticle
    pos = state.pose.pose.position
    q = state.pose.pose.orientation
    n = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w
    if n < np.finfo(float).eps * 4.0:
        heading = 0.0
    else:
        s = math.sqrt(2.0 / n)
        zs, ws = q.z * s, q.w * s
        heading = math.atan2(zs * ws, 1.0 - zs * zs)
    return float(pos.x), float(pos.y), heading


def probability_of_match_many(state, blob, features):
    n = len(features)
    x, y, th = _state_tuple(state)
    pose3 = np.tile([x, y, th], (n, 1))
    blob4 = np.tile([blob.bearing, blob.color.r, blob.color.g, blob.color.b], (n, 1)).astype(np.float64)
    mean5 = np.array([np.asarray(f.mean, dtype=np.float64).reshape(5) for f in features])
    cov = np.array([np.asarray(f.covar, dtype=np.float64) for f in features])
    return likelihood_batch(pose3, blob4, mean5, cov[:, :2, :2], cov[:, 2:, 2:])


def probability_of_match(state, blob, feature):
    return float(probability_of_match_many(state, blob, [feature])[0])
