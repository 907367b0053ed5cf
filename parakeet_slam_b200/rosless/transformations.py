"""Restatement of the two ``tf.transformations`` functions the reference calls
(``utils.py:18`` euler_from_quaternion, ``utils.py:28`` quaternion_from_euler).

``tf`` is not vendored in the reference and carries no version pin
(``package.xml:50,58``); this follows the published algorithm of ROS Indigo's
``tf/transformations.py`` (C. Gohlke's transformations module: quaternion order
x,y,z,w; ``_EPS = 4*eps``; static-frame 'sxyz' axes), operation for operation, so
the floating-point result of a heading round trip is the one the reference
would see.  Only the 'sxyz' convention is implemented.
"""
from __future__ import annotations

import math

import numpy

_EPS = numpy.finfo(float).eps * 4.0


def quaternion_from_euler(ai, aj, ak, axes="sxyz"):
    if axes != "sxyz":
        raise NotImplementedError("only the 'sxyz' convention is restated")
    ai /= 2.0
    aj /= 2.0
    ak /= 2.0
    ci = math.cos(ai)
    si = math.sin(ai)
    cj = math.cos(aj)
    sj = math.sin(aj)
    ck = math.cos(ak)
    sk = math.sin(ak)
    cc = ci * ck
    cs = ci * sk
    sc = si * ck
    ss = si * sk
    quaternion = numpy.empty((4,), dtype=numpy.float64)
    quaternion[0] = cj * sc - sj * cs
    quaternion[1] = cj * ss + sj * cc
    quaternion[2] = cj * cs - sj * sc
    quaternion[3] = cj * cc + sj * ss
    return quaternion


def quaternion_matrix(quaternion):
    q = numpy.array(quaternion[:4], dtype=numpy.float64, copy=True)
    nq = numpy.dot(q, q)
    if nq < _EPS:
        return numpy.identity(4)
    q *= math.sqrt(2.0 / nq)
    q = numpy.outer(q, q)
    return numpy.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
        (0.0, 0.0, 0.0, 1.0)), dtype=numpy.float64)


def euler_from_matrix(matrix, axes="sxyz"):
    if axes != "sxyz":
        raise NotImplementedError("only the 'sxyz' convention is restated")
    i, j, k = 0, 1, 2
    M = numpy.asarray(matrix, dtype=numpy.float64)[:3, :3]
    cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
    if cy > _EPS:
        ax = math.atan2(M[k, j], M[k, k])
        ay = math.atan2(-M[k, i], cy)
        az = math.atan2(M[j, i], M[i, i])
    else:
        ax = math.atan2(-M[j, k], M[j, j])
        ay = math.atan2(-M[k, i], cy)
        az = 0.0
    return ax, ay, az


def euler_from_quaternion(quaternion, axes="sxyz"):
    return euler_from_matrix(quaternion_matrix(quaternion), axes)
