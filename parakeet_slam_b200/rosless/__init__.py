"""ROS-free stand-ins for the handful of ROS Indigo modules the reference filter
touches (``rospy``, ``geometry_msgs.msg``, ``nav_msgs.msg``, ``viz_feature_sim.msg``
and ``tf.transformations``).

They exist for two reasons:

* the device filter must run without a ROS install (the drop-in module falls
  back to these message classes when ``rospy`` is not importable), and
* the *unmodified* reference sources (``prkt_ros.py``, ``prkt_core_v2.py`` ...)
  can be loaded in-process on top of them (``install()``), which is how the
  parity oracle and the adapter harness run without ROS.

Only the fields the reference reads are modelled:
``Blob{bearing,size,color{r,g,b}}``, ``VizScan{observes}``, nested
``Odometry.pose.pose.position/orientation``, ``Twist.linear.x / angular.z``,
``Time``/``Duration`` arithmetic (reference ``prkt_core_v2.py:158,165,174``).
"""
from __future__ import annotations

import sys
import types

from . import messages, rostime, transformations  # noqa: F401
from .messages import (Blob, ColorRGBA, Header, Observation, Odometry, Point, Pose,
                       PoseWithCovariance, Quaternion, Twist, TwistWithCovariance,
                       Vector3, VizScan)
from .rostime import Duration, Rate, Time, clock

__all__ = [
    "Blob", "ColorRGBA", "Header", "Observation", "Odometry", "Point", "Pose",
    "PoseWithCovariance", "Quaternion", "Twist", "TwistWithCovariance", "Vector3",
    "VizScan", "Duration", "Rate", "Time", "clock", "install", "uninstall",
]

_FAKE_NAMES = ("rospy", "geometry_msgs", "geometry_msgs.msg", "nav_msgs", "nav_msgs.msg",
               "viz_feature_sim", "viz_feature_sim.msg", "tf", "tf.transformations",
               "std_msgs", "std_msgs.msg")


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__rosless__ = True
    return mod


def build_modules() -> dict:
    """Return {module name: module} for every fake ROS module."""
    from . import fake_rospy

    rospy = _module("rospy", **{k: getattr(fake_rospy, k) for k in fake_rospy.__all__})
    gm_msg = _module("geometry_msgs.msg", Twist=Twist, Quaternion=Quaternion, Point=Point,
                     Pose=Pose, Vector3=Vector3, PoseWithCovariance=PoseWithCovariance,
                     TwistWithCovariance=TwistWithCovariance)
    gm = _module("geometry_msgs", msg=gm_msg)
    nm_msg = _module("nav_msgs.msg", Odometry=Odometry)
    nm = _module("nav_msgs", msg=nm_msg)
    vf_msg = _module("viz_feature_sim.msg", Blob=Blob, VizScan=VizScan, Observation=Observation)
    vf = _module("viz_feature_sim", msg=vf_msg)
    sm_msg = _module("std_msgs.msg", Header=Header, ColorRGBA=ColorRGBA)
    sm = _module("std_msgs", msg=sm_msg)
    tf = _module("tf", transformations=transformations)
    return {
        "rospy": rospy,
        "geometry_msgs": gm, "geometry_msgs.msg": gm_msg,
        "nav_msgs": nm, "nav_msgs.msg": nm_msg,
        "viz_feature_sim": vf, "viz_feature_sim.msg": vf_msg,
        "std_msgs": sm, "std_msgs.msg": sm_msg,
        "tf": tf, "tf.transformations": transformations,
    }


def install(force: bool = False) -> dict:
    """Register the fake modules in ``sys.modules``.

    A real ROS install always wins unless ``force`` is set: if ``rospy`` is
    already importable nothing is touched and ``{}`` is returned.
    """
    if not force:
        try:
            import rospy  # noqa: F401
            if not getattr(sys.modules["rospy"], "__rosless__", False):
                return {}
        except ImportError:
            pass
    mods = build_modules()
    sys.modules.update(mods)
    return mods


def uninstall() -> None:
    for name in _FAKE_NAMES:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__rosless__", False):
            del sys.modules[name]
