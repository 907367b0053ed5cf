"""Plain-Python message classes with the field names the reference reads.

Default-constructed ``Odometry()`` carries the all-zero quaternion (0,0,0,0),
as genpy does; the reference's tests rely on that reading back as heading 0
(``test_prkt_ros2.py:100,230`` through ``utils.py:8-19``).
"""
from __future__ import annotations


class _Msg(object):
    __slots__ = ()

    def __repr__(self):
        body = ", ".join("%s=%r" % (s, getattr(self, s)) for s in self.__slots__)
        return "%s(%s)" % (type(self).__name__, body)

    def __eq__(self, other):
        return type(other) is type(self) and all(
            getattr(self, s) == getattr(other, s) for s in self.__slots__)

    def __ne__(self, other):
        return not self == other

    __hash__ = None


class Vector3(_Msg):
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class Point(Vector3):
    __slots__ = ()


class Quaternion(_Msg):
    __slots__ = ("x", "y", "z", "w")

    def __init__(self, x=0.0, y=0.0, z=0.0, w=0.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class Twist(_Msg):
    __slots__ = ("linear", "angular")

    def __init__(self, linear=None, angular=None):
        self.linear = linear if linear is not None else Vector3()
        self.angular = angular if angular is not None else Vector3()


class Pose(_Msg):
    __slots__ = ("position", "orientation")

    def __init__(self, position=None, orientation=None):
        self.position = position if position is not None else Point()
        self.orientation = orientation if orientation is not None else Quaternion()


class PoseWithCovariance(_Msg):
    __slots__ = ("pose", "covariance")

    def __init__(self, pose=None, covariance=None):
        self.pose = pose if pose is not None else Pose()
        self.covariance = covariance if covariance is not None else [0.0] * 36


class TwistWithCovariance(_Msg):
    __slots__ = ("twist", "covariance")

    def __init__(self, twist=None, covariance=None):
        self.twist = twist if twist is not None else Twist()
        self.covariance = covariance if covariance is not None else [0.0] * 36


class Header(_Msg):
    __slots__ = ("seq", "stamp", "frame_id")

    def __init__(self, seq=0, stamp=None, frame_id=""):
        self.seq, self.stamp, self.frame_id = seq, stamp, frame_id


class Odometry(_Msg):
    __slots__ = ("header", "child_frame_id", "pose", "twist")

    def __init__(self):
        self.header = Header()
        self.child_frame_id = ""
        self.pose = PoseWithCovariance()
        self.twist = TwistWithCovariance()


class ColorRGBA(_Msg):
    __slots__ = ("r", "g", "b", "a")

    def __init__(self, r=0.0, g=0.0, b=0.0, a=0.0):
        self.r, self.g, self.b, self.a = r, g, b, a


class Blob(_Msg):
    """viz_feature_sim/Blob: only bearing and color.{r,g,b} are read
    (reference ``matrix.py:35-39``, ``prkt_core_v2.py:409,425-427``)."""
    __slots__ = ("bearing", "size", "color")

    def __init__(self, bearing=0.0, size=0.0, color=None):
        self.bearing = bearing
        self.size = size
        self.color = color if color is not None else ColorRGBA()


class VizScan(_Msg):
    """viz_feature_sim/VizScan: ``observes`` is the list of Blobs
    (reference ``prkt_core_v2.py:344``)."""
    __slots__ = ("header", "observes")

    def __init__(self, observes=None):
        self.header = Header()
        self.observes = observes if observes is not None else []


class Observation(_Msg):
    """Only imported by the dead v1 core (``prkt_core.py``); never read."""
    __slots__ = ("bearing", "color")

    def __init__(self):
        self.bearing = 0.0
        self.color = ColorRGBA()
