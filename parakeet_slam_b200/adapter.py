"""ROS-free adapter node with the call pattern of the reference's ``CamSlam360``
(``/root/reference/src/prkt_ros.py:17-127``) for machines that have a GPU but neither ROS nor
a copy of the reference.  It is NOT a re-implementation of the ROS node: it only wires the
in-process pub/sub of ``parakeet_slam_b200.rosless`` to the filter core the way the reference
node does, so the threading / clock contract of the drop-in core can be exercised:

* ``/cmd_vel``  Twist   -> ``core.motion_update(msg)`` then publish ``summary()``  (``prkt_ros.py:113-121``)
* ``/camera/features`` VizScan -> remember as ``last_sensor_reading``             (``prkt_ros.py:103-111``)
* main loop at 10 Hz: ``core.cam_cb(self)`` then publish ``summary()``            (``prkt_ros.py:63-85``)

With real ROS the unmodified ``prkt_ros.py`` is used instead.
"""
from __future__ import annotations

import numpy as np

from . import rosless
from .core import FastSLAM, Feature, heading_to_quaternion
from .rosless import fake_rospy as rospy
from .rosless.messages import Odometry, Twist, VizScan


def preset_map():
    """The four immutable landmarks the reference node is constructed with
    (``prkt_ros.py:33-52``): covariance 0.25 * I5, ``__immutable__ = True``."""
    cov = np.identity(5) * 0.25
    feats = []
    for mean in ([0, 25, 161, 77, 137], [10, 25, 75, 55, 230], [0, 15, 82, 120, 68], [10, 15, 224, 37, 192]):
        f = Feature(mean=np.array(mean, dtype=np.float64), covar=cov.copy())
        f.__immutable__ = True
        feats.append(f)
    return feats


class SlamNode(object):
    def __init__(self, preset_features=None, **core_kwargs):
        rospy.init_node("CAMSLAM360")
        self.last_sensor_reading = None
        self.core = FastSLAM(preset_map() if preset_features is None else preset_features, **core_kwargs)
        self.cam_sub = rospy.Subscriber("/camera/features", VizScan, self.measurement_update)
        self.twist_sub = rospy.Subscriber("/cmd_vel", Twist, self.motion_update)
        self.odom_pub = rospy.Publisher("/slam_estimate", Odometry, queue_size=1)
        self.estimates = []

    def easy_odom(self):
        x, y, heading = self.core.summary()
        otto = Odometry()
        otto.header.frame_id = "odom"
        otto.header.stamp = rospy.Time.now()
        otto.pose.pose.position.x = x
        otto.pose.pose.position.y = y
        otto.pose.pose.orientation = heading_to_quaternion(heading)
        self.estimates.append((x, y, heading))
        return otto

    def measurement_update(self, msg):
        self.last_sensor_reading = msg
        self.odom_pub.publish(self.easy_odom())

    def motion_update(self, msg):
        self.core.motion_update(msg)
        self.odom_pub.publish(self.easy_odom())

    def loop_over_particles(self):
        self.core.cam_cb(self)
        self.odom_pub.publish(self.easy_odom())

    def run(self, max_iterations=None, time_limit=40.0):
        rate = rospy.Rate(10)
        n = 0
        while not rospy.is_shutdown() and self.last_sensor_reading is None:
            if rospy.Time.now().to_sec() > time_limit:
                return 10
            rate.sleep()
        while not rospy.is_shutdown():
            self.loop_over_particles()
            n += 1
            if rospy.Time.now().to_sec() > time_limit or (max_iterations is not None and n >= max_iterations):
                return 10
            rate.sleep()
