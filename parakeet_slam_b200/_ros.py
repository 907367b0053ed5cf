"""Message / clock classes used by the drop-in core: the real ROS ones when ``rospy`` is
importable (reference ``prkt_core_v2.py:14,22,26``), else the ROS-free stand-ins."""
try:  # pragma: no cover - depends on the host having ROS
    import rospy
    from geometry_msgs.msg import Quaternion, Twist
    from nav_msgs.msg import Odometry
    from viz_feature_sim.msg import Blob

    def now():
        return rospy.Time.now()

    Publisher = rospy.Publisher
    HAVE_ROS = not getattr(rospy, "__rosless__", False)
except ImportError:
    from .rosless import Blob, Odometry, Quaternion, Time, Twist

    from .rosless.fake_rospy import Publisher

    def now():
        return Time.now()

    HAVE_ROS = False

__all__ = ["Blob", "Odometry", "Publisher", "Quaternion", "Twist", "now", "HAVE_ROS"]
