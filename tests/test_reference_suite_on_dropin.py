"""The reference's OWN unit tests (``test_prkt_ros2.py``) and its unmodified ROS node (``prkt_ros.py``) executed
against the DROP-IN classes: ``prkt_core_v2`` resolves to ``parakeet_slam_b200.dropin.prkt_core_v2`` exactly as it
would with ``dropin/`` ahead of the reference's ``src/`` on ``sys.path``.

The reference modules come from ``/root/reference/src`` in the development container and from their byte-compiled
form ``oracle/_ref/*.bin`` (``oracle/build_ref.py``; build outputs that travel with a gpurun snapshot) on the GPU box.
Host-side helper cases run on CPU; everything that constructs a filter or evaluates ``probability_of_match`` needs
the device and is marked ``gpu``."""
import unittest
import warnings

import numpy as np
import pytest

from oracle import ref_shim

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="neither /root/reference nor oracle/_ref present")

# cases of test_prkt_ros2.py that touch the device: FastSLAM construction (RosFunctionalityTest, prktFastSLAMTest)
# and the probe-backed probability_of_match (:98-124)
DEVICE_CASES = {"test_probability_of_match_color", "test_probability_of_match_bearing"}


def _load_on_dropin():
    from parakeet_slam_b200.dropin import prkt_core_v2 as dropin
    ref = ref_shim.load_reference(with_ros_node=True, core_module=dropin)
    assert ref.core is dropin and ref.ros.FastSLAM is dropin.FastSLAM     # prkt_ros.py:13 bound to the drop-in
    return ref, ref_shim.load_reference_tests(ref)


def _run(cases):
    suite = unittest.TestSuite()
    suite.addTests(cases)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = unittest.TextTestRunner(verbosity=0, stream=open("/dev/null", "w")).run(suite)
    return res


@needs_ref
def test_reference_helper_cases_pass_on_the_dropin_classes():
    """prktFilterParticleTest (:72-423) and prktFeatureTest (:425-439) of the reference, host-side cases."""
    ref, tm = _load_on_dropin()
    loader = unittest.defaultTestLoader
    cases = [c for cls in (tm.prktFilterParticleTest, tm.prktFeatureTest) for c in loader.loadTestsFromTestCase(cls)
             if c._testMethodName not in DEVICE_CASES]
    assert len(cases) == 16
    res = _run(cases)
    assert res.testsRun == 16 and not res.failures and not res.errors, (res.failures, res.errors)


@needs_ref
@pytest.mark.gpu
def test_whole_reference_suite_passes_on_the_dropin_core():
    """All 23 cases, including ``CamSlam360()`` construction (prkt_ros.py on the device core), ``FastSLAM()``
    attribute types, ``motion_model`` (:46-69) and the two ``probability_of_match`` gates (:98-124)."""
    from parakeet_slam_b200 import rosless
    from parakeet_slam_b200.rosless import fake_rospy
    fake_rospy.reset()
    rosless.clock.set(0.0)
    np.random.seed(11)
    ref, tm = _load_on_dropin()
    loader = unittest.defaultTestLoader
    cases = [c for cls in (tm.RosFunctionalityTest, tm.prktFastSLAMTest, tm.prktFilterParticleTest, tm.prktFeatureTest)
             for c in loader.loadTestsFromTestCase(cls)]
    res = _run(cases)
    assert res.testsRun == 23 and not res.failures and not res.errors, (res.failures, res.errors)


@needs_ref
@pytest.mark.gpu
def test_unmodified_ros_node_drives_the_device_core():
    """``prkt_ros.CamSlam360`` (unmodified) + a ``simple_driver``-style publisher for 20 frames on the device core:
    same estimates as the ROS-free ``adapter.SlamNode`` fed the same messages, and the bounded particle sub-sample
    on the three debugging topics when asked for."""
    import random
    from parakeet_slam_b200 import rosless
    from parakeet_slam_b200.adapter import SlamNode
    from parakeet_slam_b200.rosless import fake_rospy
    from parakeet_slam_b200.scenario import scan_from_observations

    def drive(make_node, frames=20):
        fake_rospy.reset()
        rosless.clock.set(0.0)
        np.random.seed(1)
        random.seed(2)
        node = make_node()
        cmd = fake_rospy.Publisher("/cmd_vel", rosless.Twist, queue_size=1)
        cam = fake_rospy.Publisher("/camera/features", rosless.VizScan, queue_size=1)
        t = rosless.Twist()
        t.linear.x = 0.2
        t.angular.z = t.linear.x / 2.0                      # simple_driver.py:19-20
        obs = np.array([[1.4, 161, 77, 137], [1.2, 75, 55, 230], [0.9, 224.4, 36.9, 191.7]])
        est = []
        for _ in range(frames):
            rosless.clock.advance(1.0 / 11.0)
            cmd.publish(t)                                   # -> CamSlam360.motion_update -> core.motion_update
            cam.publish(scan_from_observations(obs))         # -> CamSlam360.measurement_update
            node.loop_over_particles()                       # -> core.cam_cb(node) + /slam_estimate
            est.append(node.core.summary())
        return node, np.array(est), dict(fake_rospy.published)

    ref, _ = _load_on_dropin()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        node, est, counts = drive(lambda: ref.ros.CamSlam360())
    from parakeet_slam_b200.core import FastSLAM
    assert isinstance(node.core, FastSLAM) and node.core.num_particles == 50
    assert counts["/slam_estimate"] == 60 and "/particle_track" not in counts
    assert np.isfinite(est).all() and 0.2 < est[-1, 0] < 0.5
    _, est2, _ = drive(lambda: SlamNode())
    assert np.array_equal(est, est2)

    # bounded sub-sample of the per-particle topics (prkt_core_v2.py:55-57, 127, 237, 242)
    seen = []
    fake_rospy.reset()
    rosless.clock.set(0.0)
    fs = FastSLAM(num_particles=1000, publish_particles=16)
    fake_rospy.Subscriber("/particle_track", rosless.Odometry, seen.append)
    fake_rospy.Subscriber("/aged_particles", rosless.Odometry, seen.append)
    fake_rospy.Subscriber("/resampled_particles", rosless.Odometry, seen.append)

    class View(object):
        last_sensor_reading = scan_from_observations(np.array([[0.3, 10.0, 20.0, 30.0]]))
    rosless.clock.advance(0.1)
    fs.cam_cb(View())
    assert len(seen) == 48 and all(m.header.frame_id == "odom" for m in seen)
