"""Dev-container only (needs /root/reference): the shim that runs the unmodified reference is
sound -- the reference's own 23 unit tests pass through it -- and the NumPy restatement matches the
LIVE reference on a fresh scenario that is not among the committed fixtures."""
import unittest
import warnings

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")]


def test_reference_unit_tests_pass_through_the_shim():
    ref = ref_shim.load_reference()
    tm = ref_shim.load_reference_tests(ref)
    suite = unittest.TestSuite()
    for cls in (tm.RosFunctionalityTest, tm.prktFastSLAMTest, tm.prktFeatureTest, tm.prktFilterParticleTest):
        suite.addTests(unittest.defaultTestLoader.loadTestsFromTestCase(cls))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = unittest.TextTestRunner(verbosity=0, stream=open("/dev/null", "w")).run(suite)
    assert res.testsRun == 23 and not res.failures and not res.errors


def test_restatement_matches_live_reference():
    from oracle import fastslam_np as onp, ref_driver
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c1", num_particles=12, frames=10, trajectory="corridor", num_landmarks=12,
                        world_seed=31, obs_seed=32, motion_seed=33, resample_seed=34, sigma_color=1.0)
    tr = ref_driver.run_reference(scn, record_landmarks_at=(9,))
    to = onp.run_scenario(scn, record_landmarks_at=(9,))
    assert np.array_equal(tr["assoc"], to["assoc"]) and np.array_equal(tr["ancestors"], to["ancestors"])
    assert np.max(np.abs(tr["pose_post"] - to["pose_post"])) < 1e-12
    assert np.max(np.abs(tr["weight"] - to["weight"]) / np.maximum(tr["weight"], 1e-300)) < 1e-9
    assert np.max(np.abs(tr["lm_mean"][9] - to["lm_mean"][9])) < 1e-9
    assert np.max(np.abs(tr["lm_cov"][9] - to["lm_cov"][9])) < 1e-10


def test_unmodified_ros_node_runs_on_the_fakes():
    """prkt_ros.CamSlam360 + a simple_driver-style publisher, in process, on the fake rospy."""
    from parakeet_slam_b200 import rosless
    from parakeet_slam_b200.rosless import fake_rospy
    from parakeet_slam_b200.scenario import scan_from_observations
    fake_rospy.reset()
    rosless.clock.set(0.0)
    ref = ref_shim.load_reference(with_ros_node=True)
    np.random.seed(1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        node = ref.ros.CamSlam360()
        cmd = ref.rospy.Publisher("/cmd_vel", ref.msgs.Twist, queue_size=1)
        cam = ref.rospy.Publisher("/camera/features", ref.msgs.VizScan, queue_size=1)
        t = ref.msgs.Twist()
        t.linear.x = 0.2
        t.angular.z = t.linear.x / 2.0                      # simple_driver.py:19-20
        obs = np.array([[0.3, 161, 77, 137], [1.2, 75, 55, 230]])
        for _ in range(3):
            rosless.clock.advance(1.0 / 11.0)
            cmd.publish(t)                                   # -> CamSlam360.motion_update
            cam.publish(scan_from_observations(obs, ref.msgs))
            node.loop_over_particles()                       # -> core.cam_cb(node)
        x, y, h = node.core.summary()
    assert np.isfinite([x, y, h]).all() and 0.0 < x < 0.2
    assert fake_rospy.published["/slam_estimate"] >= 9
