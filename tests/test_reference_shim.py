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


def test_analysis_helpers_match_reference_utils():
    """parakeet_slam_b200.analysis.SlamAnalyzer.calc_errors against the reference's utils.calc_errors
    (utils.py:83-205) on random (location, goal) pairs; host arithmetic on both sides."""
    from parakeet_slam_b200.analysis import SlamAnalyzer
    ref = ref_shim.load_reference(with_ros_node=False)
    rs = np.random.RandomState(4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(200):
            loc = (rs.uniform(-5, 5), rs.uniform(-5, 5), rs.uniform(-3.1, 3.1))
            goal = (rs.uniform(-5, 5), rs.uniform(-5, 5), rs.uniform(-3.1, 3.1))
            want = ref.utils.calc_errors(ref.utils.easy_Odom(loc[0], loc[1], loc[2]),
                                         ref.utils.easy_Odom(goal[0], goal[1], goal[2]))
            got = SlamAnalyzer.calc_errors(loc, goal)
            assert np.allclose(got, want, rtol=0, atol=1e-12), (loc, goal, got, want)


def test_spawn_restatement_matches_live_patched_reference():
    """Spawn mode on a scenario that is not among the fixtures: the NumPy restatement against the reference run with
    the three patches of SURVEY.md A.6 (ids incl. new negative ones, ancestors, next_id, orphaned readings)."""
    from oracle import fastslam_np as onp, ref_driver
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c3", num_particles=10, num_landmarks=16, frames=25, obs_per_frame=6,
                        world_seed=41, obs_seed=42, motion_seed=43, resample_seed=44)
    ref = ref_shim.load_reference(with_ros_node=False)
    ref_shim.apply_spawn_patches(ref)
    tr = ref_driver.run_reference(scn, ref=ref, spawn=True, known_map=False, record_landmarks_at=(24,))
    to = onp.run_scenario(scn, spawn=True, known_map=False, capacity=64, record_landmarks_at=(24,))
    assert np.array_equal(tr["assoc"], to["assoc"]) and np.array_equal(tr["ancestors"], to["ancestors"])
    assert np.array_equal(tr["next_id"], to["next_id"])
    assert (tr["assoc"] < 0).any() and (tr["assoc"] > 0).any()
    for i, ps in enumerate(tr["spawn_state"][24]):
        ids = to["lm_ids"][24][i]
        mine = {int(ids[j]): to["lm_mean"][24][i, j] for j in range(len(ids)) if ids[j] != 0}
        assert set(mine) == set(ps["landmarks"])
        for id_, (mean, cov, cnt) in ps["landmarks"].items():
            assert np.max(np.abs(mean - mine[id_])) < 1e-9
        assert [o[0] for o in to["orphans"][24][i]] == [o[0] for o in ps["orphans"]]
