"""The drop-in core behind a ROS-style adapter (call pattern of prkt_ros.CamSlam360): Twist callbacks
that integrate the previous control over the clock delta, cam_cb frames with dt == 0 motion, summary
after every call -- compared with the NumPy oracle driven by the same call sequence."""
import math
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_call_sequence_matches_oracle():
    from oracle import fastslam_np as onp
    from parakeet_slam_b200 import rosless
    from parakeet_slam_b200.adapter import SlamNode, preset_map
    from parakeet_slam_b200.rosless import fake_rospy, messages
    from parakeet_slam_b200.scenario import scan_from_observations

    fake_rospy.reset()
    rosless.clock.set(0.0)
    M = 50                                              # the reference's particle count (:41)
    np.random.seed(5)
    random.seed(6)
    node = SlamNode(num_particles=M, dtype="f64")
    cmd = fake_rospy.Publisher("/cmd_vel", messages.Twist)
    cam = fake_rospy.Publisher("/camera/features", messages.VizScan)

    # oracle twin, driven by the same sequence of calls
    np_state = np.random.get_state()
    lm = np.array([np.asarray(f.mean, dtype=float) for f in preset_map()])
    st = onp.OracleState(M, lm, preset_covar=0.25, immutable=True)
    urng = random.Random(6)
    nrng = np.random.RandomState()
    nrng.set_state(np_state)
    last_control = (0.0, 0.0)
    last_update = 0.0
    est = []

    def o_motion(new, now):
        nonlocal last_control, last_update
        dt = now - last_update
        st.pose = onp.motion_update(st.pose, nrng.standard_normal((M, 3)), last_control[0], last_control[1], dt)
        last_update = last_update + dt
        last_control = new

    tw = messages.Twist()
    tw.linear.x, tw.angular.z = 0.2, 0.1
    obs = np.array([[math.atan2(25, 0) - 0.0, 161, 77, 137], [math.atan2(15, 10), 224.3, 36.8, 192.1],
                    [0.4, 10.0, 10.0, 10.0]])
    for it in range(6):
        rosless.clock.advance_nsec(90909091)
        now = rosless.clock.now().to_sec()
        cmd.publish(tw)                                  # node.motion_update -> core.motion_update
        o_motion((0.2, 0.1), now)
        est.append(onp.summary(st.pose))
        cam.publish(scan_from_observations(obs))        # node.measurement_update: stores the scan
        est.append(onp.summary(st.pose))
        node.loop_over_particles()                       # core.cam_cb(node)
        o_motion(last_control, now)                      # cam_cb's own motion_update, dt == 0 (:75-77)
        onp.measurement_update(st, obs)
        anc = onp.resample_sequential(st.weight, urng.random())
        onp.apply_ancestors(st, anc)
        est.append(onp.summary(st.pose))
    got = np.array(node.estimates)
    want = np.array(est)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 1e-9
    # immutable presets: maps never change, weights are the importance factors
    p = node.core.particles[0]
    assert sorted(p.feature_set) == [1, 2, 3, 4] and p.feature_set[1].__immutable__
    assert p.feature_set[2].update_count == 0
    with pytest.raises(KeyError):
        p.get_feature_by_id(9)
    assert p.next_id == st.next_id[0]


def test_reference_unit_cases_on_device():
    """Known answers of the reference's own unit tests (test_prkt_ros2.py) for the helpers that the
    drop-in evaluates on the device."""
    from parakeet_slam_b200.core import FastSLAM, Feature, FilterParticle
    from parakeet_slam_b200.rosless import Duration, messages
    particle = FilterParticle()
    state = messages.Odometry()                          # zero quaternion -> heading 0
    # test_probability_of_match_color :98-110 -- colours far apart -> exactly 0.0
    blob = messages.Blob()
    blob.color.r = 255
    feature = Feature(mean=np.array([1, 0, 0, 0, 0]))
    assert particle.probability_of_match(state, blob, feature) == 0.0
    # test_probability_of_match_bearing :112-124 -- bearing far off -> exactly 0.0
    blob = messages.Blob()
    blob.bearing = math.pi
    assert particle.probability_of_match(state, blob, feature) == 0.0
    # aligned, same colour -> a positive likelihood, and match_one picks that feature
    blob = messages.Blob()
    v = particle.probability_of_match(state, blob, feature)
    assert v > 0.0
    particle.feature_set[7] = Feature(mean=np.array([1, 0, 200, 0, 0]))
    particle.feature_set[3] = feature
    assert particle.match_one(state, blob) == 3
    assert particle.match_features_to_scan(messages.VizScan(observes=[blob]))[0][0] == 3
    # test_initilization :39-44, test_motion_model :46-69 (statistical bound, heading 0)
    fs = FastSLAM()
    assert isinstance(fs.last_control, messages.Twist) and isinstance(fs.Qt, np.ndarray) and len(fs.particles) == 50
    fpold = FilterParticle()
    fpold.state.pose.pose.position.y = 2.0
    twist = messages.Twist()
    twist.linear.x = 1
    fpnew = fs.motion_model(fpold, twist, Duration.from_sec(.1))
    assert abs((fpnew.state.pose.pose.position.y - 2.0) - 0.0) < .01
    assert abs((fpnew.state.pose.pose.position.x - 0.0) - 0.1) < 6 * 0.0505   # drive noise sigma = .05*v + .0005 (:185)
    # test_initialization :73-80, test_get_feature_by_id :82-96, prktFeatureTest :426-431
    p = FilterParticle()
    assert p.weight == 1 and p.next_id == 1 and p.feature_set == {} and p.hypothesis_set == {}
    f = Feature()
    assert f.update_count == 0 and f.mean.shape == (5,) and f.covar.shape == (5, 5) and f.identity.shape == (5, 5)


def test_v1_names_route_into_the_v2_pipeline():
    """prkt_core.py's names (the module itself is un-importable in the reference, finding F10)."""
    import sys
    import numpy as np
    from parakeet_slam_b200.rosless import clock, messages
    sys.path.insert(0, os.path.join(ROOT, "parakeet_slam_b200", "dropin"))
    try:
        import prkt_core
    finally:
        sys.path.pop(0)
    clock.set(0.0)
    np.random.seed(3)
    slam = prkt_core.ParticleMixedSlam(spawn=True, capacity=8, orphan_capacity=8)
    assert slam.M == 10 and len(slam.robot_particles) == 10
    tw = messages.Twist()
    tw.linear.x = 0.2
    slam.motion_update(tw)
    obs = messages.Blob()
    obs.bearing, obs.color.r, obs.color.g, obs.color.b = 0.3, 120.0, 30.0, 200.0
    for k in range(3):
        clock.advance(1.0 / 11.0)
        obs.bearing = 0.3 + 0.05 * k
        slam.measurement_update(obs)                      # -> cam_observation_update
    x, y, h = slam.summary()
    assert np.isfinite([x, y, h]).all() and x > 0.0
    p = slam.robot_particles[0]
    # reading 1 is orphaned (id 1), reading 2 pairs with it into potential landmark -2, reading 3 either associates
    # with that landmark (no id consumed) or is orphaned in its turn
    assert p.next_id in (3, 4) and len(p.hypothesis_set) >= 1
    assert len(p.hypothesis_set) + len(p.potential_features) + len(p.feature_set) == p.next_id - 1
    assert prkt_core.RobotParticle is prkt_core.FilterParticle and prkt_core.FeatureModel is prkt_core.Feature
