"""The tools around the filter: device-side scan simulator, the device-resident
scan path of the fused kernel, and the on-device accuracy analysis.  Needs a B200."""
import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clocked_filter(scn, M, dtype, **kw):
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    urng = random.Random(3)
    fs = FastSLAM(make_features(scn) if kw.pop("known_map", True) else [], num_particles=M, dtype=dtype, noise="philox",
                  seed=21, uniform=urng.random, clock=clk, **kw)
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    return fs, clk, tw


@pytest.mark.parametrize("layout,K,N", [("corridor", 8, 256), ("polar", 8, 20), ("polar", 12, 5), ("corridor", 64, 1024)])
def test_simulator_reproduces_the_scenario(layout, K, N):
    """Same landmark choice and order as scenario.make_scenario (stable K nearest), colours bit-identical
    (one multiply-add each), bearings within 1e-12 (device atan2 vs numpy's)."""
    from parakeet_slam_b200.scenario import make_scenario
    from parakeet_slam_b200.simulator import BearingSimulator
    T = 25
    scn = make_scenario("c3" if layout == "corridor" else "c1", num_particles=4, num_landmarks=N, frames=T,
                        obs_per_frame=K, layout=layout, trajectory="circle" if layout == "polar" else "corridor")
    sim = BearingSimulator(scn.landmarks, K, scn.meta["sigma_bearing"], scn.meta["sigma_color"])
    rs = np.random.RandomState(scn.meta["obs_seed"])      # the scenario's observation-noise stream
    Keff = min(K, N)
    for t in range(T):
        zb = rs.standard_normal(Keff)                     # rs.normal(0, s, n) == s * standard_normal(n) on this stream
        zc = rs.standard_normal((Keff, 3))
        noise = np.zeros((K, 4))
        noise[:Keff, 0], noise[:Keff, 1:] = zb, zc
        obs = sim.scan(scn.true_poses[t], noise).cpu().numpy()
        idx = sim.last_landmarks.cpu().numpy()
        assert np.array_equal(idx, scn.obs_landmark[t]), "frame %d" % t
        assert np.array_equal(obs[:, 1:], scn.observations[t][:, 1:])
        assert np.max(np.abs(obs[:, 0] - scn.observations[t][:, 0])) < 1e-12
    # on-device noise: right landmarks, noise of the right size
    obs = sim.scan(scn.true_poses[0]).cpu().numpy()
    assert np.array_equal(sim.last_landmarks.cpu().numpy(), scn.obs_landmark[0])
    d = obs[:Keff, 1:] - scn.landmarks[scn.obs_landmark[0][:Keff], 2:5]
    assert 0.0 < np.abs(d).max() < 6 * scn.meta["sigma_color"]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("spawn", [False, True])
def test_device_scan_path_equals_host_path(dtype, spawn):
    """pk_measurement_update_dev (+ pk_spawn_update_dev) on a device-resident scan against the host-scan entry
    points on the same numbers: identical ids and ancestors, weights and maps to rounding (the blob direction is
    built with CUDA's sincos instead of libm's)."""
    import torch
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    M, T = 4096, 8
    scn = make_scenario("c3", num_particles=M, num_landmarks=40, frames=T, obs_per_frame=8)
    kw = dict(spawn=True, capacity=48, orphan_capacity=32, known_map=False) if spawn else {}
    runs = []
    for on_device in (False, True):
        fs, clk, tw = _clocked_filter(scn, M, dtype, **dict(kw))
        fs.keep_trace = True
        trace = []
        for t in range(T):
            clk.ns += DT_NSEC
            fs.motion_update(tw)
            obs = scn.observations[t]
            fs.measurement_update(torch.from_numpy(obs.copy()).cuda() if on_device else obs)
            w = fs.last_weight.cpu().numpy()
            fs.low_variance_resample()
            trace.append((fs.last_assoc.cpu().numpy(), fs.last_ancestors.cpu().numpy(), w, fs.stats()))
        runs.append((trace, fs.export_maps(), fs.aux.cpu().numpy()))
    (tr_h, maps_h, aux_h), (tr_d, maps_d, aux_d) = runs
    for t in range(T):
        assert np.array_equal(tr_h[t][0], tr_d[t][0]), "assoc, frame %d" % t
        assert np.array_equal(tr_h[t][1], tr_d[t][1]), "ancestors, frame %d" % t
        big = tr_h[t][2] > 1e-300
        assert np.max(np.abs(tr_h[t][2][big] - tr_d[t][2][big]) / tr_h[t][2][big]) < 1e-9
        for key in ("matched", "unmatched", "spawned", "orphaned", "promoted"):
            assert tr_h[t][3][key] == tr_d[t][3][key]
    assert np.array_equal(aux_h, aux_d)
    assert np.array_equal(maps_h[4], maps_d[4])                          # landmark ids
    assert np.allclose(maps_h[0], maps_d[0], rtol=1e-6 if dtype == "f32" else 1e-10, atol=1e-9)
    if spawn:
        assert sum(s[3]["spawned"] for s in tr_d) > 0


def test_closed_loop_on_device_and_accuracy():
    """Simulator -> fused kernel -> resample with no host data in the loop; the on-device analysis against NumPy."""
    import torch
    from parakeet_slam_b200 import analysis
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    from parakeet_slam_b200.simulator import BearingSimulator
    M, T = 1 << 16, 30
    scn = make_scenario("c2", num_particles=M, num_landmarks=64, frames=T)
    fs, clk, tw = _clocked_filter(scn, M, "f32")
    sim = BearingSimulator(scn.landmarks, 8, 0.02, 0.3)
    an = analysis.SlamAnalyzer()
    scan = torch.empty((8, 4), dtype=torch.float64, device="cuda")
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        sim.scan(scn.true_poses[t], out=scan)
        fs.measurement_update(scan)
        if t == T - 1:
            acc = analysis.accuracy(fs, scn.true_poses[t])
            pose = fs.pose.cpu().numpy()
        fs.low_variance_resample()
        an.add(scn.true_poses[t], fs.summary(), analysis.accuracy(fs, scn.true_poses[t])["n_eff"])
    s = fs.stats()
    assert s["matched"] / float(M * 8) > 0.95 and s["flags"] == 0
    # accuracy kernel == NumPy on the same particles
    w = pose[:, 3]
    assert acc["sum_w"] == pytest.approx(w.sum(), rel=1e-12)
    assert acc["n_eff"] == pytest.approx(w.sum() ** 2 / (w * w).sum(), rel=1e-10)
    assert acc["max_w"] == w.max()
    tp = scn.true_poses[T - 1]
    assert acc["rms_x"] == pytest.approx(math.sqrt(((pose[:, 0] - tp[0]) ** 2).mean()), rel=1e-10)
    assert acc["rms_y"] == pytest.approx(math.sqrt(((pose[:, 1] - tp[1]) ** 2).mean()), rel=1e-10)
    dth = (pose[:, 2] - tp[2] + math.pi) % (2 * math.pi) - math.pi
    assert acc["rms_heading"] == pytest.approx(math.sqrt((dth ** 2).mean()), rel=1e-9)
    assert 1.0 < acc["n_eff"] <= M
    # the filter tracks: mean error of the estimate over the run stays small (v = 0.2 m/s, 30 frames)
    rep = an.report()
    assert rep["frames"] == T and rep["rms_x"] < 0.05 and rep["rms_y"] < 0.05 and abs(rep["heading"]) < 0.05
    assert rep["avg_x"] == pytest.approx(math.sqrt(an.x_squared))
    # map error: every particle holds all 64 preset landmarks; they moved little from the truth they started at
    rms, cnt = analysis.map_error(fs, scn.landmarks)
    assert np.array_equal(cnt, np.full(64, M)) and np.nanmax(rms) < 1.0
    mean5, _, _, _, ids, _ = fs.export_maps(0, 256)
    want = np.sqrt(((mean5[:, :, :2] - scn.landmarks[None, :, :2]) ** 2).sum(-1).mean(0))
    rms_s, _ = analysis.map_error(fs, scn.landmarks)
    assert rms_s.shape == want.shape


def test_calc_errors_matches_reference_helpers():
    """utils.calc_errors semantics (utils.py:83-205) on plain triples -- host arithmetic, no device needed."""
    from parakeet_slam_b200.analysis import SlamAnalyzer
    along, off, head = SlamAnalyzer.calc_errors((1.0, 1.0, 0.3), (0.0, 0.0, 0.0))
    assert along == pytest.approx(1.0) and off == pytest.approx(1.0) and head == pytest.approx(0.3)
    along, off, head = SlamAnalyzer.calc_errors((0.0, -2.0, 0.0), (0.0, 0.0, math.pi / 2))
    assert along == pytest.approx(-2.0) and abs(off) < 1e-12
    assert SlamAnalyzer.calc_errors((0.0, 0.0, 1.0), (0.0, 0.0, 0.25)) == (0.0, 0.0, 0.75)


def test_checkpoint_round_trip_and_layout_checks():
    """state_dict / load_state_dict: a restored filter continues bit-identically (device state, Philox frame counter,
    last control and time), and a checkpoint of another layout is refused."""
    import random
    import torch
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    scn = make_scenario("c2", num_particles=2048, num_landmarks=16, frames=9)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)

    def build(urng, clk, **kw):
        fs = FastSLAM(make_features(scn), num_particles=2048, dtype="f32", arithmetic="f32", noise="philox", seed=3,
                      uniform=urng.random, clock=clk, **kw)
        tw = messages.Twist()
        tw.linear.x, tw.angular.z = scn.v, 0.05
        fs.last_control = tw
        return fs, tw

    def advance(fs, tw, clk, frames):
        for t in frames:
            clk.ns += DT_NSEC
            fs.motion_update(tw)
            fs.measurement_update(scn.observations[t])
            fs.low_variance_resample()

    clk_a, rng_a = Clk(), random.Random(1)
    a, tw = build(rng_a, clk_a)
    advance(a, tw, clk_a, range(0, 4))
    sd = a.state_dict()
    rng_state, ns = rng_a.getstate(), clk_a.ns
    advance(a, tw, clk_a, range(4, 9))

    clk_b, rng_b = Clk(), random.Random(7)
    b, _ = build(rng_b, clk_b)
    b.last_control = messages.Twist()                       # the checkpoint brings control and time back
    b.load_state_dict(sd)
    rng_b.setstate(rng_state)
    clk_b.ns = ns
    assert b.last_control.angular.z == 0.05 and b.last_update.to_nsec() == a.last_update.to_nsec() - 5 * DT_NSEC
    advance(b, b.last_control, clk_b, range(4, 9))
    assert torch.equal(a.pose, b.pose) and torch.equal(a.aux, b.aux)
    ma, mb = a.export_maps(), b.export_maps()
    assert all(np.array_equal(x, y) for x, y in zip(ma, mb))
    c, _ = build(random.Random(1), Clk(), capacity=32)      # another capacity: another block layout
    with pytest.raises(ValueError):
        c.load_state_dict(sd)
    with pytest.raises(ValueError):
        a.import_maps(2040, *[x[:16] for x in ma[:5]])      # 16 particles from 2040: past the end
