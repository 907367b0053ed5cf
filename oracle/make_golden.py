"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the development container (where ``/root/reference`` exists):

    python -m oracle.make_golden            # all quick fixtures
    python -m oracle.make_golden --full-c1  # + BASELINE config 1 (100 x 20 x 500; ~1 h)

Every array in the fixtures is an output of reference code executed through
``oracle/ref_shim.py`` (fake ROS modules + two Py2->3 textual substitutions); inputs
come from ``parakeet_slam_b200.scenario`` with the fixed seeds of SURVEY.md 8(d).
The fixtures travel to the GPU box; ``/root/reference`` does not.
"""
from __future__ import annotations

import argparse
import math
import os
import random as _pyrandom
import sys
import warnings

import numpy as np

from parakeet_slam_b200.rosless import clock
from parakeet_slam_b200.scenario import make_scenario

from . import ref_driver, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def _blockdiag_parts(cov):
    """[..., 5, 5] -> position block [..., 2, 2], colour block [..., 3, 3], max |cross term|."""
    cross = max(float(np.abs(cov[..., :2, 2:]).max(initial=0.0)),
                float(np.abs(cov[..., 2:, :2]).max(initial=0.0)))
    return cov[..., :2, :2].copy(), cov[..., 2:, 2:].copy(), cross


def trace_fixture(name, scn, frames, checkpoints, spawn=False, known_map=True, potential_slots=()):
    tr = ref_driver.run_reference(scn, frames=frames, record_landmarks_at=checkpoints,
                                  spawn=spawn, known_map=known_map, potential_slots=potential_slots)
    out = dict(
        scenario=np.array([scn.name]), trajectory=np.array([scn.meta["trajectory"]]),
        num_particles=scn.num_particles, num_landmarks=scn.num_landmarks,
        obs_per_frame=scn.obs_per_frame, frames=frames, v=scn.v, w=scn.w, dt=scn.dt,
        immutable=scn.immutable, preset_covar=scn.preset_covar,
        landmarks=scn.landmarks, observations=scn.observations[:frames], u01=scn.u01[:frames],
        motion_seed=scn.motion_seed,
        pose_pre=tr["pose_pre"], pose_post=tr["pose_post"], assoc=tr["assoc"].astype(np.int16),
        weight=tr["weight"], ancestors=tr["ancestors"].astype(np.int32), summary=tr["summary"],
        next_id=tr["next_id"].astype(np.int32), checkpoints=np.asarray(checkpoints),
        potential_slots=np.asarray(potential_slots, dtype=np.int32),
    )
    worst_cross = 0.0
    for t in checkpoints:
        cp, cc, cross = _blockdiag_parts(tr["lm_cov"][t])
        worst_cross = max(worst_cross, cross)
        out["lm_mean_%d" % t] = tr["lm_mean"][t]
        out["lm_covp_%d" % t] = cp
        out["lm_covc_%d" % t] = cc
        out["lm_count_%d" % t] = tr["lm_count"][t].astype(np.int32)
        out["lm_potential_%d" % t] = tr["lm_potential"][t]
    out["max_cross_block"] = worst_cross
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB), unmatched frac %.3f, cross-block max %.1e" % (
        path, os.path.getsize(path) / 1024.0, float((tr["assoc"] == 0).mean()), worst_cross))


def spawn_arrays(spawn_state, n_max, o_max):
    """Pack ``ref_driver.spawn_state`` into arrays: landmarks sorted by |id| (signed ids, 0 padded),
    orphaned readings in insertion order."""
    M = len(spawn_state)
    ids = np.zeros((M, n_max), dtype=np.int32)
    mean = np.zeros((M, n_max, 5))
    cov = np.zeros((M, n_max, 5, 5))
    cnt = np.zeros((M, n_max), dtype=np.int32)
    north = np.zeros(M, dtype=np.int32)
    orph = np.zeros((M, o_max, 7))
    for i, ps in enumerate(spawn_state):
        for j, id_ in enumerate(sorted(ps["landmarks"], key=abs)):
            m, c, n = ps["landmarks"][id_]
            ids[i, j], mean[i, j], cov[i, j], cnt[i, j] = id_, m, c, n
        north[i] = len(ps["orphans"])
        for j, o in enumerate(ps["orphans"]):
            orph[i, j] = o
    return ids, mean, cov, cnt, north, orph


def spawn_fixture(name, scn, frames, checkpoints):
    """Unknown-map run of the reference with the three spawn patches of SURVEY.md A.6
    (``ref_shim.apply_spawn_patches``): new potential landmarks, promotion, orphaned readings."""
    tr = ref_driver.run_reference(scn, frames=frames, record_landmarks_at=checkpoints, spawn=True, known_map=False)
    n_max = max(len(ps["landmarks"]) for t in checkpoints for ps in tr["spawn_state"][t])
    o_max = max(len(ps["orphans"]) for t in checkpoints for ps in tr["spawn_state"][t])
    out = dict(
        scenario=np.array([scn.name]), trajectory=np.array([scn.meta["trajectory"]]),
        num_particles=scn.num_particles, num_landmarks=scn.num_landmarks,
        obs_per_frame=scn.obs_per_frame, frames=frames, v=scn.v, w=scn.w, dt=scn.dt,
        immutable=scn.immutable, preset_covar=scn.preset_covar,
        landmarks=scn.landmarks, observations=scn.observations[:frames], u01=scn.u01[:frames],
        motion_seed=scn.motion_seed,
        pose_pre=tr["pose_pre"], pose_post=tr["pose_post"], assoc=tr["assoc"].astype(np.int16),
        weight=tr["weight"], ancestors=tr["ancestors"].astype(np.int32), summary=tr["summary"],
        next_id=tr["next_id"].astype(np.int32), checkpoints=np.asarray(checkpoints),
        pair_gate=300.0 ** 0.5, n_max=n_max, o_max=o_max,
    )
    worst_cross = 0.0
    for t in checkpoints:
        ids, mean, cov, cnt, north, orph = spawn_arrays(tr["spawn_state"][t], n_max, max(o_max, 1))
        cp, cc, cross = _blockdiag_parts(cov)
        worst_cross = max(worst_cross, cross)
        out["sp_ids_%d" % t] = ids
        out["sp_mean_%d" % t] = mean
        out["sp_covp_%d" % t] = cp
        out["sp_covc_%d" % t] = cc
        out["sp_count_%d" % t] = cnt
        out["sp_north_%d" % t] = north
        out["sp_orph_%d" % t] = orph
    out["max_cross_block"] = worst_cross
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    a = tr["assoc"]
    print("wrote %s (%.1f KB): unmatched %.3f, potential %.3f, full %.3f of the pairs; <= %d landmarks, <= %d orphans "
          "per particle; cross-block max %.1e" % (path, os.path.getsize(path) / 1024.0, float((a == 0).mean()),
                                                   float((a < 0).mean()), float((a > 0).mean()), n_max, o_max, worst_cross))


def unit_fixture(ref):
    """Known-answer vectors from direct calls of the reference's scalar methods."""
    core, msgs = ref.core, ref.msgs
    rs = np.random.RandomState(99)
    particle = core.FilterParticle()

    # --- probability_of_match / prob_position_match / prob_color_match / closest_point -------
    n = 400
    pose = np.zeros((n, 3))
    blob = np.zeros((n, 4))
    mean = np.zeros((n, 5))
    cov = np.zeros((n, 5, 5))
    L = np.zeros(n)
    bp = np.zeros(n)
    cpv = np.zeros(n)
    near = np.zeros((n, 2))
    for i in range(n):
        x, y = rs.uniform(-3, 3, 2)
        th = rs.uniform(-3.1, 3.1) if i % 3 else rs.uniform(-0.2, 0.2)
        fx, fy = rs.uniform(-8, 8, 2)
        col = rs.uniform(0, 255, 3)
        A = rs.normal(size=(2, 2)) * rs.choice([0.05, 0.3, 1.0])
        Sp = A @ A.T + np.identity(2) * rs.choice([1e-3, 0.05, 0.25])
        B = rs.normal(size=(3, 3)) * rs.choice([0.05, 0.5, 3.0])
        Sc = B @ B.T + np.identity(3) * rs.choice([1e-2, 0.25, 5.0])
        # slight asymmetry, as (I-KH)Sigma produces: only the lower triangle must be used
        Sp[0, 1] += 1e-3 * rs.normal()
        Sc[0, 2] += 1e-3 * rs.normal()
        S = np.zeros((5, 5))
        S[:2, :2] = Sp
        S[2:, 2:] = Sc
        true_b = math.atan2(fy - y, fx - x) - th
        kind = i % 8
        if kind == 0:       # near the bearing gate (0.5 rad)
            b = true_b + rs.choice([-1, 1]) * (0.5 + rs.normal() * 1e-3)
        elif kind == 1:     # far off in bearing
            b = true_b + rs.uniform(0.6, 3.0)
        else:
            b = true_b + rs.normal() * 0.05
        if kind == 2:       # near the colour gate (300)
            d = rs.normal(size=3)
            d *= math.sqrt(300.0 + rs.normal() * 0.5) / np.linalg.norm(d)
            bc = col + d
        elif kind == 3:     # far colour
            bc = col + rs.uniform(20, 60, 3)
        elif kind == 4:     # deep tail: likelihood underflows toward 0
            bc = col + rs.normal(size=3) * 9.0
        else:
            bc = col + rs.normal(size=3) * 0.4
        st = msgs.Odometry()
        st.pose.pose.position.x = float(x)
        st.pose.pose.position.y = float(y)
        st.pose.pose.orientation = ref.utils.heading_to_quaternion(float(th))
        bl = msgs.Blob()
        bl.bearing = float(b)
        bl.color.r, bl.color.g, bl.color.b = float(bc[0]), float(bc[1]), float(bc[2])
        f = core.Feature(mean=np.array([fx, fy, col[0], col[1], col[2]]), covar=S)
        pose[i] = (x, y, ref.utils.quaternion_to_heading(st.pose.pose.orientation))
        blob[i] = (b, bc[0], bc[1], bc[2])
        mean[i] = f.mean
        cov[i] = S
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            L[i] = particle.probability_of_match(st, bl, f)
            bp[i] = particle.prob_position_match(f.mean, f.covar, float(x), float(y), float(b))
            cpv[i] = particle.prob_color_match(f.mean, f.covar, bl)
        near[i] = particle.closest_point(float(fx), float(fy), float(x), float(y), float(b))

    # --- EKF pieces: jacobian, Q, K, mean/cov update, importance factor ---------------------
    m = 200
    e_pose = np.zeros((m, 2))
    e_mean = np.zeros((m, 5))
    e_cov = np.zeros((m, 5, 5))
    e_blob = np.zeros((m, 4))
    e_H = np.zeros((m, 4, 5))
    e_Q = np.zeros((m, 4, 4))
    e_K = np.zeros((m, 5, 4))
    e_zhat = np.zeros((m, 4))
    e_mean2 = np.zeros((m, 5))
    e_cov2 = np.zeros((m, 5, 5))
    e_factor = np.zeros(m)
    Qt = core.FastSLAM.__new__(core.FastSLAM)
    Qt = ref.matrix.Matrix(np.identity(4) * 0.1)
    for i in range(m):
        p = core.FilterParticle()
        x, y = rs.uniform(-3, 3, 2)
        p.state.pose.pose.position.x = float(x)
        p.state.pose.pose.position.y = float(y)
        fx, fy = (x, y) if i == 0 else rs.uniform(-8, 8, 2)   # i == 0: q == 0 Jacobian branch
        col = rs.uniform(0, 255, 3)
        A = rs.normal(size=(2, 2)) * 0.4
        B = rs.normal(size=(3, 3)) * 0.4
        S = np.zeros((5, 5))
        S[:2, :2] = A @ A.T + np.identity(2) * 0.05
        S[2:, 2:] = B @ B.T + np.identity(3) * 0.05
        S[0, 1] += 1e-4 * rs.normal()
        S[3, 2] += 1e-4 * rs.normal()
        f = core.Feature(mean=np.array([fx, fy, col[0], col[1], col[2]]), covar=S.copy())
        p.feature_set[1] = f
        bl = msgs.Blob()
        bl.bearing = float(math.atan2(fy - y, fx - x) + rs.normal() * 0.1)
        bc = col + rs.normal(size=3) * 0.5
        bl.color.r, bl.color.g, bl.color.b = float(bc[0]), float(bc[1]), float(bc[2])
        pseudo = p.generate_measurement(1)
        H = p.measurement_jacobian(1)
        Q = p.measurement_covariance(H, 1, Qt)
        Qinv = ref.matrix.inverse(Q)
        Kg = p.kalman_gain(1, H, Qinv)
        e_pose[i] = (x, y)
        e_mean[i] = f.mean
        e_cov[i] = S
        e_blob[i] = (bl.bearing, bc[0], bc[1], bc[2])
        e_H[i], e_Q[i], e_K[i] = H, Q, Kg
        e_zhat[i] = (pseudo.bearing, pseudo.color.r, pseudo.color.g, pseudo.color.b)
        e_factor[i] = p.importance_factor(Q, bl, pseudo)
        f.update_mean(Kg, bl, pseudo)
        f.update_covar(Kg, H)
        e_mean2[i] = f.mean
        e_cov2[i] = f.covar
        assert f.update_count == 2

    # --- low_variance_resample on hand-made weight vectors -----------------------------------
    cases = []
    for M, kind in ((7, "uniform"), (64, "random"), (64, "peaked"), (33, "zeros_some"),
                    (16, "all_zero"), (50, "tiny"), (128, "one_hot"), (257, "random")):
        if kind == "uniform":
            w = np.full(M, 0.25)
        elif kind == "random":
            w = rs.uniform(0, 1, M)
        elif kind == "peaked":
            w = np.exp(-0.5 * ((np.arange(M) - M / 3.0) / 1.5) ** 2)
        elif kind == "zeros_some":
            w = rs.uniform(0, 1, M) * (rs.uniform(size=M) > 0.5)
        elif kind == "all_zero":
            w = np.zeros(M)
        elif kind == "tiny":
            w = rs.uniform(0, 1, M) * 1e-300
        elif kind == "one_hot":
            w = np.zeros(M)
            w[M // 2 + 3] = 3e-7
        for rep in range(3):
            seed = 1000 + len(cases)
            fs = core.FastSLAM.__new__(core.FastSLAM)
            fs.particles = [core.FilterParticle() for _ in range(M)]
            for idx, (pp, ww) in enumerate(zip(fs.particles, w)):
                pp.weight = float(ww)
                pp._trace_index = idx
            fs.aged_particles_pub = fs.resampled_particles_pub = ref.rospy.Publisher("x")
            _pyrandom.seed(seed)
            u01 = _pyrandom.random()
            _pyrandom.seed(seed)
            fs.low_variance_resample()
            anc = np.array([pp._trace_index for pp in fs.particles], dtype=np.int32)
            cases.append((w.copy(), u01, anc))
    rw = np.full((len(cases), 257), np.nan)
    ra = np.full((len(cases), 257), -1, dtype=np.int32)
    ru = np.zeros(len(cases))
    rn = np.zeros(len(cases), dtype=np.int32)
    rc = np.zeros(len(cases), dtype=np.int32)
    for i, (w, u01, anc) in enumerate(cases):
        rw[i, :len(w)] = w
        ra[i, :len(anc)] = anc
        ru[i], rn[i], rc[i] = u01, len(w), len(anc)

    # --- motion_model on single particles ------------------------------------------------------
    k = 120
    mo_in = np.zeros((k, 3))
    mo_ctl = np.zeros((k, 3))       # v, w, dt
    mo_noise = np.zeros((k, 3))
    mo_out = np.zeros((k, 3))
    fs = core.FastSLAM.__new__(core.FastSLAM)
    for i in range(k):
        p = core.FilterParticle()
        x, y = rs.uniform(-5, 5, 2)
        th = rs.uniform(-3.14, 3.14) if i % 4 else rs.choice([-math.pi, math.pi, 3.1415, -3.1415])
        p.state.pose.pose.position.x = float(x)
        p.state.pose.pose.position.y = float(y)
        p.state.pose.pose.orientation = ref.utils.heading_to_quaternion(float(th))
        th_read = ref.utils.quaternion_to_heading(p.state.pose.pose.orientation)
        tw = msgs.Twist()
        tw.linear.x = float(rs.choice([0.0, 0.2, -0.3, 1.0]))
        tw.angular.z = float(rs.choice([0.0, 0.1, -0.4, 2.0]))
        dt_ns = int(rs.choice([0, 90909091, 100000000, 1234567890]))
        np.random.seed(5000 + i)
        z = np.random.standard_normal(3)
        np.random.seed(5000 + i)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            q = fs.motion_model(p, tw, ref.rospy.Duration(0, dt_ns))
        mo_in[i] = (x, y, th_read)
        mo_ctl[i] = (tw.linear.x, tw.angular.z, ref.rospy.Duration(0, dt_ns).to_sec())
        mo_noise[i] = z
        mo_out[i] = ref_driver.pose_of(ref, q)

    # --- summary -------------------------------------------------------------------------------
    fs = core.FastSLAM.__new__(core.FastSLAM)
    fs.particles = [core.FilterParticle() for _ in range(37)]
    su_pose = np.zeros((37, 3))
    for i, p in enumerate(fs.particles):
        x, y, th = rs.uniform(-4, 4), rs.uniform(-4, 4), rs.uniform(2.6, 3.6)
        p.state.pose.pose.position.x = float(x)
        p.state.pose.pose.position.y = float(y)
        p.state.pose.pose.orientation = ref.utils.heading_to_quaternion(float(th))
        su_pose[i] = ref_driver.pose_of(ref, p)
    su_out = np.array([float(v) for v in fs.summary()])

    path = os.path.join(GOLDEN_DIR, "unit_vectors.npz")
    np.savez_compressed(
        path, like_pose=pose, like_blob=blob, like_mean=mean, like_cov=cov, like_L=L,
        like_bp=bp, like_cp=cpv, like_near=near,
        ekf_pose=e_pose, ekf_mean=e_mean, ekf_cov=e_cov, ekf_blob=e_blob, ekf_H=e_H, ekf_Q=e_Q,
        ekf_K=e_K, ekf_zhat=e_zhat, ekf_mean2=e_mean2, ekf_cov2=e_cov2, ekf_factor=e_factor,
        rs_weight=rw, rs_u01=ru, rs_anc=ra, rs_n=rn, rs_count=rc,
        mo_in=mo_in, mo_ctl=mo_ctl, mo_noise=mo_noise, mo_out=mo_out,
        su_pose=su_pose, su_out=su_out)
    print("wrote %s (%.1f KB); L>0: %d/%d, L==0: %d" % (
        path, os.path.getsize(path) / 1024.0, int((L > 0).sum()), n, int((L == 0).sum())))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-c1", action="store_true", help="also BASELINE config 1 (slow)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args(argv)
    if not ref_shim.available():
        print("reference sources not present; nothing to do", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jobs = {
        "unit": lambda: unit_fixture(ref_shim.load_reference(with_ros_node=False)),
        "circle": lambda: trace_fixture(
            "trace_circle_m32_t60", make_scenario("c1", num_particles=32, frames=60), 60,
            (0, 29, 59)),
        "corridor": lambda: trace_fixture(
            "trace_corridor_m32_t60",
            make_scenario("c1", num_particles=32, frames=60, trajectory="corridor"), 60,
            (0, 29, 59)),
        "immutable": lambda: trace_fixture(
            "trace_circle_immutable_m32_t40",
            make_scenario("c1", num_particles=32, frames=40, immutable=True), 40, (0, 39)),
        "noisy": lambda: trace_fixture(
            "trace_corridor_noisy_m48_t40",
            make_scenario("c1", num_particles=48, frames=40, trajectory="corridor",
                          sigma_color=3.0, sigma_bearing=0.08, num_landmarks=40), 40, (0, 39)),
    }
    jobs["potential"] = lambda: trace_fixture(
        "trace_corridor_potential_m24_t12",
        make_scenario("c1", num_particles=24, frames=12, trajectory="corridor", num_landmarks=16), 12,
        (0, 2, 3, 11), potential_slots=(1, 5, 8, 9, 12, 13, 14, 15))
    jobs["spawn"] = lambda: spawn_fixture(
        "trace_corridor_spawn_m24_t40",
        make_scenario("c3", num_particles=24, num_landmarks=24, frames=40, obs_per_frame=8), 40,
        (0, 1, 2, 5, 10, 20, 39))
    if args.full_c1:
        jobs["c1"] = lambda: trace_fixture(
            "trace_c1_m100_n20_t500", make_scenario("c1"), 500, (0, 99, 249, 499))
    for key, job in jobs.items():
        if args.only and key not in args.only.split(","):
            continue
        job()
    return 0


if __name__ == "__main__":
    sys.exit(main())
