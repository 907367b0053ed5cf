"""The NumPy restatement (oracle/fastslam_np.py) against golden vectors produced by the
unmodified reference (oracle/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

from oracle import fastslam_np as onp

from conftest import TRACE_FIXTURES, load_trace, scenario_from_trace


def _rel(a, b, floor=1e-300):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def test_likelihood_vectors(unit_vectors):
    g = unit_vectors
    n = len(g["like_L"])
    L = np.zeros(n)
    for i in range(n):
        L[i] = onp.match_likelihood(g["like_pose"][i:i + 1], g["like_blob"][i:i + 1],
                                    g["like_mean"][i:i + 1, None], g["like_cov"][i:i + 1, None])[0, 0, 0]
    ref = g["like_L"]
    # the match / no-match decision (exact zero, decided by fp64 underflow) must be identical
    assert np.array_equal(L == 0.0, ref == 0.0)
    nz = (ref > 1e-290)
    assert _rel(L[nz], ref[nz]) < 1e-10
    sub = (ref > 0) & ~nz  # denormal range: few significant bits
    assert np.all(np.abs(L[sub] - ref[sub]) <= 1e-6 * ref[sub] + 1e-320)


def test_position_and_colour_pdf(unit_vectors):
    g = unit_vectors
    pose, blob, mean, cov = g["like_pose"], g["like_blob"], g["like_mean"], g["like_cov"]
    for i in range(len(pose)):
        cb, sb = onp.obs_direction(float(blob[i, 0]))
        fx, fy = mean[i, 0], mean[i, 1]
        t = (fx - pose[i, 0]) * cb + (fy - pose[i, 1]) * sb
        near = (pose[i, 0], pose[i, 1]) if t < 0 else (pose[i, 0] + cb * t, pose[i, 1] + sb * t)
        assert near[0] == pytest.approx(g["like_near"][i, 0], abs=1e-14)
        assert near[1] == pytest.approx(g["like_near"][i, 1], abs=1e-14)
        cp = onp._pdf3_lower(blob[i, 1] - mean[i, 2], blob[i, 2] - mean[i, 3],
                             blob[i, 3] - mean[i, 4], cov[i, 2:, 2:])
        assert cp == pytest.approx(g["like_cp"][i], rel=1e-10, abs=1e-320)
        pse = math.atan2(fy - pose[i, 1], fx - pose[i, 0])
        if abs(pse - blob[i, 0]) > math.pi / 2:
            assert g["like_bp"][i] == 0.0
        else:
            bp = onp._pdf2_lower(near[0] - fx, near[1] - fy, cov[i, 0, 0], cov[i, 1, 0], cov[i, 1, 1])
            assert bp == pytest.approx(g["like_bp"][i], rel=1e-10, abs=1e-320)


def test_ekf_vectors(unit_vectors):
    g = unit_vectors
    m = len(g["ekf_factor"])
    for i in range(m):
        st = onp.OracleState(1, g["ekf_mean"][i:i + 1], capacity=1)
        st.cov[0, 0] = g["ekf_cov"][i]
        st.pose[0, :2] = g["ekf_pose"][i]
        ids = np.array([[1]], dtype=np.int32)
        onp.measurement_update(st, g["ekf_blob"][i:i + 1], ids=ids)
        assert _rel(st.mean[0, 0], g["ekf_mean2"][i], 1e-12) < 1e-11
        assert np.max(np.abs(st.cov[0, 0] - g["ekf_cov2"][i])) < 1e-12
        assert st.weight[0] == pytest.approx(g["ekf_factor"][i], rel=1e-11)
        assert st.count[0, 0] == 2


def test_resample_vectors(unit_vectors):
    g = unit_vectors
    for i in range(len(g["rs_n"])):
        n, c = int(g["rs_n"][i]), int(g["rs_count"][i])
        w = g["rs_weight"][i, :n]
        want = g["rs_anc"][i, :c]
        got = onp.resample_sequential(w, float(g["rs_u01"][i]))
        assert np.array_equal(got, want)
        assert c == n
        assert np.array_equal(onp.resample_searchsorted(w, float(g["rs_u01"][i])), want)


def test_motion_vectors(unit_vectors):
    g = unit_vectors
    for i in range(len(g["mo_in"])):
        v, w, dt = g["mo_ctl"][i]
        out = onp.motion_update(g["mo_in"][i:i + 1], g["mo_noise"][i:i + 1], v, w, dt)[0]
        assert np.max(np.abs(out[:2] - g["mo_out"][i, :2])) < 1e-14
        d = abs(out[2] - g["mo_out"][i, 2])
        assert min(d, abs(d - 2 * math.pi)) < 1e-12


def test_summary_vector(unit_vectors):
    g = unit_vectors
    out = onp.summary(g["su_pose"])
    assert np.allclose(out, g["su_out"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", TRACE_FIXTURES)
def test_trace(name):
    g = load_trace(name)
    scn = scenario_from_trace(g)
    cps = tuple(int(c) for c in g["checkpoints"])
    pot = tuple(int(j) for j in g["potential_slots"]) if "potential_slots" in g.files else ()
    tr = onp.run_scenario(scn, record_landmarks_at=cps, potential_slots=pot)
    assert np.array_equal(tr["assoc"], g["assoc"])           # bit-exact indices
    assert np.array_equal(tr["ancestors"], g["ancestors"])
    assert np.array_equal(tr["next_id"], g["next_id"])
    assert np.max(np.abs(tr["pose_pre"] - g["pose_pre"])) < 1e-12
    assert np.max(np.abs(tr["pose_post"] - g["pose_post"])) < 1e-12
    assert _rel(tr["weight"], g["weight"]) < 1e-9
    assert np.max(np.abs(tr["summary"] - g["summary"])) < 1e-12
    for t in cps:
        assert np.max(np.abs(tr["lm_mean"][t] - g["lm_mean_%d" % t])) < 1e-9
        assert np.max(np.abs(tr["lm_cov"][t][..., :2, :2] - g["lm_covp_%d" % t])) < 1e-10
        assert np.max(np.abs(tr["lm_cov"][t][..., 2:, 2:] - g["lm_covc_%d" % t])) < 1e-10
        assert np.array_equal(tr["lm_count"][t], g["lm_count_%d" % t])
        if pot:
            assert np.array_equal(tr["lm_potential"][t], g["lm_potential_%d" % t])
    if pot:  # the fixture really exercises negative ids and the promotion of :114-118
        assert (g["assoc"] < 0).any() and (g["assoc"][-1] > 0).all()
        assert g["lm_potential_%d" % cps[0]].any() and not g["lm_potential_%d" % cps[-1]][:, list(pot)].all()
    assert float(g["max_cross_block"]) == 0.0


def _sorted_spawn_state(ids, mean, cov, cnt, n_max):
    """Oracle landmark arrays -> the fixture's layout (landmarks sorted by |id|, zero padded)."""
    M = ids.shape[0]
    o_ids = np.zeros((M, n_max), dtype=np.int64)
    o_mean = np.zeros((M, n_max, 5))
    o_cov = np.zeros((M, n_max, 5, 5))
    o_cnt = np.zeros((M, n_max), dtype=np.int64)
    for i in range(M):
        js = [j for j in range(ids.shape[1]) if ids[i, j] != 0]
        js.sort(key=lambda j: abs(int(ids[i, j])))
        assert len(js) <= n_max
        for q, j in enumerate(js):
            o_ids[i, q], o_mean[i, q], o_cov[i, q], o_cnt[i, q] = ids[i, j], mean[i, j], cov[i, j], cnt[i, j]
    return o_ids, o_mean, o_cov, o_cnt


def test_spawn_trace():
    """Spawn mode: the restatement against the reference run with the three documented
    patches -- new potential landmarks, their promotion, the orphaned readings."""
    g = load_trace("trace_corridor_spawn_m24_t40")
    scn = scenario_from_trace(g)
    cps = tuple(int(c) for c in g["checkpoints"])
    n_max = int(g["n_max"])
    tr = onp.run_scenario(scn, record_landmarks_at=cps, spawn=True, known_map=False, capacity=2 * n_max)
    assert np.array_equal(tr["assoc"], g["assoc"])
    assert np.array_equal(tr["ancestors"], g["ancestors"])
    assert np.array_equal(tr["next_id"], g["next_id"])
    assert np.max(np.abs(tr["pose_post"] - g["pose_post"])) < 1e-12
    assert _rel(tr["weight"], g["weight"]) < 1e-9
    for t in cps:
        ids, mean, cov, cnt = _sorted_spawn_state(tr["lm_ids"][t], tr["lm_mean"][t], tr["lm_cov"][t], tr["lm_count"][t], n_max)
        assert np.array_equal(ids, g["sp_ids_%d" % t])
        assert np.max(np.abs(mean - g["sp_mean_%d" % t])) < 1e-9
        assert np.max(np.abs(cov[..., :2, :2] - g["sp_covp_%d" % t])) < 1e-10
        assert np.max(np.abs(cov[..., 2:, 2:] - g["sp_covc_%d" % t])) < 1e-10
        assert np.array_equal(cnt, g["sp_count_%d" % t])
        for i, orph in enumerate(tr["orphans"][t]):
            assert len(orph) == int(g["sp_north_%d" % t][i])
            if orph:
                assert np.max(np.abs(np.asarray(orph) - g["sp_orph_%d" % t][i, :len(orph)])) < 1e-12
    a = g["assoc"]
    assert (a == 0).any() and (a < 0).any() and (a > 0).any()      # orphans, potential and promoted landmarks all occur
