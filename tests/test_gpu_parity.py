"""Parity of the CUDA path (through the C ABI) against the golden vectors produced by the
unmodified reference and against the NumPy oracle.  Needs a B200."""
import ctypes
import math

import numpy as np
import pytest

from conftest import TRACE_FIXTURES, load_trace, scenario_from_trace

pytestmark = pytest.mark.gpu


def _rel(a, b, floor=1e-300):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


@pytest.fixture(scope="module")
def lib():
    from parakeet_slam_b200 import _lib
    _lib.require_device()
    return _lib.load()


def test_probe_likelihood_golden(lib, unit_vectors):
    from parakeet_slam_b200 import probe
    g = unit_vectors
    cov = g["like_cov"]
    L = probe.likelihood_batch(g["like_pose"], g["like_blob"], g["like_mean"], cov[:, :2, :2], cov[:, 2:, 2:])
    ref = g["like_L"]
    # match / no-match (exact zero via gates or fp64 underflow) must be identical
    assert np.array_equal(L == 0.0, ref == 0.0)
    nz = ref > 1e-290
    assert _rel(L[nz], ref[nz]) < 1e-9            # fp64 device vs fp64 reference
    sub = (ref > 0) & ~nz
    assert np.all(np.abs(L[sub] - ref[sub]) <= 1e-6 * ref[sub] + 1e-320)


def test_probe_ekf_golden(lib, unit_vectors):
    from parakeet_slam_b200 import probe
    g = unit_vectors
    cov = g["ekf_cov"]
    mean2, covp2, covc2, factor = probe.ekf_batch(g["ekf_pose"], g["ekf_blob"], g["ekf_mean"], cov[:, :2, :2],
                                                  cov[:, 2:, 2:])
    assert np.max(np.abs(mean2 - g["ekf_mean2"])) < 1e-10      # colours are O(255): 4e-13 relative
    assert np.max(np.abs(covp2 - g["ekf_cov2"][:, :2, :2])) < 1e-12
    assert np.max(np.abs(covc2 - g["ekf_cov2"][:, 2:, 2:])) < 1e-12
    assert _rel(factor, g["ekf_factor"]) < 1e-10


def test_motion_golden(lib, unit_vectors):
    import torch
    from parakeet_slam_b200 import _lib
    g = unit_vectors
    n = len(g["mo_in"])
    exact_xy = 0
    for i in range(n):
        rec = torch.tensor([[g["mo_in"][i, 0], g["mo_in"][i, 1], g["mo_in"][i, 2], 1.0]], dtype=torch.float64,
                           device="cuda")
        z = torch.from_numpy(g["mo_noise"][i:i + 1].copy()).cuda()
        v, w, dt = (float(x) for x in g["mo_ctl"][i])
        _lib.check(lib.pk_motion_update(_lib.ptr(rec), 1, _lib.ptr(z), 0, 0, 0, v, w, dt, None))
        out = rec.cpu().numpy()[0]
        assert np.max(np.abs(out[:2] - g["mo_out"][i, :2])) < 1e-14
        exact_xy += int(np.array_equal(out[:2], g["mo_out"][i, :2]))
        d = abs(out[2] - g["mo_out"][i, 2])
        assert min(d, abs(d - 2 * math.pi)) < 1e-12
        assert out[3] == 1.0
    assert exact_xy >= 0.5 * n   # most positions reproduce bit-for-bit (device sincos vs libm differ by <=1 ulp)


def test_resample_golden(lib, unit_vectors):
    import torch
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import clock
    g = unit_vectors
    clock.set(0.0)
    for i in range(len(g["rs_n"])):
        n = int(g["rs_n"][i])
        w = g["rs_weight"][i, :n]
        fs = FastSLAM([], num_particles=n, uniform=lambda u=float(g["rs_u01"][i]): u)
        fs.keep_trace = True
        fs.pose[:, 3] = torch.from_numpy(w.copy()).cuda()
        fs.pose[:, 0] = torch.arange(n, dtype=torch.float64, device="cuda")   # tag particles by x
        fs.low_variance_resample()
        anc = fs.last_ancestors.cpu().numpy()
        assert np.array_equal(anc, g["rs_anc"][i, :n]), "case %d" % i
        assert np.array_equal(fs.pose[:, 0].cpu().numpy(), anc.astype(np.float64))


def _check_trace(tr, g, cps, exact, tol_state, tol_weight, min_index_match):
    assoc_match = float((tr["assoc"] == g["assoc"]).mean())
    anc_match = float((tr["ancestors"] == g["ancestors"]).mean())
    if exact:
        assert np.array_equal(tr["assoc"], g["assoc"])
        assert np.array_equal(tr["ancestors"], g["ancestors"])
        assert np.array_equal(tr["next_id"], g["next_id"])
    assert assoc_match >= min_index_match and anc_match >= min_index_match, (assoc_match, anc_match)
    if exact:
        # pose: abs 1e-9 (x, y in metres, heading in rad) as SURVEY 8(d) states
        assert np.max(np.abs(tr["pose_pre"] - g["pose_pre"])) < 1e-9
        assert np.max(np.abs(tr["pose_post"] - g["pose_post"])) < 1e-9
        assert np.max(np.abs(tr["summary"] - g["summary"])) < 1e-9
        big = g["weight"] > 1e-300
        assert _rel(tr["weight"][big], g["weight"][big]) < tol_weight
        for t in cps:
            assert _rel(tr["lm_mean"][t], g["lm_mean_%d" % t], 1e-3) < tol_state
            assert np.max(np.abs(tr["lm_covp"][t] - g["lm_covp_%d" % t])) < tol_state
            assert np.max(np.abs(tr["lm_covc"][t] - g["lm_covc_%d" % t])) < tol_state
            assert np.array_equal(tr["lm_count"][t], g["lm_count_%d" % t])
    return assoc_match, anc_match


@pytest.mark.parametrize("name", TRACE_FIXTURES)
def test_trace_f64(lib, name):
    """fp64 storage: indices bit-exact, state within 1e-5 relative (measured ~1e-12)."""
    from device_harness import run_device
    g = load_trace(name)
    scn = scenario_from_trace(g)
    cps = tuple(int(c) for c in g["checkpoints"])
    pot = tuple(int(j) for j in g["potential_slots"]) if "potential_slots" in g.files else ()
    tr = run_device(scn, "f64", checkpoints=cps, potential_slots=pot)
    _check_trace(tr, g, cps, exact=True, tol_state=1e-5, tol_weight=1e-5, min_index_match=1.0)
    for t in cps:
        if pot:   # potential -> full promotion (:114-118) happened on the same frames as in the reference
            assert np.array_equal(tr["lm_potential"][t], g["lm_potential_%d" % t])
    # tighter: what the fp64 path actually achieves
    assert np.max(np.abs(tr["pose_pre"] - g["pose_pre"])) < 1e-11
    big = g["weight"] > 1e-300
    assert _rel(tr["weight"][big], g["weight"][big]) < 1e-8


@pytest.mark.parametrize("name", TRACE_FIXTURES)
def test_trace_f32_storage(lib, name):
    """fp32 landmark storage: >= 90 % of association / resampling indices reproduced exactly
    (BASELINE.json target); pose never touches fp32 so the motion path stays exact until the
    first differing resample."""
    from device_harness import run_device
    g = load_trace(name)
    scn = scenario_from_trace(g)
    pot = tuple(int(j) for j in g["potential_slots"]) if "potential_slots" in g.files else ()
    tr = run_device(scn, "f32", potential_slots=pot)
    a, r = _check_trace(tr, g, (), exact=False, tol_state=1e-5, tol_weight=1e-3, min_index_match=0.90)
    # first frame sees identical state: landmark presets are exactly representable or rounded once
    assert np.array_equal(tr["assoc"][0], g["assoc"][0])
    big = g["weight"][0] > 1e-300
    assert _rel(tr["weight"][0][big], g["weight"][0][big]) < 1e-3


# ---- spawn mode (reference prkt_core_v2.py:546-746 with the three documented patches) --------
def _device_spawn_state(tr, t, n_max):
    """Device landmark arrays at checkpoint t -> the fixture's layout (sorted by |id|, zero padded)."""
    ids, mean = tr["lm_ids"][t], tr["lm_mean"][t]
    covp, covc, cnt = tr["lm_covp"][t], tr["lm_covc"][t], tr["lm_count"][t]
    M = ids.shape[0]
    o = dict(ids=np.zeros((M, n_max), dtype=np.int64), mean=np.zeros((M, n_max, 5)), covp=np.zeros((M, n_max, 2, 2)),
             covc=np.zeros((M, n_max, 3, 3)), cnt=np.zeros((M, n_max), dtype=np.int64))
    for i in range(M):
        js = [j for j in range(ids.shape[1]) if ids[i, j] != 0]
        js.sort(key=lambda j: abs(int(ids[i, j])))
        assert len(js) <= n_max, (len(js), n_max)
        for q, j in enumerate(js):
            o["ids"][i, q], o["mean"][i, q], o["covp"][i, q] = ids[i, j], mean[i, j], covp[i, j]
            o["covc"][i, q], o["cnt"][i, q] = covc[i, j], cnt[i, j]
    return o


def _check_orphans(dev_list, want_rows, want_n):
    """Device readings (x, y, cos, sin, r, g, b, id) against reference readings (id, x, y, angle, r, g, b)."""
    for i, rows in enumerate(dev_list):
        n = int(want_n[i])
        assert len(rows) == n, (i, len(rows), n)
        if n == 0:
            continue
        w = np.asarray(want_rows[i][:n])
        assert np.array_equal(rows[:, 7].astype(np.int64), w[:, 0].astype(np.int64))        # ids, insertion order
        assert np.max(np.abs(rows[:, 0:2] - w[:, 1:3])) < 1e-9                               # pose copy (P2)
        assert np.max(np.abs(rows[:, 2] - np.cos(w[:, 3]))) < 1e-12
        assert np.max(np.abs(rows[:, 3] - np.sin(w[:, 3]))) < 1e-12
        assert np.array_equal(rows[:, 4:7], w[:, 4:7])                                       # blob colours, verbatim


def test_spawn_trace_f64(lib):
    """Unknown map: orphaned readings, triangulated potential landmarks (id < 0) and their promotion must follow
    the patched reference -- ids, ancestors and next_id bit-exact, state within 1e-5 relative."""
    from device_harness import run_device
    g = load_trace("trace_corridor_spawn_m24_t40")
    scn = scenario_from_trace(g)
    cps = tuple(int(c) for c in g["checkpoints"])
    n_max = int(g["n_max"])
    tr = run_device(scn, "f64", checkpoints=cps, spawn=True, known_map=False, capacity=2 * n_max, orphan_capacity=32)
    assert np.array_equal(tr["assoc"], g["assoc"])
    assert np.array_equal(tr["ancestors"], g["ancestors"])
    assert np.array_equal(tr["next_id"], g["next_id"])
    assert np.max(np.abs(tr["pose_post"] - g["pose_post"])) < 1e-9
    big = g["weight"] > 1e-300
    assert _rel(tr["weight"][big], g["weight"][big]) < 1e-5
    for t in cps:
        d = _device_spawn_state(tr, t, n_max)
        assert np.array_equal(d["ids"], g["sp_ids_%d" % t])
        assert _rel(d["mean"], g["sp_mean_%d" % t], 1e-3) < 1e-5
        assert np.max(np.abs(d["covp"] - g["sp_covp_%d" % t])) < 1e-5
        assert np.max(np.abs(d["covc"] - g["sp_covc_%d" % t])) < 1e-5
        assert np.array_equal(d["cnt"], g["sp_count_%d" % t])
        _check_orphans(tr["orphans"][t], g["sp_orph_%d" % t], g["sp_north_%d" % t])
    flags = 0
    for s in tr["stats"]:
        flags |= s["flags"]
    assert flags == 0, flags                       # no expiry, no full map, no degenerate pair in this run
    assert sum(s["spawned"] for s in tr["stats"]) > 0 and sum(s["orphaned"] for s in tr["stats"]) > 0
    assert sum(s["promoted"] for s in tr["stats"]) > 0


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_spawn_against_oracle(lib, dtype):
    """Same run, more particles and landmarks, against the NumPy restatement (itself pinned to the patched
    reference by tests/test_oracle_golden.py::test_spawn_trace)."""
    from device_harness import run_device
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c3", num_particles=192, num_landmarks=48, frames=30, obs_per_frame=8)
    cps = (0, 3, 29)
    to = onp.run_scenario(scn, record_landmarks_at=cps, spawn=True, known_map=False, capacity=64)
    tr = run_device(scn, dtype, checkpoints=cps, spawn=True, known_map=False, capacity=64, orphan_capacity=64)
    if dtype == "f64":
        assert np.array_equal(tr["assoc"], to["assoc"])
        assert np.array_equal(tr["ancestors"], to["ancestors"])
        assert np.array_equal(tr["next_id"], to["next_id"])
        for t in cps:
            ids_o = np.where(to["lm_ids"][t] != 0, to["lm_ids"][t], 0)
            assert np.array_equal(tr["lm_ids"][t], ids_o)          # same slots: both append in creation order
            live = ids_o != 0
            assert _rel(tr["lm_mean"][t][live], to["lm_mean"][t][live], 1e-3) < 1e-5
            assert np.array_equal(tr["lm_count"][t][live], to["lm_count"][t][live])
            for i in range(scn.num_particles):
                assert len(tr["orphans"][t][i]) == len(to["orphans"][t][i])
    else:
        # fp32 storage rounds the freshly triangulated (ill-conditioned, short-baseline) landmarks, so the weights --
        # and with them the resampling -- drift from the fp64 run after the first frames: associations stay >= 90 %
        # identical, ancestors are identical while the maps are, then follow their own (valid) lineage
        # (measured: 72 % of all ancestors over 30 frames).
        assert float((tr["assoc"] == to["assoc"]).mean()) >= 0.90
        assert np.array_equal(tr["ancestors"][:2], to["ancestors"][:2])
        assert float((tr["ancestors"] == to["ancestors"]).mean()) >= 0.50
        assert np.array_equal(tr["assoc"][0], to["assoc"][0]) and np.array_equal(tr["assoc"][1], to["assoc"][1])


def test_spawn_ring_and_capacity_flags(lib):
    """A 4-slot orphan ring expires readings (flagged), a 2-landmark capacity fills up (flagged); nothing overruns."""
    from device_harness import run_device
    from parakeet_slam_b200 import _lib
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c3", num_particles=64, num_landmarks=32, frames=12, obs_per_frame=8)

    def flags_of(tr):
        f = 0
        for s in tr["stats"]:
            f |= s["flags"]
        return f

    # (a) ring of 4 readings with 8 blobs per frame: every frame overwrites the previous one's readings
    tr = run_device(scn, "f64", checkpoints=(11,), spawn=True, known_map=False, capacity=8, orphan_capacity=4)
    assert flags_of(tr) & _lib.PK_FLAG_ORPHAN_EXPIRED
    rows, totals = tr["filter"].export_orphans()
    assert all(len(r) <= 4 for r in rows) and int(totals.max()) > 4
    # (b) room for two landmarks only: further pairs are dropped and flagged, n_live never exceeds the capacity
    tr = run_device(scn, "f64", checkpoints=(11,), spawn=True, known_map=False, capacity=2, orphan_capacity=32)
    assert flags_of(tr) & _lib.PK_FLAG_MAP_FULL
    fs = tr["filter"]
    assert int(fs.aux[:, 0].max()) == 2
    p = fs.particles[0]                                # host view: potential / full landmarks and hypothesis_set
    assert len(p.feature_set) + len(p.potential_features) <= 2 and 0 < len(p.hypothesis_set) <= 32


# ---- fp32 landmark algebra (PK_DTYPE_ARITH_F32): the throughput instantiation --------------------------------
@pytest.mark.parametrize("name", TRACE_FIXTURES)
def test_trace_f32_arithmetic(lib, name):
    """fp32 storage AND fp32 landmark algebra (poses, weights, resampling stay fp64): BASELINE.json asks for >= 90 % of
    the reference's association / resampling indices; what this build delivers -- and what DESIGN.md / the bench line
    claim -- is EVERY index of EVERY reference fixture (including the 500-frame config 1), asserted here per fixture.
    Weights of frame 0 to 1e-3."""
    from device_harness import run_device
    g = load_trace(name)
    scn = scenario_from_trace(g)
    pot = tuple(int(j) for j in g["potential_slots"]) if "potential_slots" in g.files else ()
    tr = run_device(scn, "f32", potential_slots=pot, arithmetic="f32")
    a, r = _check_trace(tr, g, (), exact=False, tol_state=1e-5, tol_weight=1e-3, min_index_match=0.90)
    assert np.array_equal(tr["assoc"][0], g["assoc"][0])
    assert np.array_equal(tr["ancestors"][0], g["ancestors"][0])
    big = g["weight"][0] > 1e-300
    assert _rel(tr["weight"][0][big], g["weight"][0][big]) < 1e-3
    print("fp32 arithmetic, %s: assoc %.4f ancestors %.4f identical" % (name, a, r))
    assert a == 1.0 and r == 1.0, "fp32 algebra: %s: %.6f of the association ids, %.6f of the ancestors identical" % (name, a, r)


def test_f32_arithmetic_state_against_oracle(lib):
    """One frame from identical state, 4096 particles: what fp32 algebra does to a single update -- associations
    identical, weights to 1e-4 relative, landmark means to 1e-5 relative (BASELINE tolerance), covariances to 1e-5
    absolute; then 10 frames: indices >= 90 % identical, pose estimate unchanged to 1e-3 m."""
    import random
    import torch
    from device_harness import make_features
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    M, N, T = 4096, 64, 10
    scn = make_scenario("c2", num_particles=M, num_landmarks=N, frames=T)
    rs = np.random.RandomState(5)
    blocks = [rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    urng = random.Random(8)
    fs = FastSLAM(make_features(scn), num_particles=M, dtype="f32", arithmetic="f32", noise=lambda m: next(it),
                  uniform=urng.random, clock=clk)
    fs.keep_trace = True
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng2 = random.Random(8)
    match = []
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        fs.measurement_update(scn.observations[t])
        if t == 0:
            mean5, covp, covc, meta, ids_, nlive = fs.export_maps()
            w0 = fs.pose[:, 3].cpu().numpy()
        fs.low_variance_resample()
        if t == 0:
            st0 = st.copy()
            st0.pose = onp.motion_update(st0.pose, blocks[0], scn.v, scn.w, scn.dt)
            ids0 = onp.measurement_update(st0, scn.observations[0])
            assert np.array_equal(fs.last_assoc.cpu().numpy(), ids0)
            assert np.max(np.abs(w0 - st0.weight) / st0.weight) < 1e-4
            assert np.max(np.abs(mean5 - st0.mean) / np.maximum(np.abs(st0.mean), 1e-3)) < 1e-5
            assert np.max(np.abs(covp - st0.cov[..., :2, :2])) < 1e-5
            assert np.max(np.abs(covc - st0.cov[..., 2:, 2:])) < 1e-5
        ids, wgt, anc, _ = onp.frame(st, scn.observations[t], blocks[t], scn.v, scn.w, scn.dt, urng2.random(),
                                     sequential_resample=False)
        match.append((float((fs.last_assoc.cpu().numpy() == ids).mean()),
                      float((fs.last_ancestors.cpu().numpy() == anc).mean())))
    assert min(m for m, _ in match) >= 0.90 and min(r for _, r in match) >= 0.90, match
    est, ref = np.array(fs.summary()), np.array(onp.summary(st.pose))
    assert np.max(np.abs(est - ref)) < 1e-3
    assert fs.stats()["flags"] == 0


# ---- PK_MODEL_TEXTBOOK: the textbook bearing model as an option -------------------------
@pytest.mark.parametrize("dtype,arith", [("f64", "f64"), ("f32", "f32")])
def test_textbook_measurement_model_against_oracle(lib, dtype, arith):
    """Robot-frame predicted bearing, Jacobian row [-dy/q, +dx/q], wrapped innovation (the reference's quirks F4 a/b/e
    switched off) on the circle trajectory, where the heading is not ~0 and the models really differ: the device
    against the NumPy restatement run with the same switch (no reference fixture can exist for a deviation).
    f64: every index exact, state to 1e-9; fp32 algebra: frame 0 exact, >= 90 % of all indices."""
    from device_harness import run_device
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c1", num_particles=64, num_landmarks=20, frames=40, trajectory="circle")
    to = onp.run_scenario(scn, record_landmarks_at=(39,), model="textbook")
    tr = run_device(scn, dtype, checkpoints=(39,), arithmetic=arith, measurement_model="textbook")
    ref_mode = onp.run_scenario(scn, model="reference")
    assert not np.array_equal(to["weight"], ref_mode["weight"])          # the switch changes the filter
    if dtype == "f64":
        assert np.array_equal(tr["assoc"], to["assoc"]) and np.array_equal(tr["ancestors"], to["ancestors"])
        assert np.max(np.abs(tr["pose_post"] - to["pose_post"])) < 1e-9
        big = to["weight"] > 1e-300
        assert _rel(tr["weight"][big], to["weight"][big]) < 1e-7
        assert np.max(np.abs(tr["lm_mean"][39] - to["lm_mean"][39])) < 1e-8
    else:
        assert np.array_equal(tr["assoc"][0], to["assoc"][0]) and np.array_equal(tr["ancestors"][0], to["ancestors"][0])
        assert float((tr["assoc"] == to["assoc"]).mean()) >= 0.90
        assert float((tr["ancestors"] == to["ancestors"]).mean()) >= 0.90


# ---- PK_MODEL_LOG_WEIGHTS: log-domain importance weights + log-sum-exp normaliser (north_star (3)) ---------------
def test_log_weight_normaliser_kernels(lib):
    """pk_log_weights_max / pk_log_weights_normalise against NumPy on log weights around -2000 (their linear forms
    are exactly 0.0 in fp64), with -inf entries (a zero factor), and on the all -inf case."""
    import ctypes
    import torch
    from parakeet_slam_b200 import _lib
    L = _lib.load()
    rs = np.random.RandomState(3)
    for M, all_inf in ((100003, False), (257, True)):
        lw = -2000.0 + 40.0 * rs.standard_normal(M)
        lw[::17] = -np.inf
        if all_inf:
            lw[:] = -np.inf
        pose = torch.zeros((M, 4), dtype=torch.float64, device="cuda")
        pose[:, 3] = torch.from_numpy(lw).cuda()
        mx = torch.zeros(1, dtype=torch.float64, device="cuda")
        out3 = torch.zeros(3, dtype=torch.float64, device="cuda")
        ws = torch.zeros(2 * 1024, dtype=torch.float64, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.pk_log_weights_max(_lib.ptr(pose), M, _lib.ptr(mx), _lib.ptr(ws), st), "max")
        _lib.check(L.pk_log_weights_normalise(_lib.ptr(pose), M, _lib.ptr(mx), _lib.ptr(out3), _lib.ptr(ws), st), "norm")
        w = pose[:, 3].cpu().numpy()
        o = out3.cpu().numpy()
        if all_inf:
            assert np.all(w == 1.0) and o[0] == M            # nothing to prefer: uniform
            continue
        assert float(mx.item()) == lw.max()
        want = np.exp(lw - lw.max())
        assert w.max() == 1.0 and np.all(w[::17] == 0.0)
        assert np.max(np.abs(w - want)) < 1e-14
        assert abs(o[0] - want.sum()) < 1e-9 * want.sum() and abs(o[1] - (want ** 2).sum()) < 1e-9 * (want ** 2).sum()


@pytest.mark.parametrize("dtype,arith", [("f64", "f64"), ("f32", "f32")])
def test_log_weights_filter_matches_linear_weights(lib, dtype, arith):
    """FastSLAM(weights="log") against the same filter with the reference's linear weights on a reference fixture's
    scenario.  Frame 0 (identical state going in): identical associations, log weight == log(linear weight), identical
    ancestors.  Later frames: the two filters resample from weights that differ where the linear product underflows
    (log mode keeps what fp64 flushes to zero -- its purpose), so lineages may part; associations of identical
    lineages stay identical, which shows as a high overall agreement."""
    from device_harness import run_device
    g = load_trace("trace_corridor_m32_t60")
    scn = scenario_from_trace(g)
    lin = run_device(scn, dtype, arithmetic=arith)
    log = run_device(scn, dtype, arithmetic=arith, weights="log")
    assert np.array_equal(lin["assoc"][0], log["assoc"][0])
    w = lin["weight"][0]
    big = w > 1e-300
    assert big.any()
    assert np.max(np.abs(log["weight"][0][big] - np.log(w[big]))) < (1e-9 if arith == "f64" else 1e-4)
    assert np.array_equal(lin["ancestors"][0], log["ancestors"][0])
    same_anc = float((lin["ancestors"] == log["ancestors"]).mean())
    same_assoc = float((lin["assoc"] == log["assoc"]).mean())
    print("log vs linear weights (%s/%s): ancestors %.4f, associations %.4f identical" % (dtype, arith, same_anc, same_assoc))
    assert same_assoc >= 0.90
    n_eff, lse = log["filter"].effective_sample_size()
    assert 1.0 <= n_eff <= scn.num_particles and np.isfinite(lse)
