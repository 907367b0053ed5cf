"""Parity at sizes the reference cannot reach: the device against the NumPy oracle on a sampled
subset, and size-independent properties at BASELINE config-2 size (2^20 particles x 64 landmarks)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(M, N, dtype, frames, noise="philox", arithmetic="f64", fs_kw={}, **kw):
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c2", num_particles=M, num_landmarks=N, frames=frames, **kw)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    urng = random.Random(4)
    fs = FastSLAM(make_features(scn), num_particles=M, dtype=dtype, noise=noise, seed=11, uniform=urng.random, clock=clk,
                  arithmetic=arithmetic, **fs_kw)
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    return scn, fs, clk, tw


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-5)])
def test_medium_size_against_oracle(dtype, tol):
    """4096 particles x 64 landmarks x 8 blobs, injected noise, 6 frames, whole state compared."""
    import torch
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import DT_NSEC
    M, N, T = 4096, 64, 6
    noise_rs = np.random.RandomState(77)
    blocks = [noise_rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)
    scn, fs, clk, tw = _make(M, N, dtype, T, noise=lambda m: next(it), sigma_color=1.5)
    fs.keep_trace = True
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng = random.Random(4)
    idx_match = []
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        fs.measurement_update(scn.observations[t])
        fs.low_variance_resample()
        ids, wgt, anc, pose_pre = onp.frame(st, scn.observations[t], blocks[t], scn.v, scn.w, scn.dt, urng.random(),
                                            sequential_resample=False)
        a = fs.last_assoc.cpu().numpy()
        r = fs.last_ancestors.cpu().numpy()
        idx_match.append(((a == ids).mean(), (r == anc).mean()))
        if dtype == "f64":
            assert np.array_equal(a, ids), "frame %d" % t
            assert np.array_equal(r, anc), "frame %d" % t
            w = fs.last_weight.cpu().numpy()
            big = wgt > 1e-300
            assert np.max(np.abs(w[big] - wgt[big]) / wgt[big]) < 1e-8
            assert np.max(np.abs(fs.pose[:, :3].cpu().numpy() - st.pose)) < tol
    if dtype == "f64":
        mean5, covp, covc, meta, ids_, nlive = fs.export_maps()
        assert np.max(np.abs(mean5 - st.mean) / np.maximum(np.abs(st.mean), 1e-3)) < 1e-8
        assert np.max(np.abs(covp - st.cov[..., :2, :2])) < 1e-10
        assert np.max(np.abs(covc - st.cov[..., 2:, 2:])) < 1e-10
        assert np.array_equal(meta & 0xFFFFFF, st.count)
    else:
        # fp32 landmark storage: >= 90 % of association / resampling indices identical (BASELINE target)
        assert min(m for m, _ in idx_match) >= 0.90 and min(r for _, r in idx_match) >= 0.90, idx_match
        assert idx_match[0] == (1.0, 1.0)
    s = fs.stats()
    assert s["flags"] == 0


@pytest.mark.parametrize("arithmetic,tol_weight,tol_mean", [("f64", 1e-9, 1e-6), ("f32", 1e-3, 1e-5)])
def test_config2_size_properties(arithmetic, tol_weight, tol_mean):
    """2^20 particles x 64 landmarks x 8 blobs, fp32 records -- with fp64 landmark algebra and with the fp32 algebra
    bench.py times (FastSLAM(dtype="f32", arithmetic="f32")): properties that do not need the oracle at full size, plus
    the oracle on a random subset of 2048 particles (particles are independent up to resampling): >= 99.9 % of their
    associations identical, weights and updated landmark means of the identical rows to the stated tolerance."""
    import torch
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import DT_NSEC
    M, N, T = 1 << 20, 64, 3
    scn, fs, clk, tw = _make(M, N, "f32", T, sigma_color=1.0, arithmetic=arithmetic)
    fs.keep_trace = True
    rs = np.random.RandomState(3)
    sub = np.sort(rs.choice(M, 2048, replace=False))
    for t in range(T):
        clk.ns += DT_NSEC
        pose_before = fs.pose[:, :3].clone()
        fs.motion_update(tw)
        pose_motion = fs.pose[:, :3].cpu().numpy()
        maps_before = fs.export_maps(0, 0)  # (shape only)
        # oracle twin of the sampled particles: same pose after motion, same maps
        sub_t = torch.from_numpy(sub).cuda()
        mean5, covp, covc, meta, ids_, nlive = _export_subset(fs, sub)
        st = onp.OracleState(len(sub), scn.landmarks)
        st.pose = pose_motion[sub].copy()
        st.mean = mean5.copy()
        st.cov[...] = 0.0
        st.cov[..., :2, :2] = covp
        st.cov[..., 2:, 2:] = covc
        st.count = (meta & 0xFFFFFF).astype(np.int64)
        fs.measurement_update(scn.observations[t])
        ids = onp.measurement_update(st, scn.observations[t])
        a = fs.last_assoc[sub_t].cpu().numpy()
        w = fs.last_weight[sub_t].cpu().numpy()
        assert (a == ids).mean() >= 0.999
        same = (a == ids).all(axis=1)
        big = same & (st.weight > 1e-300)
        # fp64 algebra on the stored f32 state: 1e-9; fp32 algebra: 1e-3 (a weight is the product of eight factors
        # exp(-maha/2) with maha ~ 30 in fp32: measured 1.1e-4), 1e-5 (means, BASELINE tolerance)
        assert np.max(np.abs(w[big] - st.weight[big]) / st.weight[big]) < tol_weight
        mean_after = _export_subset(fs, sub)[0]
        assert np.max(np.abs(mean_after[same] - st.mean[same]) / np.maximum(np.abs(st.mean[same]), 1e-3)) < tol_mean
        # ---- resampling: exact properties at full size ------------------------------------------------
        weights = fs.pose[:, 3].cpu().numpy().copy()
        pool_before = _block_checksums(fs)
        slot_before = fs.slot.cpu().numpy().copy()
        pose_pre = fs.pose.cpu().numpy().copy()
        u = random.Random(4)
        fs.low_variance_resample()
        anc = fs.last_ancestors.cpu().numpy()
        assert np.all(np.diff(anc) >= 0) and anc.min() >= 0 and anc.max() < M          # sorted, in range
        u01 = _nth_uniform(4, t)
        want = onp.resample_searchsorted(weights, u01)
        assert (anc != want).sum() <= 2      # identical up to exact near-ties of the threshold comparison
        # copy-on-resample: every output particle owns a distinct block that holds its ancestor's map
        slot_after = fs.slot.cpu().numpy()
        assert len(np.unique(slot_after)) == M
        pool_after = _block_checksums(fs)
        assert np.array_equal(pool_after[slot_after], pool_before[slot_before[anc]])
        assert np.array_equal(fs.pose[:, :3].cpu().numpy(), pose_pre[anc, :3])
        # summary == plain mean (reference :254-276)
        x, y, h = fs.summary()
        p = fs.pose.cpu().numpy()
        assert abs(x - p[:, 0].mean()) < 1e-9 and abs(y - p[:, 1].mean()) < 1e-9
        assert abs(h - np.arctan2(np.sin(p[:, 2]).sum(), np.cos(p[:, 2]).sum())) < 1e-9
    assert fs.stats()["flags"] == 0


def _nth_uniform(seed, n):
    r = random.Random(seed)
    v = None
    for _ in range(n + 1):
        v = r.random()
    return v


def _block_checksums(fs):
    import torch
    words = fs._pool.view(torch.int32).view(fs.num_particles, fs.block_bytes // 4).to(torch.int64)
    mult = torch.arange(1, words.shape[1] + 1, device=words.device, dtype=torch.int64)
    return ((words * mult).sum(dim=1)).cpu().numpy()


def _export_subset(fs, sub):
    """Maps of an arbitrary subset of particles (export_maps works on ranges)."""
    outs = [fs.export_maps(int(i), 1) for i in sub[:0]]
    import torch
    # gather the subset's blocks into a temporary contiguous filter view: export ranges around each index
    mean5 = np.zeros((len(sub), fs.capacity, 5))
    covp = np.zeros((len(sub), fs.capacity, 2, 2))
    covc = np.zeros((len(sub), fs.capacity, 3, 3))
    meta = np.zeros((len(sub), fs.capacity), dtype=np.int32)
    ids = np.zeros((len(sub), fs.capacity), dtype=np.int32)
    nlive = np.zeros(len(sub), dtype=np.int32)
    # export everything once in chunks and pick rows (2^20 x 64 x 18 doubles would be 9.7 GB: chunk it)
    chunk = 1 << 16
    for lo in range(0, fs.num_particles, chunk):
        sel = np.nonzero((sub >= lo) & (sub < lo + chunk))[0]
        if len(sel) == 0:
            continue
        m5, cp, cc, me, idd, nl = fs.export_maps(lo, min(chunk, fs.num_particles - lo))
        rows = sub[sel] - lo
        mean5[sel], covp[sel], covc[sel], meta[sel], ids[sel], nlive[sel] = m5[rows], cp[rows], cc[rows], me[rows], idd[rows], nl[rows]
    return mean5, covp, covc, meta, ids, nlive


@pytest.mark.parametrize("n_colours", [24, 6, 3, 2, "mixed"])
def test_colour_ambiguous_maps_against_oracle(n_colours):
    """Landmarks that share colours: every blob has several colour-compatible landmarks, so the
    bearing / position terms decide.  Exercises the two-candidate path (24 colours for 48 landmarks), the
    shared-memory hit list (6 colours, 8 landmarks each; 3 colours, 16 each = a full list) and the key re-walk
    (2 colours, 24 each: more hits than the 16-entry list holds);
    "mixed": ONE colour shared by 24 landmarks, the other 24 unique -- items with more hits than the list holds
    (key re-walk) and items with one hit sit in the same warp, which is the case the warp collectives of the
    candidate loop must survive (they are executed by every lane, outside the per-lane conditions)."""
    import torch
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import DT_NSEC
    M, N, T = 2048, 48, 5
    noise_rs = np.random.RandomState(5)
    blocks = [noise_rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)
    scn, fs, clk, tw = _make(M, N, "f64", T, noise=lambda m: next(it))
    # recolour the world: n_colours distinct colours, jittered by < 2 units
    rs = np.random.RandomState(8)
    if n_colours == "mixed":
        palette = rs.uniform(20, 235, (25, 3))
        which = np.where(np.arange(N) % 2 == 0, 0, 1 + np.arange(N) // 2)   # every other landmark: colour 0
    else:
        palette = rs.uniform(20, 235, (n_colours, 3))
        which = np.arange(N) % n_colours
    scn.landmarks[:, 2:5] = palette[which] + rs.uniform(-1.5, 1.5, (N, 3))
    for t in range(T):
        scn.observations[t, :, 1:4] = scn.landmarks[scn.obs_landmark[t], 2:5] + rs.normal(0, 0.3, (8, 3))
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    fs2 = FastSLAM(make_features(scn), num_particles=M, dtype="f64", noise=lambda m: next(it), uniform=random.Random(4).random,
                   clock=clk)
    fs2.last_control = tw
    fs2.keep_trace = True
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng = random.Random(4)
    evals = 0
    for t in range(T):
        clk.ns += DT_NSEC
        fs2.motion_update(tw)
        fs2.measurement_update(scn.observations[t])
        evals = max(evals, fs2.stats()["evaluated"] / M)
        fs2.low_variance_resample()
        ids, wgt, anc, _ = onp.frame(st, scn.observations[t], blocks[t], scn.v, scn.w, scn.dt, urng.random(),
                                     sequential_resample=False)
        assert np.array_equal(fs2.last_assoc.cpu().numpy(), ids), "frame %d" % t
        assert np.array_equal(fs2.last_ancestors.cpu().numpy(), anc), "frame %d" % t
        w = fs2.last_weight.cpu().numpy()
        big = wgt > 1e-300
        assert np.max(np.abs(w[big] - wgt[big]) / wgt[big]) < 1e-8
    assert evals > (12 if n_colours in (24, "mixed") else 40)   # the ambiguous paths really ran
    mean5 = fs2.export_maps()[0]
    assert np.max(np.abs(mean5 - st.mean) / np.maximum(np.abs(st.mean), 1e-3)) < 1e-8


@pytest.mark.parametrize("K", [1, 3, 33, 40, 64])
def test_blob_counts_against_oracle(K):
    """Scans with 1, 3 (partial lane groups), 33/40/64 blobs (two items per lane) -- several blobs hit the
    same landmark, so the sequential same-landmark updates (finding F2) are exercised too."""
    from oracle import fastslam_np as onp
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    M, N, T = 1001, 24, 4
    scn = make_scenario("c2", num_particles=M, num_landmarks=N, obs_per_frame=min(K, N), frames=T)
    rs = np.random.RandomState(K)
    obs = np.zeros((T, K, 4))
    for t in range(T):
        pick = np.concatenate([np.arange(min(K, N)), rs.randint(0, min(K, N), max(0, K - N))])
        obs[t] = scn.observations[t][pick]
        obs[t, :, 0] += rs.normal(0, 0.01, K)        # repeated blobs: same landmark, slightly different reading
        obs[t, :, 1:] += rs.normal(0, 0.2, (K, 3))

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    blocks = [rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)
    fs = FastSLAM(make_features(scn), num_particles=M, dtype="f64", noise=lambda m: next(it),
                  uniform=random.Random(2).random, clock=clk)
    fs.keep_trace = True
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng = random.Random(2)
    same = 0
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        fs.measurement_update(obs[t])
        same += fs.stats()["same_landmark"]
        fs.low_variance_resample()
        ids, wgt, anc, _ = onp.frame(st, obs[t], blocks[t], scn.v, scn.w, scn.dt, urng.random(), sequential_resample=False)
        assert np.array_equal(fs.last_assoc.cpu().numpy(), ids), "frame %d" % t
        assert np.array_equal(fs.last_ancestors.cpu().numpy(), anc), "frame %d" % t
        w = fs.last_weight.cpu().numpy()
        big = wgt > 1e-300
        assert np.max(np.abs(w[big] - wgt[big]) / wgt[big]) < 1e-8
    mean5, covp, covc, meta, _, _ = fs.export_maps()
    assert np.max(np.abs(mean5 - st.mean) / np.maximum(np.abs(st.mean), 1e-3)) < 1e-8
    assert np.max(np.abs(covc - st.cov[..., 2:, 2:])) < 1e-10
    assert np.array_equal(meta & 0xFFFFFF, st.count)
    if K > N:
        assert same > 0


def test_empty_scan_and_too_many_blobs():
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import messages
    from parakeet_slam_b200.scenario import make_scenario
    scn = make_scenario("c1", num_particles=64, frames=1)
    fs = FastSLAM(make_features(scn), num_particles=64, noise="philox")
    fs.keep_trace = True

    class V(object):
        last_sensor_reading = messages.VizScan()
    before = fs.export_maps()[0]
    fs.cam_cb(V())                               # no blobs: weights stay 1, resampling keeps everyone
    assert np.array_equal(fs.last_ancestors.cpu().numpy(), np.arange(64))
    assert np.array_equal(fs.export_maps()[0], before)
    assert fs.particles[3].weight == 1.0 and fs.particles[3].next_id == 21
    with pytest.raises(ValueError):
        fs.measurement_update(np.zeros((65, 4)))
    V.last_sensor_reading = None
    with pytest.raises(AttributeError):          # as the reference: scan.observes on None (:344)
        fs.cam_cb(V())


def _spawn_twin(fs, lo, count, capacity):
    """NumPy-oracle twin of particles [lo, lo + count) of a spawn-mode filter, built from the exported device state
    (poses, maps with signed ids, orphaned readings, next_id)."""
    from oracle import fastslam_np as onp
    mean5, covp, covc, meta, ids, nlive = fs.export_maps(lo, count)
    st = onp.OracleState(count, None, capacity=capacity)
    st.pose = fs.pose[lo:lo + count, :3].cpu().numpy().copy()
    st.next_id = fs.aux[lo:lo + count, 1].cpu().numpy().astype(np.int64)
    live = np.arange(capacity)[None, :] < nlive[:, None]
    st.live = live
    st.mean = np.where(live[..., None], mean5, 0.0)
    st.cov[..., :2, :2] = covp
    st.cov[..., 2:, 2:] = covc
    st.cov[~live] = 0.0
    st.count = np.where(live, meta & 0xFFFFFF, 0).astype(np.int64)
    st.potential = live & ((meta & 0x20000000) != 0)
    st.immutable = live & ((meta & 0x10000000) != 0)
    st.ids = np.where(live, np.abs(ids), 0).astype(np.int64)
    rows, _ = fs.export_orphans(lo, count)
    st.orphans = [[(int(r[7]), float(r[0]), float(r[1]), float(np.arctan2(r[3], r[2])), float(r[4]), float(r[5]),
                    float(r[6])) for r in rr] for rr in rows]
    return st


def test_config3_size_spawn_properties():
    """BASELINE config 3 shape -- 2^22 particles, capacity 256, UNKNOWN map, spawn mode (82 GB of particle blocks):
    size-independent properties of the new-landmark path (id bookkeeping along every lineage, sign <-> potential
    flag, block permutation through the resamples, counters that add up), plus the NumPy oracle in spawn mode on 1024
    sampled particles rebuilt from the device state before every measurement update."""
    import torch
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    free, _ = torch.cuda.mem_get_info()
    M = 1 << 22
    if free < 110e9:
        M = 1 << 20
    N, T, K = 256, 5, 8
    scn = make_scenario("c3", num_particles=M, num_landmarks=N, frames=T, obs_per_frame=K)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    urng = random.Random(9)
    fs = FastSLAM([], num_particles=M, capacity=N, dtype="f32", noise="philox", seed=5, uniform=urng.random, clock=clk,
                  spawn=True, orphan_capacity=32)
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    fs.keep_trace = True
    spawned = orphaned = 0
    ranges = [(int(lo), 256) for lo in np.linspace(0, M - 256, 4).astype(np.int64)]   # 1024 sampled particles
    id_match = []
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        twins = [_spawn_twin(fs, lo, cnt, N) for lo, cnt in ranges]      # oracle twins of the sampled particles
        fs.measurement_update(scn.observations[t])
        for (lo, cnt), st in zip(ranges, twins):
            ids = onp.measurement_update(st, scn.observations[t], spawn=True, orphan_capacity=32)
            a = fs.last_assoc[lo:lo + cnt].cpu().numpy()
            id_match.append(float((a == ids).mean()))
            same = (a == ids).all(axis=1)
            aux_s = fs.aux[lo:lo + cnt].cpu().numpy()
            # identical association rows => identical bookkeeping: ids consumed, landmarks alive, new landmark ids
            assert np.array_equal(aux_s[same, 1], st.next_id[same]), "frame %d next_id" % t
            assert np.array_equal(aux_s[same, 0], st.live[same].sum(axis=1)), "frame %d n_live" % t
            mean5, _, _, _, ids_dev, nl = fs.export_maps(lo, cnt)
            for i in np.nonzero(same)[0][:64]:
                n = int(nl[i])
                want = np.where(st.potential[i, :n], -st.ids[i, :n], st.ids[i, :n])
                assert np.array_equal(ids_dev[i, :n], want), "frame %d particle %d landmark ids" % (t, lo + i)
                if n:
                    assert np.max(np.abs(mean5[i, :n] - st.mean[i, :n]) / np.maximum(np.abs(st.mean[i, :n]), 1.0)) < 1e-4
        s = fs.stats()
        assert s["matched"] + s["unmatched"] == M * K
        assert s["spawned"] + s["orphaned"] == s["unmatched"]      # every unseen blob ends in exactly one of the two
        # allowed: ring expiry (16) and, with fp32 storage, PK_FLAG_SINGULAR_COV (1): a landmark triangulated from two
        # nearly coincident rays sits almost on the robot, its first update collapses the position covariance to
        # ~0.1/h^2 along h, and the fp32 lower triangle can then round to det <= 0 (the likelihood is NaN -> no match,
        # as `nan > max` is False in the reference loop)
        assert (s["flags"] & ~(16 | 1)) == 0, s["flags"]
        spawned += s["spawned"]
        orphaned += s["orphaned"]
        fs.low_variance_resample()
    assert spawned > 0 and orphaned > 0
    # sampled-particle oracle comparison (spawn mode, fp32 records): >= 99.9 % of the association ids identical
    assert min(id_match) >= 0.999, id_match
    aux = fs.aux.cpu().numpy()
    n_live, next_id = aux[:, 0].astype(np.int64), aux[:, 1].astype(np.int64)
    assert n_live.min() >= 0 and n_live.max() <= N
    # along every surviving lineage: ids consumed = readings stored + landmarks spawned
    sample = np.linspace(0, M - 1, 2048).astype(np.int64)
    for i in sample[:64]:
        rows, totals = fs.export_orphans(int(i), 1)
        assert next_id[i] - 1 == int(totals[0]) + n_live[i], (i, next_id[i], totals[0], n_live[i])
        mean5, covp, covc, meta, ids, nl = fs.export_maps(int(i), 1)
        live_ids = ids[0, :int(nl[0])]
        assert len(set(np.abs(live_ids))) == len(live_ids) and (live_ids != 0).all()
        assert np.all((live_ids < 0) == ((meta[0, :int(nl[0])] & 0x20000000) != 0))    # sign <-> potential flag
        assert np.isfinite(mean5[0, :int(nl[0])]).all()
    slot = fs.slot.cpu().numpy()
    assert len(np.unique(slot)) == M                               # blocks stay a permutation through the resamples


@pytest.mark.parametrize("K", [3, 8, 40, 64])
def test_blob_counts_f32_arithmetic(K):
    """The fp32-algebra instantiation over the same blob counts (one and two items per lane, repeated blobs on one
    landmark): frame 0 identical to the oracle, >= 90 % of all indices, weights of frame 0 to 1e-3, no flags."""
    from oracle import fastslam_np as onp
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    M, N, T = 1001, 24, 4
    scn = make_scenario("c2", num_particles=M, num_landmarks=N, obs_per_frame=min(K, N), frames=T)
    rs = np.random.RandomState(K)
    obs = np.zeros((T, K, 4))
    for t in range(T):
        pick = np.concatenate([np.arange(min(K, N)), rs.randint(0, min(K, N), max(0, K - N))])
        obs[t] = scn.observations[t][pick]
        obs[t, :, 0] += rs.normal(0, 0.01, K)
        obs[t, :, 1:] += rs.normal(0, 0.2, (K, 3))

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)
    clk = Clk()
    blocks = [rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)
    fs = FastSLAM(make_features(scn), num_particles=M, dtype="f32", arithmetic="f32", noise=lambda m: next(it),
                  uniform=random.Random(2).random, clock=clk)
    fs.keep_trace = True
    tw = messages.Twist()
    tw.linear.x, tw.angular.z = scn.v, scn.w
    fs.last_control = tw
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng = random.Random(2)
    match = []
    for t in range(T):
        clk.ns += DT_NSEC
        fs.motion_update(tw)
        fs.measurement_update(obs[t])
        assert fs.stats()["flags"] == 0
        fs.low_variance_resample()
        ids, wgt, anc, _ = onp.frame(st, obs[t], blocks[t], scn.v, scn.w, scn.dt, urng.random(), sequential_resample=False)
        a, r = fs.last_assoc.cpu().numpy(), fs.last_ancestors.cpu().numpy()
        match.append((float((a == ids).mean()), float((r == anc).mean())))
        if t == 0:
            assert np.array_equal(a, ids) and np.array_equal(r, anc)
            w = fs.last_weight.cpu().numpy()
            big = wgt > 1e-300
            assert np.max(np.abs(w[big] - wgt[big]) / wgt[big]) < 1e-3
    assert min(m for m, _ in match) >= 0.90 and min(r for _, r in match) >= 0.90, match


def test_block_copies_on_the_copy_stream_change_nothing():
    """copy-on-resample (deepcopy, prkt_core_v2.py:243) runs on the filter's copy stream beside the next frame's motion
    update: the state must be bit-identical to the single-stream order, also when the pool is read straight after a
    resampling, when two resamplings follow each other, and when weights are degenerate (most blocks copied)."""
    import torch
    from parakeet_slam_b200.scenario import DT_NSEC
    M, N, T = 1 << 16, 64, 12
    runs = []
    for overlap in (True, False):
        scn, fs, clk, tw = _make(M, N, "f32", T, arithmetic="f32", fs_kw=dict(overlap_copy=overlap))
        assert (fs._copy_stream is not None) == overlap
        snaps = []
        for t in range(T):
            clk.ns += DT_NSEC
            fs.motion_update(tw)
            fs.measurement_update(scn.observations[t])
            if t % 4 == 3:   # degenerate weights: every 16th particle survives, 15/16 of the blocks are copied
                fs.pose[:, 3] = torch.where(torch.arange(M, device="cuda") % 16 == 0, 1.0, 0.0).to(torch.float64)
            fs.low_variance_resample()
            if t == 5:
                fs.low_variance_resample()   # twice in a row: the second must not overtake the first one's copies
            if t in (3, 5, T - 1):           # read the pool right behind the resampling
                snaps.append([a.copy() for a in fs.export_maps(0, 4096)])
        torch.cuda.synchronize()
        runs.append((fs.pose.cpu().numpy(), fs.slot.cpu().numpy(), fs.aux.cpu().numpy(), fs._pool.cpu().numpy(), snaps,
                     fs.stats()))
    a, b = runs
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    for sa, sb in zip(a[4], b[4]):
        for x, y in zip(sa, sb):
            assert np.array_equal(x, y)
    assert a[5] == b[5]


@pytest.mark.parametrize("dtype,arithmetic", [("f64", "f64"), ("f32", "f32")])
def test_colour_screen_contains_the_gate_on_its_boundary(dtype, arithmetic):
    """The byte-key screen is a bound on the distance of two 8-bit keys that must CONTAIN the reference's colour gate
    (`abs(colour distance) > 300` rejects, prkt_core_v2.py:441).  Hard cases: the blob's colour differs from the
    landmark's by the same amount in every channel (largest L1 distance for a given squared distance), with landmark
    colours at x.5 (key rounding) and near 0 / 255 (key clamping).  Blobs just inside the gate must still be matched, blobs
    just outside must not -- exactly as the oracle decides."""
    from oracle import fastslam_np as onp
    from parakeet_slam_b200.scenario import DT_NSEC
    M, N, T = 1024, 48, 4
    noise_rs = np.random.RandomState(15)
    blocks = [noise_rs.standard_normal((M, 3)) for _ in range(T)]
    it = iter(blocks)
    scn, fs, clk, tw = _make(M, N, dtype, T, noise=lambda m: next(it))
    rs = np.random.RandomState(18)
    col = np.floor(rs.uniform(30, 225, (N, 3))) + 0.5          # x.5: the key rounds to even
    col[::7] = rs.uniform(0.0, 6.0, (len(col[::7]), 3))         # near the clamp at 0
    col[3::7] = rs.uniform(249.0, 255.0, (len(col[3::7]), 3))   # ... and at 255
    scn.landmarks[:, 2:5] = col
    # per blob: a signed offset of equal size per channel, squared distance 3 d^2 around the gate (d = 10 <=> 300)
    d = np.array([9.9, 9.99, 9.999, 10.001, 10.01, 10.2, 9.5, 0.3])
    for t in range(T):
        sign = rs.choice([-1.0, 1.0], (8, 3))
        lmc = scn.landmarks[scn.obs_landmark[t], 2:5]
        scn.observations[t, :, 1:4] = lmc + sign * np.roll(d, t)[:, None]
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    fs2 = FastSLAM(make_features(scn), num_particles=M, dtype=dtype, arithmetic=arithmetic, noise=lambda m: next(it),
                   uniform=random.Random(4).random, clock=clk)
    fs2.last_control = tw
    fs2.keep_trace = True
    st = onp.OracleState(M, scn.landmarks, preset_covar=scn.preset_covar)
    urng = random.Random(4)
    seen_match = seen_gate = 0
    for t in range(T):
        clk.ns += DT_NSEC
        fs2.motion_update(tw)
        fs2.measurement_update(scn.observations[t])
        fs2.low_variance_resample()
        ids, wgt, anc, _ = onp.frame(st, scn.observations[t], blocks[t], scn.v, scn.w, scn.dt, urng.random(),
                                     sequential_resample=False)
        got = fs2.last_assoc.cpu().numpy()
        if arithmetic == "f64":
            assert np.array_equal(got, ids), "frame %d" % t
            assert np.array_equal(fs2.last_ancestors.cpu().numpy(), anc), "frame %d" % t
        else:
            # fp32 algebra evaluates the gate on fp32 colours: a blob within 1e-3 of the gate may fall on either side
            dd = np.roll(d, t)
            sure = np.abs(3.0 * dd * dd - 300.0) > 0.5
            assert np.array_equal(got[:, sure], ids[:, sure]), "frame %d" % t
            break   # (the two filters' maps differ from here on if a borderline blob was decided differently)
        seen_match += int((ids > 0).sum())
        seen_gate += int((ids == 0).sum())
    if arithmetic == "f64":
        assert seen_match > 0 and seen_gate > 0   # both sides of the gate were exercised
