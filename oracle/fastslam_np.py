"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, fp64) of the reference's
FastSLAM hot path, vectorised over particles.

Nothing under ``parakeet_slam_b200/`` may import this.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm use it,
and only as the checker / the timed CPU baseline -- never as the product path.

Parity status: PINNED.  This restatement is checked (``tests/test_oracle_golden.py``)
against golden traces produced by running the *unmodified reference sources*
(``/root/reference/src/prkt_core_v2.py``, ``matrix.py``, ``utils.py``) through
``oracle/ref_shim.py`` (``oracle/make_golden.py`` is the generating script, fixtures
under ``tests/golden/``), and, in the development container, directly against the
live reference (``tests/test_reference_shim.py``).

Every function cites the reference lines it follows (paths relative to
``/root/reference/src``).  Formulas are reproduced as written, including the ones
that differ from the textbook (SURVEY.md finding F4): world-frame predicted
bearing in the EKF update, ``[+dy/q, +dx/q]`` Jacobian row, Frobenius norm in the
importance factor, no angle wrapping of innovations.

Third-party arithmetic that is not in the reference repository is restated from
its published algorithm:
* ``scipy.stats.multivariate_normal.pdf`` (unpinned dependency; call sites
  ``prkt_core_v2.py:489-490, 543-544``): ``exp(-0.5*(k*log(2*pi) + log det + maha))``
  with the covariance symmetrised from its LOWER triangle (``eigh(lower=True)``).
* ``tf.transformations`` heading round trip (``utils.py:8-35``): see
  ``parakeet_slam_b200/rosless/transformations.py``.
* ``numpy.linalg.inv`` / ``numpy.dot`` / ``numpy.linalg.norm`` are used directly.
"""
from __future__ import annotations

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)

# Literals of the reference (SURVEY.md section 5 "Config / flags").
BEARING_GATE = 0.5            # prkt_core_v2.py:433
POSITION_GATE = math.pi / 2   # prkt_core_v2.py:474
COLOR_GATE = 300.0            # prkt_core_v2.py:441
NO_MATCH_WEIGHT = 0.1         # prkt_core_v2.py:857
QT_DIAG = 0.1                 # prkt_core_v2.py:50-53
PROMOTE_COUNT = 5             # prkt_core_v2.py:114


# --------------------------------------------------------------------------------------
# State
# --------------------------------------------------------------------------------------
class OracleState(object):
    """Structure-of-arrays particle set.

    pose [M,3] (x, y, heading as read back through the quaternion), weight [M],
    mean [M,N,5], cov [M,N,5,5], count [M,N] (``Feature.update_count``),
    immutable [M,N] bool, live [M,N] bool, next_id [M].
    Known-map mode: landmark id = slot + 1 (``load_feature_list`` ``:294-299``).
    """

    def __init__(self, num_particles, landmarks=None, preset_covar=0.25, immutable=False,
                 capacity=None):
        M = int(num_particles)
        n = 0 if landmarks is None else len(landmarks)
        N = n if capacity is None else int(capacity)
        self.pose = np.zeros((M, 3))                      # FilterParticle.__init__ :279-285
        self.weight = np.ones(M)                          # :288
        self.mean = np.zeros((M, N, 5))
        self.cov = np.zeros((M, N, 5, 5))
        self.count = np.zeros((M, N), dtype=np.int64)
        self.immutable = np.zeros((M, N), dtype=bool)
        self.live = np.zeros((M, N), dtype=bool)
        self.potential = np.zeros((M, N), dtype=bool)     # lives in potential_features, id = -(slot+1)  :287
        self.next_id = np.ones(M, dtype=np.int64)         # :292
        self.ids = np.zeros((M, N), dtype=np.int64)       # |reference id| of each slot (0 = empty)
        # spawn mode: hypothesis_set of every particle, insertion order,
        # entries (id, x, y, heading + bearing, r, g, b)   :291, :745
        self.orphans = [[] for _ in range(M)]
        if n:
            self.ids[:, :n] = np.arange(1, n + 1)[None]
            self.mean[:, :n] = np.asarray(landmarks, dtype=np.float64)[None]
            self.cov[:, :n] = (np.identity(5) * preset_covar)[None, None]
            self.immutable[:, :n] = bool(immutable)
            self.live[:, :n] = True
            self.next_id += n                             # :298-299

    @property
    def num_particles(self):
        return self.pose.shape[0]

    def copy(self):
        out = OracleState.__new__(OracleState)
        for k, v in self.__dict__.items():
            setattr(out, k, [list(o) for o in v] if k == "orphans" else v.copy())
        return out


# --------------------------------------------------------------------------------------
# Heading round trip  (utils.py:8-35 -> tf.transformations)
# --------------------------------------------------------------------------------------
def wrap_heading(h):
    """heading -> quaternion (0,0,sin h/2,cos h/2) -> heading, as the reference stores and
    re-reads it (``prkt_core_v2.py:206`` then ``:188/:404``).  Vectorised transcription of
    ``quaternion_from_euler(0,0,h)`` / ``quaternion_matrix`` / ``euler_from_matrix('sxyz')``."""
    h = np.asarray(h, dtype=np.float64)
    half = h / 2.0
    z = np.sin(half)           # cj*cs - sj*sc with ci=cj=1, si=sj=0  ->  sk
    w = np.cos(half)           # cj*cc + sj*ss                        ->  ck
    nq = z * z + w * w         # numpy.dot(q, q), q = (0, 0, z, w)
    s = np.sqrt(2.0 / nq)
    zs = z * s
    ws = w * s
    m10 = zs * ws              # q[0,1] + q[2,3], q[0,1] == 0
    m00 = 1.0 - zs * zs        # 1 - q[1,1] - q[2,2], q[1,1] == 0
    return np.arctan2(m10, m00)


# --------------------------------------------------------------------------------------
# Motion  (prkt_core_v2.py:148-208)
# --------------------------------------------------------------------------------------
def motion_sigmas(v, w):
    """Noise scales of ``motion_model`` (``:185,190,193``)."""
    sd = abs(.05 * v) + abs(.005 * w) + .0005
    sh = abs(.025 * w) + abs(.005 * v) + .0005
    return sd, sh


def motion_update(pose, noise, v, w, dt):
    """All particles through ``motion_model`` (``:168-208``) with standard normals
    ``noise[M,3]`` standing for the three ``normal(0, sigma, 1)`` draws (``sigma * z``)."""
    pose = np.asarray(pose, dtype=np.float64)
    sd, sh = motion_sigmas(v, w)
    dheading = w * dt                                   # :183
    ds = v * dt + sd * noise[:, 0]                      # :185-186
    h1 = pose[:, 2] + dheading / 2 + sh * noise[:, 1]   # :190-191
    h2 = h1 + dheading / 2 + sh * noise[:, 2]           # :193-194
    out = np.empty_like(pose)
    out[:, 0] = pose[:, 0] + ds * np.cos(h1)            # :198, :203
    out[:, 1] = pose[:, 1] + ds * np.sin(h1)            # :199, :204
    out[:, 2] = wrap_heading(h2)                        # :206
    return out


# --------------------------------------------------------------------------------------
# Association likelihood  (prkt_core_v2.py:383-544)
# --------------------------------------------------------------------------------------
def obs_direction(bearing):
    """``unit((cos b, sin b, 0.0))`` of ``closest_point`` (``:510``, ``utils.py:69-76``)."""
    c = math.cos(bearing)
    s = math.sin(bearing)
    length = math.sqrt(c * c + s * s + 0.0 * 0.0)
    inv = 1.0 / length
    return c * inv, s * inv


def _pdf2_lower(ex, ey, a, b10, d):
    """2-D normal pdf, covariance [[a, b10],[b10, d]] (lower triangle), SciPy formula."""
    det = a * d - b10 * b10
    with np.errstate(invalid="ignore", divide="ignore"):
        maha = (d * ex * ex - 2.0 * b10 * ex * ey + a * ey * ey) / det
        return np.exp(-0.5 * (2 * LOG_2PI + np.log(det) + maha))


def _pdf3_lower(e0, e1, e2, C):
    """3-D normal pdf with covariance symmetrised from the lower triangle of C[...,3,3]."""
    a, b, c = C[..., 0, 0], C[..., 1, 0], C[..., 2, 0]
    d, e, f = C[..., 1, 1], C[..., 2, 1], C[..., 2, 2]
    # symmetric matrix [[a,b,c],[b,d,e],[c,e,f]]
    c00 = d * f - e * e
    c01 = c * e - b * f
    c02 = b * e - c * d
    c11 = a * f - c * c
    c12 = b * c - a * e
    c22 = a * d - b * b
    det = a * c00 + b * c01 + c * c02
    with np.errstate(invalid="ignore", divide="ignore"):
        maha = (c00 * e0 * e0 + c11 * e1 * e1 + c22 * e2 * e2
                + 2.0 * (c01 * e0 * e1 + c02 * e0 * e2 + c12 * e1 * e2)) / det
        return np.exp(-0.5 * (3 * LOG_2PI + np.log(det) + maha))


def match_likelihood(pose, obs, mean, cov):
    """``probability_of_match`` (``:383-455``) for every (particle, blob, landmark).

    pose [M,3], obs [K,4] (bearing,r,g,b), mean [M,N,5], cov [M,N,5,5] -> L [M,K,N].
    """
    x = pose[:, 0][:, None, None]
    y = pose[:, 1][:, None, None]
    th = pose[:, 2][:, None, None]
    fx = mean[:, None, :, 0]
    fy = mean[:, None, :, 1]
    beta = obs[:, 0][None, :, None]

    pse = np.arctan2(fy - y, fx - x)                    # :408 / :473 (same expression)
    expected = pse - th                                 # :408
    del_bearing = beta - expected                       # :415
    gate_bearing = np.abs(del_bearing) > BEARING_GATE   # :433

    # prob_position_match :457-494 (observed robot-frame bearing used as if world frame)
    gate_pos = np.abs(pse - beta) > POSITION_GATE       # :474
    dirs = np.array([obs_direction(float(b)) for b in obs[:, 0]])
    cb = dirs[:, 0][None, :, None]
    sb = dirs[:, 1][None, :, None]
    magmag = (fx - x) * cb + (fy - y) * sb + 0.0 * 0.0  # :513, utils.py:37-43
    near_x = np.where(magmag < 0, x, x + cb * magmag)   # :515-522
    near_y = np.where(magmag < 0, y, y + sb * magmag)
    bp = _pdf2_lower(near_x - fx, near_y - fy, cov[:, None, :, 0, 0], cov[:, None, :, 1, 0],
                     cov[:, None, :, 1, 1])              # :482-490
    bp = np.where(gate_pos, 0.0, bp)
    bearing_prob = 500.0 * bp                           # :439

    dr = obs[:, 1][None, :, None] - mean[:, None, :, 2]
    dg = obs[:, 2][None, :, None] - mean[:, None, :, 3]
    db = obs[:, 3][None, :, None] - mean[:, None, :, 4]
    color_distance = dr * dr + dg * dg + db * db        # :425-427
    gate_color = np.abs(color_distance) > COLOR_GATE    # :441
    cp = _pdf3_lower(dr, dg, db, cov[:, None, :, 2:, 2:])  # :524-544
    color_prob = 500.0 * cp                             # :446

    with np.errstate(invalid="ignore", over="ignore", under="ignore"):
        L = bearing_prob * color_prob / 250000.0        # :455
    L = np.where(gate_bearing | gate_color, 0.0, L)
    return L


def associate(state: OracleState, obs):
    """``match_features_to_scan`` (``:317-351``) / ``match_one`` (``:353-381``): per blob the
    arg-max likelihood over live slots in slot order, strict ``>`` from 0.0 so the first
    maximum wins and "all exactly zero" gives id 0.  Returns ids [M,K] (slot+1, or 0)."""
    L = match_likelihood(state.pose, obs, state.mean, state.cov)
    L = np.where(state.live[:, None, :], L, 0.0)
    L = np.where(np.isnan(L), 0.0, L)   # ``nan > max`` is False in the reference loop
    best = np.argmax(L, axis=2)          # first occurrence of the maximum
    best_val = np.take_along_axis(L, best[:, :, None], axis=2)[:, :, 0]
    ids = np.where(best_val > 0.0, np.take_along_axis(state.ids, best, axis=1), 0)
    # potential features carry negative ids (:336, :367); with full features in the lower slots the slot
    # order is the reference's iteration order (feature_set first, then potential_features)
    pot = np.take_along_axis(state.potential, best, axis=1)
    ids = np.where(pot, -ids, ids)
    state.last_best_slot = best
    return ids.astype(np.int32), best_val


# --------------------------------------------------------------------------------------
# EKF update + importance weight  (prkt_core_v2.py:88-124, 748-849, 897-930)
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# Spawn mode  (prkt_core_v2.py:546-746 with the patches P1-P3 of SURVEY.md A.6, i.e. what
# oracle/ref_shim.apply_spawn_patches makes the reference execute)
# --------------------------------------------------------------------------------------
PAIR_GATE = 300.0 ** 0.5


def ray_intersect(x1, y1, b1, x3, y3, b3):
    """``ray_intersect`` (``:610-640``), operation for operation."""
    as_ = (x1, y1)
    ad_ = (math.cos(b1), math.sin(b1))
    bs_ = (x3, y3)
    bd_ = (math.cos(b3), math.sin(b3))
    if (ad_[1] * bd_[0] - ad_[0] * bd_[1]) == 0:
        return False
    v = ((ad_[0] * bs_[1] - ad_[1] * bs_[0] + ad_[1] * as_[0] - ad_[0] * as_[1]) /
         (ad_[1] * bd_[0] - ad_[0] * bd_[1]))
    if abs(ad_[1]) < abs(ad_[0]):
        u = (bs_[0] + bd_[0] * v - as_[0]) / (ad_[0])
    else:
        u = (bs_[1] + bd_[1] * v - as_[1]) / (ad_[1])
    return u >= 0 and v >= 0


def cross_readings(x1, y1, h1, x3, y3, h3):
    """``cross_readings`` (``:682-738``): line-line intersection; ``None`` for parallel lines."""
    x2 = x1 + math.cos(h1)
    y2 = y1 + math.sin(h1)
    x4 = x3 + math.cos(h3)
    y4 = y3 + math.sin(h3)
    t0 = x1 * y2 - y1 * x2
    t1 = x3 - x4
    t2 = x1 - x2
    t3 = x3 * y4 - x4 * y3
    t5 = y3 - y4
    t6 = y1 - y2
    den = t2 * t5 - t6 * t1
    if den == 0:
        return None
    return ((t0 * t1 - t2 * t3) / den, (t0 * t5 - t6 * t3) / den)


def add_hypothesis(state: OracleState, i, blob, gate=PAIR_GATE, orphan_capacity=None, flags=None):
    """``add_hypothesis`` (``:546-563``) for particle ``i`` and one unseen blob (bearing, r, g, b).
    Returns "spawned" or "orphaned"."""
    x, y, th = (float(v) for v in state.pose[i])
    beta, r, g, b = (float(v) for v in blob)
    ang = beta + th                                               # :603
    # find_nearest_reading (P1): hypothesis_set in insertion order, strict '<'
    min_d, min_o = float("inf"), None
    for o in state.orphans[i]:
        _, ox, oy, oang, orr, og, ob = o
        if ray_intersect(ox, oy, oang, x, y, ang):                # :605
            d = math.sqrt(math.pow(orr - r, 2) + math.pow(og - g, 2) + math.pow(ob - b, 2))   # :642-651
        else:
            d = float("inf")
        if d < min_d:
            min_d, min_o = d, o
    new_id = int(state.next_id[i])
    if min_o is not None and min_d <= gate:
        # add_new_feature :653-680
        _, ox, oy, oang, orr, og, ob = min_o
        inter = cross_readings(ox, oy, oang, x, y, ang)
        if inter is not None:
            live = state.live[i]
            n_live = int(live.sum())
            state.next_id[i] += 1                                 # :680
            if n_live >= live.shape[0]:
                if flags is not None:
                    flags.add("map_full")
                return "dropped"
            j = n_live
            state.mean[i, j] = (inter[0], inter[1], (orr + r) / 2, (og + g) / 2, (ob + b) / 2)
            state.cov[i, j] = np.identity(5)
            state.count[i, j] = 0
            state.immutable[i, j] = False
            state.live[i, j] = True
            state.potential[i, j] = True                          # potential_features[-new_id] :679
            state.ids[i, j] = new_id
            return "spawned"
        if flags is not None:
            flags.add("degenerate")   # the reference raises TypeError here (:665-667)
    # add_orphaned_reading :740-746 (P2: the pose is copied)
    state.orphans[i].append((new_id, x, y, ang, r, g, b))
    if orphan_capacity is not None and len(state.orphans[i]) > orphan_capacity:
        state.orphans[i].pop(0)      # ring of the device build (reported deviation)
        if flags is not None:
            flags.add("expired")
    state.next_id[i] += 1
    return "orphaned"


def wrap_pi(a):
    """Angle difference wrapped to (-pi, pi] (textbook model only; the reference never wraps, finding F4e)."""
    a = a - 2.0 * math.pi * np.rint(a / (2.0 * math.pi))
    return np.where(a <= -math.pi, a + 2.0 * math.pi, a)


def measurement_update(state: OracleState, obs, ids=None, spawn=False, gate=PAIR_GATE, orphan_capacity=None,
                       model="reference"):
    """One frame of ``cam_cb``'s per-particle body after the motion update: weights <- 1
    (``:73``), association of all blobs against the pre-update map (``:84``), then the K
    sequential updates in scan order (``:88-124``).  Mutates ``state``; returns ids [M,K].

    ``model="textbook"`` is NOT reference behaviour: the documented deviation ``PK_MODEL_TEXTBOOK`` of the device
    library -- robot-frame predicted bearing, Jacobian row ``[-dy/q, +dx/q]``, wrapped bearing
    innovation -- restated here so that the device build of it has something to be checked against."""
    textbook = model == "textbook"
    if model not in ("reference", "textbook"):
        raise ValueError("model must be 'reference' or 'textbook'")
    M = state.num_particles
    K = obs.shape[0]
    slots = None
    if ids is None:
        ids, _ = associate(state, obs)
        slots = state.last_best_slot
    state.weight = np.ones(M)
    Qt = np.identity(4) * QT_DIAG
    I5 = np.identity(5)
    rows = np.arange(M)
    for k in range(K):
        idk = ids[:, k]
        un = idk == 0
        # unseen blob: add_hypothesis -> add_orphaned_reading (:92-95, :546-563, :740-746)
        if spawn:
            for i in np.nonzero(un)[0]:
                add_hypothesis(state, int(i), obs[k], gate=gate, orphan_capacity=orphan_capacity)
        else:
            state.next_id[un] += 1
        factor = np.full(M, NO_MATCH_WEIGHT)            # :95, :851-857
        m = ~un
        if m.any():
            r = rows[m]
            j = (np.abs(idk[m]) - 1) if slots is None else slots[m, k]
            mu = state.mean[r, j]                       # [m,5]
            Sg = state.cov[r, j]                        # [m,5,5]
            px = state.pose[r, 0]
            py = state.pose[r, 1]
            dx = mu[:, 0] - px
            dy = mu[:, 1] - py
            zhat = np.stack([np.arctan2(dy, dx), mu[:, 2], mu[:, 3], mu[:, 4]], axis=1)  # :859-877
            q = dx ** 2 + dy ** 2                       # :785
            with np.errstate(divide="ignore", invalid="ignore"):
                hx = np.where(q == 0, 0.0, dy / q)      # :788-791
                hy = np.where(q == 0, 0.0, dx / q)      # :794-797
            if textbook:
                zhat[:, 0] = zhat[:, 0] - state.pose[r, 2]
                hx = -hx
            H = np.zeros((len(r), 4, 5))                # :799-802
            H[:, 0, 0] = hx
            H[:, 0, 1] = hy
            H[:, 1, 2] = 1.0
            H[:, 2, 3] = 1.0
            H[:, 3, 4] = 1.0
            Ht = np.transpose(H, (0, 2, 1))
            Q = H @ Sg @ Ht + Qt                        # :817-819
            Qinv = np.linalg.inv(Q)                     # :102, matrix.py:11-12
            Kg = Sg @ Ht @ Qinv                         # :833
            z = np.broadcast_to(obs[k], (len(r), 4))
            delz = z - zhat                             # :911, :846 (no wrapping)
            if textbook:
                delz[:, 0] = wrap_pi(delz[:, 0])
            mut = ~state.immutable[r, j]                # :909, :926
            new_mu = mu + np.einsum("mij,mj->mi", Kg, delz)          # :912-913
            new_Sg = (I5 - Kg @ H) @ Sg                              # :928-929
            state.mean[r[mut], j[mut]] = new_mu[mut]
            state.cov[r[mut], j[mut]] = new_Sg[mut]
            state.count[r[mut], j[mut]] += 2                         # :914, :930
            # importance_factor :835-849 (pre-update Q and zhat; Frobenius norm of Q)
            v1 = (2.0 * math.pi * np.sqrt(np.sum(np.abs(Q) ** 2, axis=(1, 2)))) ** -0.5
            expo = -0.5 * np.einsum("mi,mij,mj->m", delz, Qinv, delz)
            with np.errstate(under="ignore"):
                fm = v1 * np.exp(expo)
            neg = idk[m] < 0
            # potential feature :109-118: weight as if unseen; promoted when update_count > 5
            fm = np.where(neg, NO_MATCH_WEIGHT, fm)
            factor[m] = fm
            promote = neg & (state.count[r, j] > PROMOTE_COUNT)
            state.potential[r[promote], j[promote]] = False
        state.weight = state.weight * factor            # :124 / :95
    return ids


# --------------------------------------------------------------------------------------
# Resampling  (prkt_core_v2.py:210-252)
# --------------------------------------------------------------------------------------
def resample_sequential(weight, u01):
    """Literal restatement of the running-``step`` sweep (``:216-250``).  Pure-Python loop:
    use for small M; it is the definition the vectorised form is checked against."""
    M = len(weight)
    sum_ = 0
    for wgt in weight:
        sum_ += float(wgt)                               # :218-220
    range_ = sum_ / float(M)                             # :225
    step = u01 * range_                                  # :226
    anc = []
    count = 0
    for i in range(M):
        step = step - float(weight[i])                   # :238
        while step <= 0.0 and count < M:                 # :239
            anc.append(i)
            step += range_                               # :248
            count += 1
    return np.asarray(anc, dtype=np.int64)


def resample_searchsorted(weight, u01):
    """Vectorised equivalent: ``anc[k] = min{i : C_i >= u0 + k*r}`` with C the left-fold
    prefix sum (SURVEY.md finding F6; agrees with the sweep except at exact near-ties)."""
    weight = np.asarray(weight, dtype=np.float64)
    M = len(weight)
    C = np.cumsum(weight)
    total = C[-1]
    r = total / float(M)
    u0 = u01 * r
    if not total > 0.0:
        return np.zeros(M, dtype=np.int64)              # step == 0 <= 0: particle 0, M times
    targets = u0 + np.arange(M, dtype=np.float64) * r
    anc = np.searchsorted(C, targets, side="left")
    return np.minimum(anc, M - 1).astype(np.int64)


def apply_ancestors(state: OracleState, anc):
    """``temp_particles.append(deepcopy(particle))`` (``:243``) for every ancestor."""
    for name in ("pose", "weight", "mean", "cov", "count", "immutable", "live", "potential", "next_id", "ids"):
        setattr(state, name, getattr(state, name)[anc].copy())
    state.orphans = [list(state.orphans[int(a)]) for a in anc]


# --------------------------------------------------------------------------------------
# Summary  (prkt_core_v2.py:254-276)
# --------------------------------------------------------------------------------------
def summary(pose):
    """Unweighted mean x, mean y and circular-mean heading; left-fold sums as the loop."""
    M = float(pose.shape[0])
    xs = np.cumsum(pose[:, 0])[-1]
    ys = np.cumsum(pose[:, 1])[-1]
    hdx = np.cumsum(np.cos(pose[:, 2]))[-1]
    hdy = np.cumsum(np.sin(pose[:, 2]))[-1]
    return xs / M, ys / M, math.atan2(hdy, hdx)


# --------------------------------------------------------------------------------------
# One full frame  (cam_cb :59-137)
# --------------------------------------------------------------------------------------
def frame(state: OracleState, obs, noise, v, w, dt, u01, sequential_resample=None, spawn=False, orphan_capacity=None,
          model="reference"):
    """motion (``:75-77``) -> association + updates (``:84-124``) -> resample (``:137``).
    Returns (ids [M,K], pre-resample weights [M], ancestors [M], pose_pre [M,3])."""
    state.pose = motion_update(state.pose, noise, v, w, dt)
    pose_pre = state.pose.copy()
    ids = measurement_update(state, obs, spawn=spawn, orphan_capacity=orphan_capacity, model=model)
    wgt = state.weight.copy()
    if sequential_resample is None:
        sequential_resample = state.num_particles <= 4096
    anc = resample_sequential(wgt, u01) if sequential_resample else resample_searchsorted(wgt, u01)
    if len(anc) < state.num_particles:  # reference would shrink the list; pad like the device
        anc = np.concatenate([anc, np.full(state.num_particles - len(anc), state.num_particles - 1)])
    apply_ancestors(state, anc)
    return ids, wgt, anc, pose_pre


def run_scenario(scn, frames=None, num_particles=None, record_landmarks_at=(), chunk=None, potential_slots=(),
                 spawn=False, known_map=True, capacity=None, orphan_capacity=None, model="reference"):
    """Run the restatement over a ``Scenario`` (same trace layout as
    ``oracle.ref_driver.run_reference``)."""
    T = scn.frames if frames is None else frames
    M = scn.num_particles if num_particles is None else num_particles
    K = scn.obs_per_frame
    st = OracleState(M, scn.landmarks if known_map else None, preset_covar=scn.preset_covar, immutable=scn.immutable,
                     capacity=capacity)
    for j in potential_slots:
        st.potential[:, j] = True
    trace = dict(pose_pre=np.zeros((T, M, 3)), pose_post=np.zeros((T, M, 3)),
                 assoc=np.zeros((T, M, K), dtype=np.int32), weight=np.zeros((T, M)),
                 ancestors=np.zeros((T, M), dtype=np.int32), summary=np.zeros((T, 3)),
                 next_id=np.zeros((T, M), dtype=np.int64), lm_mean={}, lm_cov={}, lm_count={})
    stream = scn.motion_noise_stream(M)
    for t in range(T):
        noise = next(stream)
        ids, wgt, anc, pose_pre = frame(st, scn.observations[t], noise, scn.v, scn.w, scn.dt,
                                        float(scn.u01[t]), spawn=spawn, orphan_capacity=orphan_capacity, model=model)
        trace["assoc"][t] = ids
        trace["weight"][t] = wgt
        trace["ancestors"][t] = anc
        trace["pose_pre"][t] = pose_pre
        trace["pose_post"][t] = st.pose
        trace["summary"][t] = summary(st.pose)
        trace["next_id"][t] = st.next_id
        if t in record_landmarks_at:
            trace["lm_mean"][t] = st.mean.copy()
            trace["lm_cov"][t] = st.cov.copy()
            trace["lm_count"][t] = st.count.copy()
            trace.setdefault("lm_potential", {})[t] = st.potential.copy()
            trace.setdefault("lm_ids", {})[t] = np.where(st.potential, -st.ids, st.ids) * st.live
            trace.setdefault("orphans", {})[t] = [list(o) for o in st.orphans]
    trace["state"] = st
    return trace
