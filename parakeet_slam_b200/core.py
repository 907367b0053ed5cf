"""Drop-in replacement for the reference's ``prkt_core_v2`` module
(``/root/reference/src/prkt_core_v2.py``): same class names, constructor, methods and
attributes (``FastSLAM``, ``FilterParticle``, ``Feature``), with every per-particle loop
executed by hand-written sm_100a CUDA kernels reached through the C ABI of
``libparakeet_b200.so`` (``include/parakeet_b200.h``).

Filter state lives in PyTorch CUDA tensors (structure described in DESIGN.md); the library only
receives raw pointers.  There is no CPU path: constructing ``FastSLAM`` without a CUDA device
or without the shared library raises.

Reference behaviour that is kept on purpose (SURVEY.md findings F2-F8): association of all
blobs before any update, fp64-underflow match decision, world-frame predicted bearing in the
EKF update, ``[+dy/q, +dx/q]`` Jacobian row, Frobenius norm in the importance factor, no angle
wrapping, systematic resampling every frame with one ``random.random()`` draw, motion noise as
three NumPy normals per particle in index order from the global legacy stream, the previous
control being integrated by ``motion_update``.
"""
from __future__ import annotations

import ctypes
import math
import random as _pyrandom
import threading

import numpy as np

from . import _lib
from ._ros import Odometry, Publisher, Quaternion, Twist, now as _ros_now
from .particle_math import ParticleMathMixin

__all__ = ["FastSLAM", "FilterParticle", "Feature", "ParticleList", "Matrix"]


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_CTX = _NullCtx()


def Matrix(array_like):
    """``matrix.Matrix`` of the reference (``matrix.py:6-9``): a NumPy array."""
    return np.array(array_like)


# --------------------------------------------------------------------------------------------
# Heading <-> quaternion on the host, for the message views only (utils.py:8-35)
# --------------------------------------------------------------------------------------------
def heading_to_quaternion(heading):
    """``utils.heading_to_quaternion`` (``utils.py:21-35``): quaternion_from_euler(0,0,h)."""
    q = Quaternion()
    q.x = 0.0
    q.y = 0.0
    q.z = math.sin(float(heading) / 2.0)
    q.w = math.cos(float(heading) / 2.0)
    return q


# --------------------------------------------------------------------------------------------
# Feature  (prkt_core_v2.py:881-930)
# --------------------------------------------------------------------------------------------
class Feature(object):
    """Landmark container with the reference's attributes.  Inside a running filter the landmark
    state lives on the device; ``Feature`` objects are what goes in (presets) and what comes
    out (``particles[i].feature_set``)."""

    def __init__(self, mean=None, covar=None):
        self.__immutable__ = False
        if mean is None:
            mean = Matrix([0, 0, 0, 0, 0])
        if covar is None:
            covar = Matrix(np.identity(5, dtype=int))
        self.mean = mean
        self.covar = covar
        self.identity = np.identity(covar.shape[0])
        self.update_count = 0

    # The two update methods are part of the reference's public surface (:897-930).  A filter
    # never calls them -- the fused kernel does this arithmetic -- they act on this host object.
    def update_mean(self, kalman_gain, measure, expected_measure):
        if self.__immutable__:
            return None
        delz = _blob_to_matrix(measure) - _blob_to_matrix(expected_measure)
        self.mean = self.mean + np.dot(kalman_gain, delz)
        self.update_count += 1

    def update_covar(self, kalman_gain, bigH):
        if self.__immutable__:
            return None
        adjust = np.subtract(self.identity, np.dot(kalman_gain, bigH))
        self.covar = np.dot(adjust, self.covar)
        self.update_count += 1


def _blob_to_matrix(blob):
    """``matrix.blob_to_matrix`` (``matrix.py:35-39``)."""
    if hasattr(blob, "bearing"):
        return np.array([blob.bearing, blob.color.r, blob.color.g, blob.color.b])
    return blob


def _feature_arrays(features, capacity, first_id=1):
    """Features -> fp64 SoA arrays (mean5, covp, covc, meta, ids) of length ``capacity``."""
    n = len(features)
    mean5 = np.zeros((capacity, 5))
    covp = np.zeros((capacity, 4))
    covc = np.zeros((capacity, 9))
    meta = np.zeros(capacity, dtype=np.int32)
    ids = np.zeros(capacity, dtype=np.int32)
    for j, f in enumerate(features):
        mean = np.asarray(f.mean, dtype=np.float64).reshape(-1)
        cov = np.asarray(f.covar, dtype=np.float64)
        if mean.shape != (5,) or cov.shape != (5, 5):
            raise ValueError("Feature %d: mean must have 5 entries and covar must be 5x5" % j)
        if np.any(cov[:2, 2:] != 0.0) or np.any(cov[2:, :2] != 0.0):
            raise ValueError(
                "Feature %d: the device stores the landmark covariance as its position (2x2) and "
                "colour (3x3) blocks; non-zero cross terms are not representable (the reference "
                "never produces them, prkt_core_v2.py:799-802)" % j)
        mean5[j] = mean
        covp[j] = cov[:2, :2].reshape(4)
        covc[j] = cov[2:, 2:].reshape(9)
        count = int(getattr(f, "update_count", 0)) & _lib.PK_META_COUNT_MASK
        meta[j] = count | (_lib.PK_META_IMMUTABLE if getattr(f, "__immutable__", False) else 0)
        ids[j] = first_id + j
    return n, mean5, covp, covc, meta, ids


def _feature_from_arrays(mean5, covp, covc, meta):
    cov = np.zeros((5, 5))
    cov[:2, :2] = covp.reshape(2, 2)
    cov[2:, 2:] = covc.reshape(3, 3)
    f = Feature(mean=np.array(mean5, dtype=np.float64), covar=cov)
    f.update_count = int(meta) & _lib.PK_META_COUNT_MASK
    f.__immutable__ = bool(int(meta) & _lib.PK_META_IMMUTABLE)
    return f


# --------------------------------------------------------------------------------------------
# FilterParticle  (prkt_core_v2.py:278-879) -- host-side view / container
# --------------------------------------------------------------------------------------------
class FilterParticle(ParticleMathMixin):
    """One particle as the reference exposes it: ``state`` (Odometry), ``weight``,
    ``feature_set`` {id>0: Feature}, ``potential_features`` {id<0: Feature},
    ``hypothesis_set``, ``next_id`` (``:279-292``).  Objects returned by
    ``FastSLAM.particles[i]`` are snapshots copied from the device.  The scalar helper methods
    of ``:457-877`` (``prob_position_match`` ... ``generate_measurement``) come from
    ``particle_math.ParticleMathMixin``; ``probability_of_match`` / ``match_one`` run on the device."""

    def __init__(self, state=None):
        if state is None:
            state = Odometry()
            state.pose.pose.position.x = 0.0
            state.pose.pose.position.y = 0.0
            state.pose.pose.orientation = heading_to_quaternion(0.0)
        self.state = state
        self.feature_set = {}
        self.potential_features = {}
        self.weight = 1
        self.hypothesis_set = {}
        self.next_id = 1

    def load_feature_list(self, features):
        """``:294-299``"""
        for feature in features:
            self.feature_set[self.next_id] = feature
            self.next_id += 1

    def get_feature_by_id(self, id_):
        """``:301-315`` -- raises KeyError for an unknown id."""
        if id_ < 0:
            return self.potential_features[int(id_)]
        return self.feature_set[id_]

    def no_match_weight(self):
        """``:851-857``"""
        return 0.1

    @property
    def heading(self):
        q = self.state.pose.pose.orientation
        n = q.z * q.z + q.w * q.w
        if n < np.finfo(float).eps * 4.0:
            return 0.0
        return math.atan2(2.0 * q.z * q.w / n, 1.0 - 2.0 * q.z * q.z / n)

    # -- scalar helpers of the reference, evaluated on the device through the probe ABI ----------
    def probability_of_match(self, state, blob, feature):
        """``:383-455`` for one (state, blob, feature) triple, computed by the same device
        function the fused kernel uses (``pk_probe_likelihood``)."""
        from .probe import probability_of_match
        return probability_of_match(state, blob, feature)

    def match_one(self, state, blob):
        """``:353-381``: arg-max over ``feature_set`` then ``potential_features`` in insertion
        order, strict ``>`` from 0.0."""
        from .probe import probability_of_match_many
        features = list(self.feature_set.items()) + list(self.potential_features.items())
        if not features:
            return 0
        vals = probability_of_match_many(state, blob, [f for _, f in features])
        max_match, max_match_id = 0.0, 0
        for (id_, _), v in zip(features, vals):
            if v > max_match:
                max_match, max_match_id = v, id_
        return max_match_id

    def match_features_to_scan(self, scan):
        """``:317-351``"""
        return [(self.match_one(self.state, blob), blob) for blob in scan.observes]

    def add_orphaned_reading(self, state, blob):
        """``:740-746``"""
        self.hypothesis_set[self.next_id] = ((state, blob,))
        self.next_id += 1


class ParticleList(list):
    """``FastSLAM.particles``: a read-only sequence whose items are fetched from the device on
    access (copying a million Python objects per frame is what the reference spends 76 % of its
    time on; the device keeps particles as structure-of-arrays instead).  A ``list`` subclass because
    the reference's attribute is one (``test_prkt_ros2.py:43``); the list storage itself stays empty."""

    def __init__(self, owner):
        super(ParticleList, self).__init__()
        self._owner = owner

    def __bool__(self):
        return len(self) > 0

    def __repr__(self):
        return "<ParticleList of %d particles on %s>" % (len(self), self._owner._device)

    def __len__(self):
        return self._owner.num_particles

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("particle index out of range")
        return self._owner._particle_view(i)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


# --------------------------------------------------------------------------------------------
# FastSLAM  (prkt_core_v2.py:37-276)
# --------------------------------------------------------------------------------------------
class FastSLAM(object):
    """``FastSLAM(preset_features=[])`` as in the reference (``:38``).  Keyword-only extras:

    ``num_particles`` (reference hard-codes 50, ``:41``), ``capacity`` (landmark slots per
    particle, default ``len(preset_features)``), ``dtype`` ``"f64"`` (parity, default) or
    ``"f32"`` (landmark storage for throughput), ``noise`` ``"numpy"`` (three normals per particle
    from NumPy's global legacy stream, exactly what ``motion_model`` consumes ``:185-193``),
    ``"philox"`` (on-device counter RNG) or a callable ``f(M) -> [M,3]`` standard normals;
    ``uniform`` callable for the resampling draw (default ``random.random`` as ``:226``);
    ``clock`` callable returning a ROS-like time (default ``rospy.Time.now``).
    ``spawn=True`` enables the new-landmark path the reference's docstrings describe but its code never
    reaches (SURVEY.md finding F5 / A.6: ``add_hypothesis`` ``:546-746`` with three documented patches):
    unseen blobs are paired with earlier orphaned readings whose rays intersect and whose colours lie
    within ``pair_gate`` (default ``sqrt(300)``), the pair is triangulated into a potential landmark
    (id < 0) that is promoted after three updates.  ``orphan_capacity`` readings are kept per particle
    (a ring; the reference keeps them for ever).  Needs ``capacity`` > number of preset landmarks.
    ``measurement_model="textbook"``: the EKF update with the textbook bearing model (robot-frame predicted bearing,
    Jacobian row ``[-dy/q, +dx/q]``, wrapped innovation) instead of the reference's as-written one (``:785-797, 871``,
    SURVEY.md finding F4 a/b/e) -- ``PK_MODEL_TEXTBOOK``; a documented deviation, default ``"reference"``.
    ``weights="log"``: importance factors and particle weights are carried as logarithms (``PK_MODEL_LOG_WEIGHTS``) and
    turned into linear weights ``exp(lw - max)`` by a log-sum-exp normaliser (warp-shuffle reductions; across shards the
    maximum and the sums are NCCL all-reduces) right before the resampling scan.  Unlike the reference's fp64 product
    (``:124``) a frame of very unlikely observations cannot zero every weight.  Default ``"linear"`` = the reference.
    ``publish_particles=n`` (> 0): publish the pose of a bounded, evenly strided sub-sample of at most ``n`` particles per
    frame on the reference's three debugging topics ``/particle_track``, ``/aged_particles`` and ``/resampled_particles``
    (``:55-57, 127, 237, 242``; the reference publishes EVERY particle, which at 10^6 particles would be the whole
    frame time).  Default 0: nothing is published and no pose leaves the device.
    ``arithmetic="f32"`` (with ``dtype="f32"`` only): the landmark algebra of the fused kernel -- gates, Mahalanobis
    forms, EKF gain and covariance update -- runs in fp32 on the fp32 records; poses, the importance weights and the
    whole resampling stay fp64 (``PK_DTYPE_ARITH_F32``).  The throughput mode: >= 90 % of the reference's indices.
    """

    def __init__(self, preset_features=[], *, num_particles=50, capacity=None, dtype="f64",
                 device=None, noise="numpy", seed=0, uniform=None, clock=None, params=None,
                 spawn=False, orphan_capacity=32, pair_gate=300.0 ** 0.5, arithmetic="f64", publish_particles=0,
                 measurement_model="reference", weights="linear", overlap_copy=True):
        import torch

        _lib.require_device()
        self._torch = torch
        self._lib = _lib.load()
        self._lock = threading.RLock()
        self._overlap_copy = bool(overlap_copy)
        self._clock = clock if clock is not None else _ros_now
        self._device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())

        self.last_control = Twist()                    # :39
        self.last_update = self._clock()               # :40
        self.num_particles = int(num_particles)        # :41
        self.Qt = Matrix([[.1, 0, 0, 0], [0, .1, 0, 0], [0, 0, .1, 0], [0, 0, 0, .1]])  # :50-53
        self.params = params if params is not None else _lib.default_params()
        if measurement_model not in ("reference", "textbook"):
            raise ValueError("measurement_model must be 'reference' or 'textbook'")
        self.measurement_model = measurement_model
        if measurement_model == "textbook":
            self.params.model |= _lib.PK_MODEL_TEXTBOOK
        if weights not in ("linear", "log"):
            raise ValueError("weights must be 'linear' or 'log'")
        self.weights = weights
        if weights == "log":
            self.params.model |= _lib.PK_MODEL_LOG_WEIGHTS

        if dtype not in ("f32", "f64"):
            raise ValueError("dtype must be 'f32' or 'f64'")
        self.dtype = dtype
        self._dt = _lib.PK_DTYPE_F64 if dtype == "f64" else _lib.PK_DTYPE_F32
        if arithmetic not in ("f64", "f32"):
            raise ValueError("arithmetic must be 'f64' or 'f32'")
        if arithmetic == "f32" and dtype != "f32":
            raise ValueError("arithmetic='f32' needs dtype='f32' (fp32 landmark storage)")
        self.arithmetic = arithmetic
        if arithmetic == "f32":
            self._dt |= _lib.PK_DTYPE_ARITH_F32
        self.spawn = bool(spawn)
        self.orphan_capacity = int(orphan_capacity) if self.spawn else 0
        self.pair_gate = float(pair_gate)
        if self.spawn:
            if not 0 < self.orphan_capacity <= _lib.PK_MAX_ORPHANS:
                raise ValueError("orphan_capacity must be in [1, %d]" % _lib.PK_MAX_ORPHANS)
            # layout code: storage type | orphan slots (the orphan region lives inside each particle's block)
            self._dt = _lib.dtype_with_orphans(self._dt, self.orphan_capacity)
        self._noise = noise
        if not (noise in ("numpy", "philox") or callable(noise)):
            raise ValueError("noise must be 'numpy', 'philox' or a callable")
        self._seed = int(seed)
        self._frame = 0
        self._uniform = uniform if uniform is not None else _pyrandom.random
        self.particle_offset = 0       # global index of local particle 0 (sharded filters)

        preset_features = list(preset_features)
        n = len(preset_features)
        self.capacity = int(capacity) if capacity is not None else n
        if self.capacity < n:
            raise ValueError("capacity smaller than the preset map")
        self._alloc()
        self._load_presets(preset_features)
        self.particles = ParticleList(self)           # :42-49
        self.publish_particles = int(publish_particles)
        self.aged_particles_pub = self.resampled_particles_pub = self.particle_track_pub = None
        if self.publish_particles > 0:                # :55-57
            self.aged_particles_pub = Publisher('/aged_particles', Odometry, queue_size=1)
            self.resampled_particles_pub = Publisher('/resampled_particles', Odometry, queue_size=1)
            self.particle_track_pub = Publisher('/particle_track', Odometry, queue_size=1)
        self.last_stats = {}
        self.last_assoc = None
        self.last_ancestors = None
        self.keep_trace = False

    # -- allocation ------------------------------------------------------------------------------
    def _alloc(self):
        torch, M, dev = self._torch, self.num_particles, self._device
        lib = self._lib
        self.block_bytes = int(lib.pk_block_bytes(self.capacity, self._dt))
        f64, i32, i64 = torch.float64, torch.int32, torch.int64
        self._pose = [torch.zeros((M, 4), dtype=f64, device=dev) for _ in range(2)]
        self._aux = [torch.zeros((M, 2), dtype=i32, device=dev) for _ in range(2)]
        self._slot = [torch.zeros((M,), dtype=i32, device=dev) for _ in range(2)]
        self._cur = 0
        self._pool_t = torch.zeros((max(1, M * self.block_bytes),), dtype=torch.uint8, device=dev)
        # copy-on-resample runs on its own stream beside the next frame's motion update (see low_variance_resample)
        self._copy_stream = torch.cuda.Stream(device=dev) if self._overlap_copy else None
        self._ev_permuted = torch.cuda.Event()
        self._ev_blocks = torch.cuda.Event()
        self._blocks_pending = False
        nb = int(lib.pk_num_scan_blocks(max(M, 1)))
        self._nb = nb
        self._cumsum = torch.zeros((max(M, 1),), dtype=f64, device=dev)
        self._block_sums = torch.zeros((nb,), dtype=f64, device=dev)
        self._plan = torch.zeros((_lib.PK_PLAN_DOUBLES,), dtype=f64, device=dev)
        self._block_prefix = torch.zeros((nb, 2), dtype=f64, device=dev)
        self._block_count = torch.zeros((nb + 1,), dtype=i64, device=dev)
        self._out_lo = torch.zeros((max(M, 1),), dtype=i64, device=dev)
        self._offspring = torch.zeros((max(M, 1),), dtype=i32, device=dev)
        self._ancestors = torch.zeros((max(M, 1),), dtype=i64, device=dev)
        self._big_runs = torch.zeros((4 + 3 * (M // 16 + 2),), dtype=i64, device=dev)
        self._gather_ws = torch.zeros((max(256, int(lib.pk_gather_workspace_bytes(max(M, 1)))),),
                                      dtype=torch.uint8, device=dev)
        self._n_copied = torch.zeros((1,), dtype=i64, device=dev)
        self._stats = torch.zeros((_lib.PK_NUM_STATS,), dtype=i64, device=dev)
        self._red_ws = torch.zeros((5 * 1024,), dtype=f64, device=dev)
        self._out5 = torch.zeros((5,), dtype=f64, device=dev)
        self._best2 = torch.zeros((2,), dtype=f64, device=dev)
        self._wmax = torch.zeros((1,), dtype=f64, device=dev)     # log-weight normaliser: maximum log weight
        self._wstats = torch.zeros((3,), dtype=f64, device=dev)   # ... sum w, sum w^2, max (after normalisation)
        self._pinned = {}
        self._assoc = None
        self._obs_table = None
        self._noise_pinned = None
        self._noise_dev = None

    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self._device).cuda_stream)

    def _on_device(self):
        """Context that makes the filter's device current (free when it already is)."""
        torch = self._torch
        if torch.cuda.current_device() == self._device.index:
            return _NULL_CTX
        return torch.cuda.device(self._device)

    @property
    def _pool(self):
        """The landmark pool.  Every access first orders the current stream behind block copies of the last resampling
        that may still be running on the copy stream, so no reader or writer of the pool can overtake them."""
        self.wait_blocks()
        return self._pool_t

    def __del__(self):
        # block copies still running on the copy stream must not outlive the pool they write into
        try:
            if getattr(self, "_blocks_pending", False):
                self._ev_blocks.synchronize()
        except Exception:
            pass

    def wait_blocks(self):
        """Make the current stream wait for the block copies of the last resampling (no-op when none are pending)."""
        if self._blocks_pending:
            self._torch.cuda.current_stream(self._device).wait_event(self._ev_blocks)
            self._blocks_pending = False

    @property
    def pose(self):
        """Device tensor [M,4]: x, y, heading, weight (current buffer)."""
        return self._pose[self._cur]

    @property
    def aux(self):
        return self._aux[self._cur]

    @property
    def slot(self):
        return self._slot[self._cur]

    def _load_presets(self, features):
        torch, lib, M = self._torch, self._lib, self.num_particles
        n, mean5, covp, covc, meta, ids = _feature_arrays(features, max(self.capacity, 1))
        with self._on_device():
            _lib.check(lib.pk_init_particles(_lib.ptr(self.pose), _lib.ptr(self.slot), _lib.ptr(self.aux), M, n,
                                             1 + n, self._stream()), "pk_init_particles")
            if n and M:
                d = [torch.from_numpy(a).to(self._device) for a in (mean5, covp, covc, meta, ids)]
                _lib.check(lib.pk_map_broadcast(_lib.ptr(self._pool), self.capacity, self._dt, 0, M, n,
                                                *[_lib.ptr(t) for t in d], self._stream()), "pk_map_broadcast")
                torch.cuda.current_stream(self._device).synchronize()

    # -- frame driver: cam_cb (:59-137) ------------------------------------------------------------
    def cam_cb(self, ros_view):
        """One measurement frame: motion update with the last control (``:75-77``), association
        of all blobs then the sequential EKF updates and weight product (``:84-124``), then the
        low-variance resample (``:137``)."""
        with self._lock:
            if self.num_particles == 0:
                return
            # (the scan is unpacked first so that the motion and measurement kernels are issued back to back)
            scan = ros_view.last_sensor_reading
            try:
                obs = self._scan_to_array(scan)
            except AttributeError:
                self.motion_update(self.last_control)      # the reference moves the particles before it trips (:75-82)
                raise
            self.motion_update(self.last_control)
            self.measurement_update(obs)
            if self.particle_track_pub is not None:
                self._publish_sample(self.particle_track_pub)          # :126-127
            if self.aged_particles_pub is not None:
                self._publish_sample(self.aged_particles_pub)          # :241-242
            self.low_variance_resample()
            if self.resampled_particles_pub is not None:
                self._publish_sample(self.resampled_particles_pub)     # :236-237 (the copies that were emitted)

    def _publish_sample(self, pub):
        """Publish the current pose of at most ``publish_particles`` evenly strided particles as Odometry messages
        with ``frame_id = 'odom'`` (``:126, 236, 241``): one small device-to-host copy."""
        torch, M = self._torch, self.num_particles
        n = min(self.publish_particles, M)
        if n <= 0:
            return
        idx = torch.linspace(0, M - 1, n, device=self._device).round().to(torch.int64)
        rows = self.pose.index_select(0, idx).cpu().numpy()
        for x, y, heading, _w in rows:
            msg = Odometry()
            msg.header.frame_id = 'odom'
            msg.pose.pose.position.x = float(x)
            msg.pose.pose.position.y = float(y)
            msg.pose.pose.orientation = heading_to_quaternion(float(heading))
            pub.publish(msg)

    @staticmethod
    def _scan_to_array(scan):
        """K x (bearing, r, g, b) from ``scan.observes`` (``:344``, ``matrix.py:35-39``)."""
        if scan is None:
            raise AttributeError("'NoneType' object has no attribute 'observes'")
        if isinstance(scan, np.ndarray):
            return np.ascontiguousarray(scan, dtype=np.float64).reshape(-1, 4)
        if hasattr(scan, "is_cuda"):
            return scan            # device-resident scan (BearingSimulator): handed to the kernels as is
        blobs = scan.observes
        obs = np.empty((len(blobs), 4), dtype=np.float64)
        for k, b in enumerate(blobs):
            obs[k, 0] = b.bearing
            obs[k, 1] = b.color.r
            obs[k, 2] = b.color.g
            obs[k, 3] = b.color.b
        return obs

    def measurement_update(self, obs):
        """The per-particle body of ``cam_cb`` (``:73, :84-124``) for a ``[K,4]`` array of blobs, as ONE fused
        kernel.  ``obs`` is a host array (the blob table rides in the kernel arguments) or a CUDA tensor, e.g.
        a ``BearingSimulator`` scan (the table is then built on the device and nothing visits the host).  Also
        the v1 core's name for the same step (``prkt_core.py:159-236``)."""
        torch, lib, M = self._torch, self._lib, self.num_particles
        on_device = hasattr(obs, "is_cuda") and obs.is_cuda
        if on_device:
            if obs.dtype != torch.float64 or obs.dim() != 2 or obs.shape[1] != 4 or not obs.is_contiguous():
                raise ValueError("a device scan must be a contiguous float64 [K, 4] tensor")
        else:
            obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 4)
        K = int(obs.shape[0])
        if K > _lib.PK_MAX_OBS:
            raise ValueError("at most %d blobs per frame (got %d)" % (_lib.PK_MAX_OBS, K))
        with self._lock, self._on_device():
            if self._assoc is None or self._assoc.shape[1] != max(K, 1):
                self._assoc = torch.zeros((M, max(K, 1)), dtype=torch.int32, device=self._device)
            if on_device:
                if self._obs_table is None:
                    self._obs_table = torch.zeros((int(lib.pk_obs_table_bytes()),), dtype=torch.uint8,
                                                  device=self._device)
                _lib.check(lib.pk_measurement_update_dev(
                    _lib.ptr(self.pose), _lib.ptr(self.aux), _lib.ptr(self.slot), _lib.ptr(self._pool),
                    self.capacity, self._dt, M, _lib.ptr(obs), K, ctypes.byref(self.params),
                    _lib.ptr(self._assoc), _lib.ptr(self._stats), _lib.ptr(self._obs_table), self._stream()),
                    "pk_measurement_update_dev")
            else:
                _lib.check(lib.pk_measurement_update(
                    _lib.ptr(self.pose), _lib.ptr(self.aux), _lib.ptr(self.slot), _lib.ptr(self._pool),
                    self.capacity, self._dt, M, obs.ctypes.data, K, ctypes.byref(self.params),
                    _lib.ptr(self._assoc), _lib.ptr(self._stats), self._stream()), "pk_measurement_update")
            if self.spawn and K:
                # add_hypothesis for every unseen blob (:92-94), after the frame's associations were made
                fn = lib.pk_spawn_update_dev if on_device else lib.pk_spawn_update
                _lib.check(fn(_lib.ptr(self.pose), _lib.ptr(self.aux), _lib.ptr(self.slot), _lib.ptr(self._pool),
                              self.capacity, self._dt, M, _lib.ptr(obs) if on_device else obs.ctypes.data, K,
                              _lib.ptr(self._assoc), self.pair_gate, _lib.ptr(self._stats), self._stream()),
                           "pk_spawn_update")
            self._last_K = K
            if self.keep_trace:
                self.last_assoc = self._assoc[:, :K].clone()
                self.last_weight = self.pose[:, 3].clone()
                self.last_pose_pre = self.pose[:, :3].clone()

    cam_observation_update = measurement_update       # v1 lineage name (prkt_core.py:205)

    def odom_motion_update(self, odom):
        """``:140-146`` -- "***Alpha feature***": the reference's body is ``pass``; kept for interface parity."""
        return None

    # -- motion: motion_update / motion_model (:148-208) -------------------------------------------
    def motion_update(self, new_twist):
        """``:148-166``: integrate the PREVIOUS control over ``now - last_update`` for every
        particle, then store the new control."""
        with self._lock:
            dt = self._clock() - self.last_update                     # :158
            self._motion_all(self.last_control, dt.to_sec())          # :159-163
            self.last_update = self.last_update + dt                  # :165
            self.last_control = new_twist                             # :166

    def _draw_noise(self, M):
        if self._noise == "philox":
            return None
        if self._noise == "numpy":
            z = np.random.standard_normal((M, 3))
        else:
            z = np.ascontiguousarray(self._noise(M), dtype=np.float64)
            if z.shape != (M, 3):
                raise ValueError("noise callable must return an [M,3] array")
        return z

    def _motion_all(self, twist, dt):
        torch, lib, M = self._torch, self._lib, self.num_particles
        if M == 0:
            return
        v = float(twist.linear.x)                                     # :176
        w = float(twist.angular.z)                                    # :177
        with self._on_device():
            z = self._draw_noise(M)
            nptr = 0
            if z is not None:
                if self._noise_pinned is None:
                    self._noise_pinned = torch.empty((M, 3), dtype=torch.float64, pin_memory=True)
                    self._noise_dev = torch.empty((M, 3), dtype=torch.float64, device=self._device)
                self._noise_pinned.numpy()[...] = z
                self._noise_dev.copy_(self._noise_pinned, non_blocking=True)
                nptr = _lib.ptr(self._noise_dev)
            _lib.check(lib.pk_motion_update(_lib.ptr(self.pose), M, nptr, self._seed, self._frame,
                                            self.particle_offset, v, w, float(dt), self._stream()),
                       "pk_motion_update")
            self._frame += 1
            if z is not None:
                # the pinned staging buffer is reused next frame
                torch.cuda.current_stream(self._device).synchronize()

    def motion_model(self, particle, twist, dt):
        """``:168-208`` for ONE host-side particle (the form ``test_prkt_ros2.py:53`` calls):
        returns a new ``FilterParticle``; the arithmetic runs in the motion kernel."""
        import copy
        torch, lib = self._torch, self._lib
        dt = dt.to_sec()                                              # :174
        new_particle = copy.deepcopy(particle)                        # :181
        pos = particle.state.pose.pose.position
        heading = particle.heading if isinstance(particle, FilterParticle) else FilterParticle.heading.fget(particle)
        with self._lock, self._on_device():
            rec = torch.tensor([[float(pos.x), float(pos.y), float(heading), 1.0]], dtype=torch.float64,
                               device=self._device)
            z = self._draw_noise(1)
            zt = None if z is None else torch.from_numpy(z).to(self._device)
            _lib.check(lib.pk_motion_update(_lib.ptr(rec), 1, _lib.ptr(zt), self._seed, self._frame, 0,
                                            float(twist.linear.x), float(twist.angular.z), float(dt),
                                            self._stream()), "pk_motion_update")
            out = rec.cpu().numpy()[0]
        new_particle.state.pose.pose.position.x = float(out[0])
        new_particle.state.pose.pose.position.y = float(out[1])
        new_particle.state.pose.pose.orientation = heading_to_quaternion(float(out[2]))
        return new_particle

    # -- resampling: low_variance_resample (:210-252) ----------------------------------------------
    def low_variance_resample(self):
        """Systematic resampling, every frame, one uniform draw (``:226``)."""
        torch, lib, M = self._torch, self._lib, self.num_particles
        if M == 0:
            return
        with self._lock, self._on_device():
            u01 = float(self._uniform())
            st = self._stream()
            cur, nxt = self._cur, 1 - self._cur
            self.wait_blocks()  # two resamplings in a row: the copy lists in the workspace are still being read
            if self.weights == "log":
                self._normalise_log_weights()
            _lib.check(lib.pk_weight_scan(_lib.ptr(self._pose[cur]), M, _lib.ptr(self._cumsum),
                                          _lib.ptr(self._block_sums), st), "pk_weight_scan")
            _lib.check(lib.pk_resample_thresholds(_lib.ptr(self._block_sums), self._nb, M, u01,
                                                  _lib.ptr(self._plan), _lib.ptr(self._block_prefix),
                                                  _lib.ptr(self._block_count), st), "pk_resample_thresholds")
            # K4 + dead-particle scan in one kernel, then free list / permutation / block copies (6 launches in all)
            _lib.check(lib.pk_resample_plan(_lib.ptr(self._cumsum), M, 0, 0, _lib.ptr(self._plan),
                                            _lib.ptr(self._block_prefix), _lib.ptr(self._block_count), M, 0, M,
                                            _lib.ptr(self._out_lo), _lib.ptr(self._offspring),
                                            _lib.ptr(self._ancestors), _lib.ptr(self._gather_ws), st),
                       "pk_resample_plan")
            overlap = self._copy_stream is not None and self.capacity > 0
            _lib.check(lib.pk_resample_gather_planned(_lib.ptr(self._ancestors), M,
                                                      _lib.ptr(self._pose[cur]), _lib.ptr(self._pose[nxt]),
                                                      _lib.ptr(self._aux[cur]), _lib.ptr(self._aux[nxt]),
                                                      _lib.ptr(self._slot[cur]), _lib.ptr(self._slot[nxt]),
                                                      _lib.ptr(self._pool_t), 0 if overlap else self.capacity, self._dt,
                                                      _lib.ptr(self._gather_ws), _lib.ptr(self._n_copied), st),
                       "pk_resample_gather_planned")
            if overlap:
                # The block copies (deepcopy :243) touch nothing but the landmark pool: they go to the copy stream,
                # ordered behind the permutation, and the next frame's motion update (poses only, bound by its fp64
                # arithmetic) runs beside them (bound by HBM).  Whatever touches the pool next waits: `_pool`.
                main = torch.cuda.current_stream(self._device)
                self._ev_permuted.record(main)
                self._copy_stream.wait_event(self._ev_permuted)
                _lib.check(lib.pk_resample_copy_blocks(_lib.ptr(self._pool_t), self.capacity, self._dt, M,
                                                       _lib.ptr(self._gather_ws), _lib.ptr(self._n_copied),
                                                       ctypes.c_void_p(self._copy_stream.cuda_stream)),
                           "pk_resample_copy_blocks")
                self._ev_blocks.record(self._copy_stream)
                self._blocks_pending = True
            self._cur = nxt
            if self.keep_trace:
                self.last_ancestors = self._ancestors.clone()

    # -- log-domain weights (weights="log") ----------------------------------------------------------------
    def _all_reduce_weight_stat(self, tensor, op):
        """Cross-shard reduction hook of the weight normaliser (no-op on one GPU; NCCL all-reduce when sharded)."""
        return None

    def _normalise_log_weights(self):
        """log weights -> linear weights exp(lw - max) in place, with sum w and sum w^2 on the side (log-sum-exp
        normaliser: max reduction, all-reduce(max), exp + sums, all-reduce(sum))."""
        lib, M, st = self._lib, self.num_particles, self._stream()
        _lib.check(lib.pk_log_weights_max(_lib.ptr(self.pose), M, _lib.ptr(self._wmax), _lib.ptr(self._red_ws), st),
                   "pk_log_weights_max")
        self._all_reduce_weight_stat(self._wmax, "max")
        _lib.check(lib.pk_log_weights_normalise(_lib.ptr(self.pose), M, _lib.ptr(self._wmax), _lib.ptr(self._wstats),
                                                _lib.ptr(self._red_ws), st), "pk_log_weights_normalise")
        self._all_reduce_weight_stat(self._wstats[:2], "sum")

    def effective_sample_size(self):
        """N_eff = (sum w)^2 / sum w^2 of the weights of the last measurement update.  ``weights="log"``: from the
        normaliser of the last resample (also returns the log-sum-exp of the log weights); linear weights: reduced on
        demand from the current weights (call it before ``low_variance_resample`` permutes them)."""
        torch, lib, M = self._torch, self._lib, self.num_particles
        with self._lock, self._on_device():
            if self.weights == "log":
                s = self._wstats.cpu().numpy()
                return float(s[0] * s[0] / s[1]), float(s[2] + math.log(s[0]))
            w = self.pose[:, 3]
            t = torch.stack([w.sum(), (w * w).sum()])
            self._all_reduce_weight_stat(t, "sum")
            s = t.cpu().numpy()
        return float(s[0] * s[0] / s[1]) if s[1] > 0 else 0.0, float(math.log(s[0])) if s[0] > 0 else float("-inf")

    # -- queries -------------------------------------------------------------------------------------
    def summary(self):
        """``:254-276``: unweighted mean x, mean y and circular-mean heading."""
        torch, lib, M = self._torch, self._lib, self.num_particles
        with self._lock, self._on_device():
            _lib.check(lib.pk_summary_partial(_lib.ptr(self.pose), M, _lib.ptr(self._out5), _lib.ptr(self._red_ws),
                                              self._stream()), "pk_summary_partial")
            s = self._read_back(self._out5)
        count = float(M)
        return (float(s[0] / count), float(s[1] / count), math.atan2(float(s[2]), float(s[3])),)

    def _read_back(self, dev_tensor):
        """Small device tensor -> NumPy through a pinned staging buffer (asynchronous copy + one stream
        synchronisation; `.cpu()` would allocate pageable memory and take the slow synchronous-copy path)."""
        torch = self._torch
        key = (dev_tensor.dtype, tuple(dev_tensor.shape))
        host = self._pinned.get(key)
        if host is None:
            host = self._pinned[key] = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True)
        host.copy_(dev_tensor, non_blocking=True)
        torch.cuda.current_stream(self._device).synchronize()
        return host.numpy().copy()

    def best_particle(self):
        """Additive API: (index, weight) of the first particle with the largest weight (weights
        are those of the last measurement update; after resampling they are the ancestors')."""
        torch, lib, M = self._torch, self._lib, self.num_particles
        with self._lock, self._on_device():
            _lib.check(lib.pk_best_particle(_lib.ptr(self.pose), M, _lib.ptr(self._best2), _lib.ptr(self._red_ws),
                                            self._stream()), "pk_best_particle")
            b = self._best2.cpu().numpy()
        return int(b[1]), float(b[0])

    def stats(self):
        """Counters of the last measurement update (matched / unmatched pairs, exact likelihood
        evaluations, flags) and the number of landmark blocks copied by the last resample."""
        s = self._stats.cpu().numpy()
        return dict(matched=int(s[_lib.PK_STAT_MATCHED]), unmatched=int(s[_lib.PK_STAT_UNMATCHED]),
                    evaluated=int(s[_lib.PK_STAT_EVALUATED]), flags=int(s[_lib.PK_STAT_FLAGS]),
                    same_landmark=int(s[_lib.PK_STAT_SAME_LANDMARK]), promoted=int(s[_lib.PK_STAT_PROMOTED]),
                    spawned=int(s[_lib.PK_STAT_SPAWNED]), orphaned=int(s[_lib.PK_STAT_ORPHANED]),
                    blocks_copied=int(self._n_copied.item()))

    def export_maps(self, lo=0, count=None):
        """Landmark state of particles [lo, lo+count) as fp64 NumPy arrays
        (mean [c,N,5], covp [c,N,2,2], covc [c,N,3,3], meta [c,N], ids [c,N], n_live [c])."""
        torch, lib = self._torch, self._lib
        count = self.num_particles - lo if count is None else count
        N = max(self.capacity, 1)
        dev = self._device
        with self._lock, self._on_device():
            mean5 = torch.zeros((count, N, 5), dtype=torch.float64, device=dev)
            covp = torch.zeros((count, N, 4), dtype=torch.float64, device=dev)
            covc = torch.zeros((count, N, 9), dtype=torch.float64, device=dev)
            meta = torch.zeros((count, N), dtype=torch.int32, device=dev)
            ids = torch.zeros((count, N), dtype=torch.int32, device=dev)
            if self.capacity and count:
                _lib.check(lib.pk_map_export(_lib.ptr(self._pool), self.capacity, self._dt, _lib.ptr(self.slot), lo,
                                             count, _lib.ptr(mean5), _lib.ptr(covp), _lib.ptr(covc), _lib.ptr(meta),
                                             _lib.ptr(ids), self._stream()), "pk_map_export")
            nlive = self.aux[lo:lo + count, 0].cpu().numpy()
            return (mean5.cpu().numpy(), covp.cpu().numpy().reshape(count, N, 2, 2),
                    covc.cpu().numpy().reshape(count, N, 3, 3), meta.cpu().numpy(), ids.cpu().numpy(), nlive)

    def export_orphans(self, lo=0, count=None):
        """Spawn mode: the stored orphan readings of particles [lo, lo+count), oldest first, as a list of
        ``[n_i, 8]`` arrays (x, y, cos(ray), sin(ray), r, g, b, id) and the totals ever stored."""
        torch, lib = self._torch, self._lib
        if not self.spawn:
            raise ValueError("export_orphans needs spawn=True")
        count = self.num_particles - lo if count is None else count
        S = self.orphan_capacity
        with self._lock, self._on_device():
            totals = torch.zeros((max(count, 1),), dtype=torch.int32, device=self._device)
            readings = torch.zeros((max(count, 1), S, 8), dtype=torch.float64, device=self._device)
            if count:
                _lib.check(lib.pk_orphans_export(_lib.ptr(self._pool), self.capacity, self._dt, _lib.ptr(self.slot), lo,
                                                 count, _lib.ptr(totals), _lib.ptr(readings), self._stream()),
                           "pk_orphans_export")
            totals = totals.cpu().numpy()[:count]
            readings = readings.cpu().numpy()[:count]
        out = []
        for i in range(count):
            t = int(totals[i])
            live = min(t, S)
            start = t % S if t > S else 0
            out.append(readings[i][[(start + j) % S for j in range(live)]].copy())
        return out, totals

    def import_maps(self, lo, mean5, covp, covc, meta, ids, n_live=None):
        """Overwrite the landmark state of particles [lo, lo+count) from fp64 arrays shaped like the
        output of ``export_maps`` (``meta`` = update_count | PK_META_IMMUTABLE | PK_META_POTENTIAL, ``ids`` the
        reference ids: > 0 full feature, < 0 potential feature)."""
        torch, lib = self._torch, self._lib
        count = len(mean5)
        dev = self._device
        N = self.capacity
        if not (0 <= lo and lo + count <= self.num_particles):
            raise ValueError("import_maps: particles [%d, %d) are outside the filter" % (lo, lo + count))
        shapes = (np.shape(mean5), np.shape(np.reshape(covp, (count, -1, 4))), np.shape(np.reshape(covc, (count, -1, 9))),
                  np.shape(meta), np.shape(ids))
        if shapes != ((count, N, 5), (count, N, 4), (count, N, 9), (count, N), (count, N)):
            raise ValueError("import_maps: arrays must be shaped [count, capacity=%d, ...] like export_maps' (got %s)"
                             % (N, shapes,))
        if n_live is not None and (np.shape(n_live) != (count,) or np.min(n_live, initial=0) < 0
                                   or np.max(n_live, initial=0) > N):
            raise ValueError("import_maps: n_live must hold count values in [0, capacity]")
        with self._lock, self._on_device():
            t = [torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev) for a, dt in
                 ((mean5, np.float64), (np.reshape(covp, (count, -1, 4)), np.float64),
                  (np.reshape(covc, (count, -1, 9)), np.float64), (meta, np.int32), (ids, np.int32))]
            _lib.check(lib.pk_map_import(_lib.ptr(self._pool), self.capacity, self._dt, _lib.ptr(self.slot), lo, count,
                                         *[_lib.ptr(x) for x in t], self._stream()), "pk_map_import")
            if n_live is not None:
                self.aux[lo:lo + count, 0] = torch.from_numpy(np.asarray(n_live, dtype=np.int32)).to(dev)
            torch.cuda.current_stream(dev).synchronize()

    def get_map(self, i):
        """Additive API: ``{id: Feature}`` of particle ``i`` (full and potential landmarks)."""
        p = self._particle_view(i)
        out = dict(p.feature_set)
        out.update(p.potential_features)
        return out

    def _particle_view(self, i):
        mean5, covp, covc, meta, ids, nlive = self.export_maps(i, 1)
        rec = self.pose[i].cpu().numpy()
        aux = self.aux[i].cpu().numpy()
        p = FilterParticle()
        p.state.pose.pose.position.x = float(rec[0])
        p.state.pose.pose.position.y = float(rec[1])
        p.state.pose.pose.orientation = heading_to_quaternion(float(rec[2]))
        p.weight = float(rec[3])
        p.next_id = int(aux[1])
        for j in range(int(nlive[0])):
            f = _feature_from_arrays(mean5[0, j], covp[0, j], covc[0, j], meta[0, j])
            if int(meta[0, j]) & _lib.PK_META_POTENTIAL:
                p.potential_features[int(ids[0, j])] = f
            else:
                p.feature_set[int(ids[0, j])] = f
        if self.spawn:
            # hypothesis_set {id: (state, blob)} (:291, :745); the device keeps the reading's world-frame
            # ray, so the view reports it as heading = ray angle with a zero bearing (the same ray)
            from .rosless import messages
            readings, _ = self.export_orphans(i, 1)
            for row in readings[0]:
                st = Odometry()
                st.pose.pose.position.x = float(row[0])
                st.pose.pose.position.y = float(row[1])
                st.pose.pose.orientation = heading_to_quaternion(math.atan2(float(row[3]), float(row[2])))
                blob = messages.Blob()
                blob.bearing = 0.0
                blob.color.r, blob.color.g, blob.color.b = float(row[4]), float(row[5]), float(row[6])
                p.hypothesis_set[int(row[7])] = (st, blob)
        return p

    def state_dict(self):
        """Checkpoint of the device state (host tensors) plus what the next frame needs from the host side: the control
        and time of the last motion update (``:162-166``), the Philox frame counter and the layout it was saved with."""
        with self._lock, self._on_device():
            self._torch.cuda.current_stream(self._device).synchronize()
            lu = self.last_update
            return dict(pose=self.pose.cpu(), aux=self.aux.cpu(), slot=self.slot.cpu(), pool=self._pool.cpu(),
                        frame=self._frame, capacity=self.capacity, dtype=self.dtype, layout=int(self._dt),
                        arithmetic=self.arithmetic, num_particles=self.num_particles, block_bytes=self.block_bytes,
                        spawn=self.spawn, orphan_capacity=self.orphan_capacity, pair_gate=self.pair_gate,
                        particle_offset=self.particle_offset,
                        last_control=(float(self.last_control.linear.x), float(self.last_control.angular.z)),
                        last_update=(int(getattr(lu, "secs", 0)), int(getattr(lu, "nsecs", 0))))

    def load_state_dict(self, sd):
        """Restore ``state_dict()``.  The filter must have been built with the same layout (particles, capacity,
        storage type, arithmetic, spawn mode / orphan slots): anything else is refused, not reinterpreted."""
        mine = dict(capacity=self.capacity, dtype=self.dtype, layout=int(self._dt), arithmetic=self.arithmetic,
                    num_particles=self.num_particles, block_bytes=self.block_bytes, spawn=self.spawn,
                    orphan_capacity=self.orphan_capacity)
        for k, v in mine.items():
            if k in sd and sd[k] != v:
                raise ValueError("checkpoint layout mismatch: %s is %r here, %r in the checkpoint" % (k, v, sd[k]))
        if sd["capacity"] != self.capacity or sd["dtype"] != self.dtype:
            raise ValueError("checkpoint layout mismatch")
        for name, t in (("pose", self.pose), ("aux", self.aux), ("slot", self.slot), ("pool", self._pool)):
            if tuple(sd[name].shape) != tuple(t.shape) or sd[name].dtype != t.dtype:
                raise ValueError("checkpoint tensor %r has shape %s / %s, expected %s / %s"
                                 % (name, tuple(sd[name].shape), sd[name].dtype, tuple(t.shape), t.dtype))
        with self._lock, self._on_device():
            self.pose.copy_(sd["pose"])
            self.aux.copy_(sd["aux"])
            self.slot.copy_(sd["slot"])
            self._pool.copy_(sd["pool"])
            self._frame = int(sd["frame"])
            self.particle_offset = int(sd.get("particle_offset", self.particle_offset))
            if "last_control" in sd:
                tw = Twist()
                tw.linear.x, tw.angular.z = sd["last_control"]
                self.last_control = tw
            if "last_update" in sd and hasattr(self.last_update, "secs"):
                self.last_update = type(self.last_update)(*sd["last_update"])
            # the host tensors may be reused or freed by the caller as soon as this returns
            self._torch.cuda.current_stream(self._device).synchronize()
