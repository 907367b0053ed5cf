"""Host-side scalar forms of the ``FilterParticle`` helper methods the reference exposes
(``/root/reference/src/prkt_core_v2.py:457-877``) and its unit tests call directly
(``test_prkt_ros2.py:126-423``).  SURVEY.md section 8(b): "helper methods ... for test compatibility (host-side
scalar versions are fine)".

These act on ONE host object (a ``FilterParticle`` snapshot, a ``Feature``, a ``Blob``): they are the
interface surface, not the hot path.  A running filter never calls them -- the same arithmetic for a
million particles is the fused measurement kernel (``csrc/pk_measure.cu``) and the spawn kernel
(``csrc/pk_spawn.cu``), which the parity tests pin against the reference directly.

Third-party arithmetic restated from its published algorithm (absent from ``/root/reference``):
``scipy.stats.multivariate_normal.pdf`` (eigendecomposition of the covariance read from its LOWER
triangle, log-pdf = -(k log 2pi + log det + Mahalanobis)/2; a singular covariance raises
``numpy.linalg.LinAlgError`` as SciPy does with ``allow_singular=False``) and ``tf.transformations``
(``rosless/transformations.py``).
"""
from __future__ import annotations

import math

import numpy as np

from .rosless import transformations as _tft

_LOG_2PI = math.log(2.0 * math.pi)


def heading_of(orientation):
    """``utils.quaternion_to_heading`` (``utils.py:8-19``): yaw of a geometry_msgs Quaternion (or a 4-sequence)."""
    try:
        quat = [orientation.x, orientation.y, orientation.z, orientation.w]
    except AttributeError:
        quat = orientation
    return _tft.euler_from_quaternion(quat)[2]


def gaussian_pdf(x, mean, cov):
    """``scipy.stats.multivariate_normal.pdf(x, mean=mean, cov=cov)`` for one point."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    mean = np.asarray(mean, dtype=np.float64).reshape(-1)
    cov = np.asarray(cov, dtype=np.float64)
    k = x.shape[0]
    if cov.shape != (k, k) or mean.shape != (k,):
        raise ValueError("gaussian_pdf: dimension mismatch")
    s, u = np.linalg.eigh(cov, UPLO="L")
    eps = 1e6 * np.finfo(np.float64).eps * max(float(np.max(np.abs(s))), 0.0)
    if np.min(s) < -eps:
        raise ValueError("the input matrix must be positive semidefinite")
    if np.any(s <= eps):
        raise np.linalg.LinAlgError("singular matrix")
    dev = x - mean
    maha = float(np.sum(np.square(np.dot(dev, u)) / s))
    log_det = float(np.sum(np.log(s)))
    return float(math.exp(-0.5 * (k * _LOG_2PI + log_det + maha)))


class ParticleMathMixin(object):
    """Scalar helper methods of ``FilterParticle``; ``core.FilterParticle`` derives from this."""

    #: ``False``: ``find_nearest_reading`` behaves exactly as the reference is written -- it walks
    #: ``potential_features`` (``:579``, finding F5), so on a running filter it can only return ids <= 0 and
    #: ``add_hypothesis`` always stores an orphaned reading.  ``True``: the documented intent (SURVEY A.6 patch P1,
    #: what ``FastSLAM(spawn=True)`` runs on the device): walk ``hypothesis_set`` and accept the nearest
    #: reading when its colour distance is within ``pair_gate``.
    intent_pairing = False
    pair_gate = 300.0 ** 0.5

    # ---- association likelihood pieces (:457-544) ---------------------------------------------------------------
    def prob_position_match(self, f_mean, f_covar, s_x, s_y, bearing):
        """``:457-494``: 2-D Gaussian density, centred on the landmark, of the point of the observation ray that is
        closest to it; 0.0 when the landmark lies more than a quarter turn off the (robot-frame, finding F4c) bearing."""
        f_x, f_y = float(f_mean[0]), float(f_mean[1])
        if abs(math.atan2(f_y - s_y, f_x - s_x) - bearing) > math.pi / 2:
            return 0.0
        near = self.closest_point(f_x, f_y, s_x, s_y, bearing)
        return gaussian_pdf(near, (f_x, f_y), np.asarray(f_covar)[0:2, 0:2])

    def closest_point(self, f_x, f_y, s_x, s_y, obs_bearing):
        """``:496-522`` with ``utils.unit / dot_product / scale`` (``utils.py:37-81``) inlined: foot of the
        perpendicular from the landmark onto the ray, or the ray origin when the landmark is behind it."""
        c, s = math.cos(obs_bearing), math.sin(obs_bearing)
        length = math.sqrt(c * c + s * s + 0.0 * 0.0)
        if length < .0001:
            raise ZeroDivisionError("vector length 0 cannot be scaled to a unit vector")
        inv = 1.0 / length
        ux, uy, uz = c * inv, s * inv, 0.0 * inv
        along = (f_x - s_x) * ux + (f_y - s_y) * uy + 0.0 * uz
        if along < 0:
            return (s_x, s_y)
        return (float(s_x + ux * along), float(s_y + uy * along))

    def prob_color_match(self, f_mean, f_covar, blob):
        """``:524-544``: 3-D Gaussian density of the blob colour under the landmark's colour block."""
        seen = (blob.color.r, blob.color.g, blob.color.b)
        return gaussian_pdf(seen, (f_mean[2], f_mean[3], f_mean[4]), np.asarray(f_covar)[2:, 2:])

    # ---- new-landmark bookkeeping (:546-746) ---------------------------------------------------------------------
    def add_hypothesis(self, state, blob):
        """``:546-563``: pair the unseen blob with an earlier reading (-> potential landmark) or remember it."""
        pair_id = self.find_nearest_reading(state, blob)
        if pair_id > 0:
            self.add_new_feature(pair_id, state, blob)
        else:
            self.add_orphaned_reading(state, blob)

    def find_nearest_reading(self, state, blob):
        """``:565-590``: id of the stored reading nearest to (state, blob) under ``reading_distance_function``
        (first minimum in insertion order, strict ``<``); 0 when none intersects."""
        readings = self.hypothesis_set if self.intent_pairing else self.potential_features
        best_id, best = 0, float("inf")
        for id_, reading in readings.items():
            d = self.reading_distance_function(reading[0], reading[1], state, blob)
            if d < best:
                best, best_id = d, id_
        if self.intent_pairing:
            return best_id if best <= self.pair_gate else -best_id
        return best_id

    def reading_distance_function(self, state1, blob1, state2, blob2):
        """``:592-608``: colour distance of two readings whose world-frame rays cross, else infinity."""
        p1, p2 = state1.pose.pose.position, state2.pose.pose.position
        ray1 = blob1.bearing + heading_of(state1.pose.pose.orientation)
        ray2 = blob2.bearing + heading_of(state2.pose.pose.orientation)
        if not self.ray_intersect(p1.x, p1.y, ray1, p2.x, p2.y, ray2):
            return float("inf")
        return self.color_distance(blob1, blob2)

    def ray_intersect(self, x1, y1, b1, x3, y3, b3):
        """``:610-640``: do the half-lines (x1, y1, b1) and (x3, y3, b3) meet?  Solves a + u*ad = b + v*bd and asks
        for u >= 0 and v >= 0; parallel directions never do."""
        adx, ady = math.cos(b1), math.sin(b1)
        bdx, bdy = math.cos(b3), math.sin(b3)
        cross = ady * bdx - adx * bdy
        if cross == 0:
            return False
        v = (adx * y3 - ady * x3 + ady * x1 - adx * y1) / cross
        if abs(ady) < abs(adx):
            u = (x3 + bdx * v - x1) / adx
        else:
            u = (y3 + bdy * v - y1) / ady
        return u >= 0 and v >= 0

    def color_distance(self, blob1, blob2):
        """``:642-651``: Euclidean distance of two blob colours."""
        return math.sqrt(math.pow(blob1.color.r - blob2.color.r, 2) + math.pow(blob1.color.g - blob2.color.g, 2) +
                         math.pow(blob1.color.b - blob2.color.b, 2))

    def add_new_feature(self, old_id, state, blob):
        """``:653-680``: triangulate reading ``old_id`` with (state, blob) into ``potential_features[-next_id]``
        (mean = crossing point + averaged colour, identity covariance).  The consumed reading stays stored."""
        from .core import Feature, Matrix
        old_state, old_blob = self.hypothesis_set[old_id]
        crossing = self.cross_readings((old_state, old_blob), (state, blob,))
        mean = Matrix([crossing[0], crossing[1], (old_blob.color.r + blob.color.r) / 2,
                       (old_blob.color.g + blob.color.g) / 2, (old_blob.color.b + blob.color.b) / 2])
        self.potential_features[-self.next_id] = Feature(mean=mean, covar=Matrix(np.identity(5, dtype=int)))
        self.next_id += 1

    def cross_readings(self, old_reading, new_reading):
        """``:682-738``: crossing point of the two (infinite) lines through the readings' rays, by the determinant
        form of line-line intersection on two points per line one unit apart; ``None`` for parallel lines."""
        def two_points(reading):
            pos = reading[0].pose.pose.position
            ray = heading_of(reading[0].pose.pose.orientation) + reading[1].bearing
            return pos.x, pos.y, pos.x + math.cos(ray), pos.y + math.sin(ray)
        x1, y1, x2, y2 = two_points(old_reading)
        x3, y3, x4, y4 = two_points(new_reading)
        d12, d34 = x1 * y2 - y1 * x2, x3 * y4 - x4 * y3
        den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        if den == 0:
            return None
        return ((d12 * (x3 - x4) - (x1 - x2) * d34) / den, (d12 * (y3 - y4) - (y1 - y2) * d34) / den,)

    # ---- EKF pieces (:748-877) -----------------------------------------------------------------------------------
    def measurement_jacobian(self, feature_id):
        """``:748-802``: 4x5 Jacobian of (bearing, r, g, b) with respect to the landmark state, at the landmark
        mean; first row ``[+dy/q, +dx/q, 0, 0, 0]`` as the reference writes it (finding F4b), zeros when q == 0."""
        from .core import Matrix
        pos = self.state.pose.pose.position
        mean = self.get_feature_by_id(feature_id).mean
        dx, dy = mean[0] - pos.x, mean[1] - pos.y
        q = float(pow(dx, 2) + pow(dy, 2))
        row = []
        for num in (dy, dx):
            try:
                row.append(float(num) / q)
            except ZeroDivisionError:
                row.append(0.0)
        return Matrix([[row[0], row[1], 0.0, 0.0, 0.0],
                       [0.0, 0.0, 1.0, 0.0, 0.0],
                       [0.0, 0.0, 0.0, 1.0, 0.0],
                       [0.0, 0.0, 0.0, 0.0, 1.0]])

    def measurement_covariance(self, bigH, feature_id, Qt):
        """``:804-819``: Q = H Sigma H^T + Qt."""
        from .core import Matrix
        sigma = self.get_feature_by_id(feature_id).covar
        return Matrix(np.add(np.dot(np.dot(bigH, sigma), bigH.T), Qt))

    def kalman_gain(self, feature_id, bigH, Qinv):
        """``:821-833``: K = Sigma H^T Q^-1."""
        from .core import Matrix
        sigma = self.get_feature_by_id(feature_id).covar
        return Matrix(np.dot(np.dot(sigma, bigH.T), Qinv))

    def importance_factor(self, bigQ, blob, pseudoblob):
        """``:835-849``: (2 pi ||Q||_F)^(-1/2) exp(-innovation^T Q^-1 innovation / 2) -- Frobenius norm, not the
        determinant (finding F4d), no angle wrapping (F4e)."""
        from .core import _blob_to_matrix
        scale = pow(2.0 * math.pi * np.linalg.norm(bigQ), -0.5)
        innovation = _blob_to_matrix(blob) - _blob_to_matrix(pseudoblob)
        return scale * math.exp(-0.5 * np.dot(np.dot(innovation.T, np.linalg.inv(bigQ)), innovation))

    def generate_measurement(self, featureid):
        """``:859-877``: the blob this particle expects from landmark ``featureid``: WORLD-frame bearing (no heading
        subtraction, finding F4a) and the landmark's colour."""
        from ._ros import Blob
        pos = self.state.pose.pose.position
        mean = self.get_feature_by_id(featureid).mean
        expected = Blob()
        expected.bearing = math.atan2(mean[1] - pos.y, mean[0] - pos.x)
        expected.color.r, expected.color.g, expected.color.b = mean[2], mean[3], mean[4]
        return expected
