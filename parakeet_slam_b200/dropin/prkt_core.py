"""Alias module for the reference's v1 core (``src/prkt_core.py``).

The v1 module cannot be imported in the reference itself (SURVEY.md finding F10: ``utils.version`` raises at
class-definition time) and nothing there is parity-checkable; its *names* are kept so that code written against
them runs on the B200 core: ``ParticleMixedSlam`` is ``FastSLAM`` with v1's constructor defaults and method names
(``prkt_core.py:107-236``), each a thin route into the v2 pipeline (motion update -> fused association + EKF +
weight -> low-variance resample)."""
import numpy as np

from parakeet_slam_b200.core import FastSLAM, Feature, FilterParticle, Matrix  # noqa: F401


class ParticleMixedSlam(FastSLAM):
    """``ParticleMixedSlam()`` (``prkt_core.py:128-145``): M = 10 particles, empty map.  Keyword arguments are
    those of ``FastSLAM``."""

    def __init__(self, **kw):
        kw.setdefault("num_particles", 10)                    # self.M = 10  :133
        super(ParticleMixedSlam, self).__init__(kw.pop("preset_features", []), **kw)
        self.M = self.num_particles
        self.hypothesis_features = []                         # :143

    @property
    def robot_particles(self):                                # :134
        return self.particles

    @property
    def last_twist(self):                                     # :146
        return self.last_control

    @property
    def last_time(self):                                      # :145
        return self.last_update

    def measurement_update(self, measurement):
        """``:177-189``: dispatch on the message type; anything with a bearing and a colour is a camera observation.
        A ``[K, 4]`` array (host or device) goes straight to the fused kernel as in ``FastSLAM``."""
        if hasattr(measurement, "bearing") and hasattr(measurement, "color"):
            return self.cam_observation_update(measurement)
        if isinstance(measurement, np.ndarray) or hasattr(measurement, "is_cuda"):
            return FastSLAM.measurement_update(self, measurement)
        return None

    def cam_observation_update(self, cam_obs):
        """``:191-236``: one bearing-colour observation = motion update with the last twist, per-particle
        association / EKF / weight, resample."""
        zt = np.array([[cam_obs.bearing, cam_obs.color.r, cam_obs.color.g, cam_obs.color.b]], dtype=np.float64)
        with self._lock:
            self.motion_update(self.last_control)             # :196
            FastSLAM.measurement_update(self, zt)
            self.low_variance_resample()


SlamAlgorithm = FastSLAM          # base-class name of v1 (:25-105): motion_update / motion_model live on FastSLAM
RobotParticle = FilterParticle    # :293-339
FeatureModel = Feature            # :342-366
