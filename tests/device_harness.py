"""Run the device filter over a Scenario and collect the same trace the oracle produces."""
import random as _pyrandom

import numpy as np

from parakeet_slam_b200.rosless import clock, messages
from parakeet_slam_b200.scenario import DT_NSEC, scan_from_observations


class View(object):
    def __init__(self):
        self.last_sensor_reading = None


def make_features(scn):
    from parakeet_slam_b200.core import Feature
    feats = []
    for row in scn.landmarks:
        f = Feature(mean=np.array(row, dtype=np.float64), covar=np.identity(5) * scn.preset_covar)
        f.__immutable__ = bool(scn.immutable)
        feats.append(f)
    return feats


def run_device(scn, dtype="f64", frames=None, num_particles=None, checkpoints=(), noise="numpy", potential_slots=(),
               spawn=False, known_map=True, capacity=None, orphan_capacity=32, arithmetic="f64",
               measurement_model="reference", weights="linear"):
    from parakeet_slam_b200.core import FastSLAM
    T = scn.frames if frames is None else frames
    M = scn.num_particles if num_particles is None else num_particles
    K = scn.obs_per_frame
    clock.set(0.0)
    fs = FastSLAM(make_features(scn) if known_map else [], num_particles=M, dtype=dtype, noise=noise, spawn=spawn,
                  capacity=capacity, orphan_capacity=orphan_capacity, arithmetic=arithmetic,
                  measurement_model=measurement_model, weights=weights)
    fs.keep_trace = True
    if len(potential_slots):
        # turn some preset landmarks into POTENTIAL features (id < 0), as potential_features[-id] of the reference
        mean5, covp, covc, meta, ids, nlive = fs.export_maps()
        sl = list(potential_slots)
        meta[:, sl] |= 0x20000000
        ids[:, sl] = -ids[:, sl]
        fs.import_maps(0, mean5, covp, covc, meta, ids)
    tw = messages.Twist()
    tw.linear.x = scn.v
    tw.angular.z = scn.w
    fs.last_control = tw
    np.random.seed(scn.motion_seed)
    _pyrandom.seed(scn.meta.get("resample_seed", 12345))
    tr = dict(pose_pre=np.zeros((T, M, 3)), pose_post=np.zeros((T, M, 3)),
              assoc=np.zeros((T, M, K), dtype=np.int32), weight=np.zeros((T, M)),
              ancestors=np.zeros((T, M), dtype=np.int64), summary=np.zeros((T, 3)),
              next_id=np.zeros((T, M), dtype=np.int64), lm_mean={}, lm_covp={}, lm_covc={}, lm_count={},
              stats=[])
    view = View()
    dt_ns = int(round(scn.dt * 1e9))
    assert dt_ns == DT_NSEC or True
    for t in range(T):
        clock.advance_nsec(dt_ns)
        view.last_sensor_reading = scan_from_observations(scn.observations[t])
        fs.cam_cb(view)
        tr["assoc"][t] = fs.last_assoc.cpu().numpy()
        tr["weight"][t] = fs.last_weight.cpu().numpy()
        tr["pose_pre"][t] = fs.last_pose_pre.cpu().numpy()
        tr["ancestors"][t] = fs.last_ancestors.cpu().numpy()
        tr["pose_post"][t] = fs.pose[:, :3].cpu().numpy()
        tr["summary"][t] = fs.summary()
        tr["next_id"][t] = fs.aux[:, 1].cpu().numpy()
        tr["stats"].append(fs.stats())
        if t in checkpoints:
            mean5, covp, covc, meta, ids, nlive = fs.export_maps()
            tr["lm_mean"][t], tr["lm_covp"][t], tr["lm_covc"][t] = mean5, covp, covc
            tr["lm_count"][t] = meta & 0x00FFFFFF
            tr.setdefault("lm_potential", {})[t] = (meta & 0x20000000) != 0
            live = np.arange(ids.shape[1])[None, :] < nlive[:, None]
            tr.setdefault("lm_ids", {})[t] = np.where(live, ids, 0)
            if spawn:
                tr.setdefault("orphans", {})[t] = fs.export_orphans()[0]
    tr["filter"] = fs
    return tr
