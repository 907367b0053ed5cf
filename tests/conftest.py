import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (dev container only)")


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name)


@pytest.fixture(scope="session")
def unit_vectors():
    import numpy as np
    return np.load(golden_path("unit_vectors.npz"))


TRACE_FIXTURES = ["trace_circle_m32_t60", "trace_corridor_m32_t60",
                  "trace_circle_immutable_m32_t40", "trace_corridor_noisy_m48_t40",
                  "trace_corridor_potential_m24_t12"]
if os.path.exists(golden_path("trace_c1_m100_n20_t500.npz")):
    TRACE_FIXTURES.append("trace_c1_m100_n20_t500")


def load_trace(name):
    import numpy as np
    return np.load(golden_path(name + ".npz"))


def scenario_from_trace(g):
    """Rebuild the Scenario a golden trace was generated from (inputs are stored in it)."""
    import numpy as np
    from parakeet_slam_b200.scenario import Scenario
    T = int(g["frames"])
    return Scenario(name=str(g["scenario"][0]), num_particles=int(g["num_particles"]),
                    num_landmarks=int(g["num_landmarks"]), obs_per_frame=int(g["obs_per_frame"]),
                    frames=T, v=float(g["v"]), w=float(g["w"]), dt=float(g["dt"]),
                    landmarks=g["landmarks"], true_poses=np.zeros((T, 3)),
                    observations=g["observations"], obs_landmark=np.zeros((T, 1), dtype=np.int64),
                    u01=g["u01"], motion_seed=int(g["motion_seed"]),
                    preset_covar=float(g["preset_covar"]), immutable=bool(g["immutable"]),
                    meta=dict(trajectory=str(g["trajectory"][0])))
