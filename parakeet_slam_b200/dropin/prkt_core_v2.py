"""Alias module: put this directory ahead of the reference's ``src/`` on ``sys.path`` and
``from prkt_core_v2 import FastSLAM, Feature`` (``prkt_ros.py:13``) resolves to the B200 core."""
from parakeet_slam_b200.core import FastSLAM, Feature, FilterParticle, Matrix  # noqa: F401
