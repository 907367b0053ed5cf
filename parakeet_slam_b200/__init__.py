"""parakeet_slam_b200 -- B200-native FastSLAM 1.0 particle-filter core."""
__version__ = "0.1.0"
