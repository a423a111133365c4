"""pantea_b200 -- B200-native HDNNP energy/force hot path behind pantea's Python API."""
__version__ = "0.1.0"
