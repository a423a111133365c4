from pantea_b200.descriptors.acsf.symmetry import G1, G2, RadialSymmetryFunction

__all__ = ["G1", "G2", "RadialSymmetryFunction"]
