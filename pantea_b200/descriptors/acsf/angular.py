from pantea_b200.descriptors.acsf.symmetry import G3, G9, AngularSymmetryFunction

__all__ = ["G3", "G9", "AngularSymmetryFunction"]
