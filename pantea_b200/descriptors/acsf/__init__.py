from pantea_b200.descriptors.acsf.acsf import ACSF
from pantea_b200.descriptors.acsf.cutoff import CutoffFunction
from pantea_b200.descriptors.acsf.symmetry import G1, G2, G3, G9, NeighborElements

__all__ = ["ACSF", "CutoffFunction", "NeighborElements", "G1", "G2", "G3", "G9"]
