from pantea_b200.atoms.box import Box
from pantea_b200.atoms.element import ElementMap
from pantea_b200.atoms.structure import Structure

__all__ = ["Structure", "Box", "ElementMap", "Neighbor", "calculate_distances"]


def __getattr__(name):  # lazy: these two pull in the CUDA library
    if name == "Neighbor":
        from pantea_b200.atoms.neighbor import Neighbor
        return Neighbor
    if name == "calculate_distances":
        from pantea_b200.atoms.distance import calculate_distances
        return calculate_distances
    raise AttributeError(name)
