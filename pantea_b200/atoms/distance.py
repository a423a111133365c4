"""`calculate_distances` (API of reference `pantea/atoms/distance.py:17-60`)."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from pantea_b200 import engine
from pantea_b200.atoms.neighbor import _workspace
from pantea_b200.types import Array


def calculate_distances(structure, atom_index: Optional[Array] = None, neighbor_atom_index: Optional[Array] = None,
                        with_aux: bool = False) -> Union[Array, Tuple[Array, Array]]:
    """Minimum-image distances (and optionally r_i - r_j) between two atom subsets of the structure."""
    ws = _workspace(structure)
    dev = structure.positions.device
    types = torch.ones(structure.natoms, dtype=torch.int32, device=dev)
    box = engine.box_lengths(structure)
    # the distance kernel only needs the packed records: any positive cutoff will do
    ws.bind(structure.positions, types, box, 1.0e-6, check=False)

    def _idx(i):
        return None if i is None else torch.atleast_1d(torch.as_tensor(i, device=dev)).to(torch.int32).contiguous()

    return ws.distances(_idx(atom_index), _idx(neighbor_atom_index), with_aux)
