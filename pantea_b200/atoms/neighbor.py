"""Neighbour search (API of reference `pantea/atoms/neighbor.py:19-115`).

The reference stores a dense N x N boolean mask; here the CUDA cell-list / all-pairs kernel
produces CSR neighbour lists (`row_ptr`, `col`, ascending columns) and the dense `masks` view is
materialised lazily only when asked for.  Predicate: `(r <= r_cutoff) & (r > 0)` on the
single-shift minimum image (`neighbor.py:102-107`).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from pantea_b200 import engine
from pantea_b200.types import Array, asarray


class Neighbor:
    def __init__(self, r_cutoff: Array, row_ptr: Array, col: Array, natoms: int) -> None:
        self.r_cutoff = r_cutoff
        self.row_ptr = row_ptr
        self.col = col
        self.natoms = natoms
        self._masks: Optional[Array] = None

    @classmethod
    def from_structure(cls, structure, r_cutoff: float, with_aux: bool = False) -> Union["Neighbor", Tuple["Neighbor", Tuple[Array, Array]]]:
        ws = _workspace(structure)
        types = torch.ones(structure.natoms, dtype=torch.int32, device=structure.positions.device)
        ws.bind(structure.positions, types, engine.box_lengths(structure), float(r_cutoff))
        row_ptr, col = ws.neighbor_lists()
        nb = cls(asarray(float(r_cutoff), dtype=structure.dtype), row_ptr, col, structure.natoms)
        if with_aux:
            return nb, ws.distances(None, None, True)
        return nb

    @property
    def masks(self) -> Array:
        """Dense boolean [N, N] view of the lists (the reference's representation)."""
        if self._masks is None:
            n = self.natoms
            masks = torch.zeros((n, n), dtype=torch.bool, device=self.col.device)
            counts = (self.row_ptr[1:] - self.row_ptr[:-1])
            rows = torch.repeat_interleave(torch.arange(n, device=self.col.device), counts)
            masks[rows, self.col.long()] = True
            self._masks = masks
        return self._masks

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(r_cutoff={float(self.r_cutoff)})"


_WS_CACHE = {}


def _workspace(structure) -> engine.Workspace:
    """Potential-less workspace for pure geometry queries, cached per (dtype, capacity)."""
    key = (structure.dtype, max(64, structure.natoms))
    ws = _WS_CACHE.get(key)
    if ws is None:
        ws = engine.Workspace(None, key[1], min(max(structure.natoms - 1, 32), 1024), structure.dtype)
        _WS_CACHE[key] = ws
    return ws
