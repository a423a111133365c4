"""Orthorhombic periodic box (reference `pantea/atoms/box.py:14-126`).

Only the lattice diagonal is used; `apply_pbc` is the reference's single-shift minimum image
(`box.py:112-117`) and `wrap_into_box` its floored remainder (`box.py:123-126`). These two
elementwise host-API helpers are torch expressions on the resident device; the hot path
(neighbour search / ACSF / MD) applies the same arithmetic inside the CUDA kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from pantea_b200.types import Array, Dtype, as_torch_dtype, asarray, default_dtype


def _apply_pbc(dx: Array, lattice: Array) -> Array:
    box = torch.diagonal(lattice)
    dx = torch.where(dx > 0.5 * box, dx - box, dx)
    dx = torch.where(dx < -0.5 * box, dx + box, dx)
    return dx


def _wrap_into_box(positions: Array, lattice: Array) -> Array:
    return torch.remainder(positions, torch.diagonal(lattice))


@dataclass
class Box:
    lattice: Array

    @classmethod
    def from_list(cls, data: Sequence[float], dtype: Optional[Dtype] = None) -> "Box":
        dt = as_torch_dtype(dtype) if dtype is not None else default_dtype.FLOATX
        return cls(asarray(data, dtype=dt).reshape(3, 3))

    def apply_pbc(self, dx: Array) -> Array:
        return _apply_pbc(dx, self.lattice)

    def wrap_into_box(self, positions: Array) -> Array:
        return _wrap_into_box(positions, self.lattice)

    @property
    def lx(self) -> Array:
        return self.lattice[0, 0]

    @property
    def ly(self) -> Array:
        return self.lattice[1, 1]

    @property
    def lz(self) -> Array:
        return self.lattice[2, 2]

    @property
    def length(self) -> Array:
        return torch.diagonal(self.lattice)

    @property
    def volume(self) -> Array:
        return torch.prod(self.length)

    @property
    def dtype(self) -> Dtype:
        return self.lattice.dtype

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(lattice={self.lattice.tolist()}, dtype={self.dtype})"
